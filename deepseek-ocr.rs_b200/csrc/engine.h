// Engine: device-resident weights + the per-page forward path (vision encode, prefill, batched decode).
// Mirrors DeepseekOcrModel (crates/infer-deepseek/src/model/mod.rs) behind the C ABI of include/dsocr.h.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <functional>
#include <string>
#include <vector>

#include "dsocr.h"
#include "dsq.h"
#include "kernels.h"
#include "util.h"

namespace dsocr {

struct ModelConfig {
  // SAM ViT-B (vision/sam.rs:44-111)
  int sam_image = 1024, sam_patch = 16, sam_dim = 768, sam_depth = 12, sam_heads = 12, sam_window = 14;
  int sam_neck = 256, sam_out0 = 512, sam_out1 = 1024;
  std::vector<int> sam_global = {2, 5, 8, 11};
  // CLIP-L (vision/clip.rs:35-52)
  int clip_dim = 1024, clip_layers = 24, clip_heads = 16, clip_image = 224, clip_patch = 14;
  // projector
  int proj_in = 2048, n_embed = 1280;
  // decoder
  int vocab = 129280, hidden = 1280, layers = 12, heads = 10, inter = 6848, moe_inter = 896;
  int n_experts = 64, n_shared = 2, topk = 6, first_dense = 1;
  float rope_theta = 10000.f, rms_eps = 1e-6f;
  int eos = 1;
  int head_dim() const { return hidden / heads; }
  bool sam_is_global(int i) const { for (int g : sam_global) if (g == i) return true; return false; }
};

struct SamBlockW {
  DevBuf ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
  std::vector<float> rel_h, rel_w;  // host copies [rel_rows, 64]
  int rel_rows = 0;
  std::map<int, DevBuf> rel_table;  // token-grid size -> 16-bit [2*zhalf, 64] ([rel_h ; rel_w])
  // decomposed rel-pos logits as extra output features of the block input: Z = q . tab^T = xn . (tab Wq)^T + tab bq
  std::vector<float> q_w_host, q_b_host;  // query rows of attn.qkv, values as rounded to the engine dtype: [D, D], [D]
  struct RelFused { DevBuf w, b; };       // 16-bit [heads*2*zhalf, D], f32 [heads*2*zhalf]
  std::map<int, RelFused> rel_fused;      // token-grid size -> fused weight
};
struct ClipBlockW {
  DevBuf ln1_w, ln1_b, qkv_w, qkv_b, out_w, out_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
};
struct DecLayerW {
  DevBuf ln1, ln2, qkv_w, o_w;
  // dense
  DevBuf gate_w, up_w, down_w;
  // MoE
  DevBuf router_wt;                          // f32 [H, E] (transposed gate weight)
  DevBuf exp_gate, exp_up, exp_down;         // [E*mi, H], [E*mi, H], [E*H, mi]
  DevBuf sh_gate, sh_up, sh_down;            // shared experts fused: [S, H], [S, H], [H, S]
  bool moe = false;
  // DSQ-quantised variants (engine created with a snapshot): q/k/v/o, dense MLP, stacked experts, shared experts
  QuantWeight q_q, q_k, q_v, q_o, q_gate, q_up, q_down, q_exp_gate, q_exp_up, q_exp_down, q_sh_gate, q_sh_up, q_sh_down;
};

struct Timings { double prepare = 0, vision = 0, prefill = 0, iterative = 0, generate = 0; };

class Engine {
 public:
  Engine(const std::string& config_path, const std::string& weights_path, const std::string& dsq_path, int device,
         DType dtype);
  ~Engine();

  // compute_image_embeddings for a batch of pages.  Views are 16-bit patch rows already gathered on the device.
  struct PageViews { int n_tiles = 0; int crop_w = 1, crop_h = 1; };
  // globals: u8 [n_pages, G, G, 3] device or f32 [n_pages,3,G,G] device; tiles likewise (concatenated).
  // Output: rows_all = concatenated [sum n_rows, hidden] f32 rows in page order (engine-owned workspace,
  // valid until the next vision_encode) and the per-page row counts.
  const float* vision_encode(int n_pages, const void* globals_dev, bool globals_f32, int G, const void* tiles_dev,
                             bool tiles_f32, int P, const std::vector<PageViews>& pages, std::vector<int>* n_rows);

  struct GenRequest {
    int n_pages = 0;
    const int64_t* const* input_ids = nullptr;
    const uint8_t* const* mask = nullptr;
    const int* n_tokens = nullptr;
    // image rows: host pointers per page, or one device buffer of the concatenated rows in page order
    const float* const* image_rows_host = nullptr;
    const float* image_rows_dev = nullptr;
    const int* n_image_rows = nullptr;
    dsocr_decode_params params{};
    const int64_t* const* forced = nullptr;  // teacher forcing
    int n_forced_steps = 0;
    float* const* logits_out = nullptr;
    dsocr_token_cb cb = nullptr;
    void* user = nullptr;
    int page_offset = 0;  // added to the page index the callback reports (calls that are one group of a larger batch)
  };
  void generate(const GenRequest& rq, int64_t* const* out_tokens, int* n_out);

  int tap(const std::string& name, float* out, size_t capacity, size_t* n_written);
  void set_record_taps(bool on) { record_taps_ = on; }
  void set_kv_f16(bool on) { kv_f16_ = on; }  // KV cache storage: f32 (reference semantics, default) or f16
  void set_moe_stats(bool on);
  // diagnostics: {sum over decode steps of non-empty (layer, expert) segments, decode steps counted}; resets them
  void moe_stats(unsigned long long out[2]);
  void set_stream(cudaStream_t s);  // adopt a caller-owned stream (e.g. torch's current stream)

  const ModelConfig& cfg() const { return cfg_; }
  DType dtype() const { return dt_; }
  int device() const { return device_; }
  int sm_count() const { return num_sms_; }
  cudaStream_t stream() const { return stream_; }
  // optional: called by the vision tower before it reads the views [.., last_view] of a chunk (local = crop tiles);
  // the C ABI uses it to wait for pages that are still being copied in on a side stream
  std::function<void(bool local, int last_view)> view_ready;
  Timings timings;
  std::string device_name;

 private:
  void load_weights(const std::string& path, const DsqReader* dsq);
  void decoder_forward_dsq(float* x, long long rows, const int* row_page, const int* row_pos, int smax,
                           const int* final_rows, int n_final, float* logits);
  // one decode step for <= 4 pages with 6 launches per layer (dsq_decode.cu), DSQ snapshot or 16-bit weights
  void decoder_step_fused_small(float* x, long long rows, const int* row_page, const int* row_pos, int smax, float* logits);
 public:
  bool quantized() const { return quantized_; }
  // ---- expert-parallel decode (BASELINE configs[4]; csrc/capi.cpp dsocr_ep_group_create wires a group of engines up)
  // ep_attach allocates this rank's shared buffers (segment counters, dispatched rows, expert outputs, barrier flags) and
  // the contiguous weight stacks of its local experts; ep_set_peers installs the peer address tables.
  struct EpBuffers { int* counts; void* xperm; float* y; int* flags; };
  EpBuffers ep_attach(int rank, int world, int cap);
  void ep_set_peers(const EpPeers& peers);
  void ep_detach();
  int ep_world() const { return ep_peers_.world; }
 private:
  void sam_forward(int Bv, int G, const void* patches16, float* sam_out);
  void clip_forward(int Bv, int g3, const float* sam_out, float* clip_out);
  void vision_views(int Bv, int G, const void* img_dev, bool is_f32, float* proj_out /*[Bv*n,1280]*/, const char* tag);
  const float* sam_pos_for(int g);
  const float* clip_pos_for(int g3);
  const void* rel_table_for(int layer, int size, int* zhalf);
  const SamBlockW::RelFused& rel_fused_for(int layer, int size, int* zhalf);
  void decoder_forward(float* x, long long rows, const int* row_page, const int* row_pos, int smax,
                       const int* final_rows, int n_final, float* logits, bool decode_mode);
  void record_tap(const std::string& name, const float* dev, size_t n);
  void record_tap16(const std::string& name, const void* dev, size_t n);
  DevBuf& ws(const std::string& name, size_t bytes);

  ModelConfig cfg_;
  DType dt_;
  int device_ = 0;
  int num_sms_ = 0;
  cudaStream_t stream_ = nullptr;
  bool owns_stream_ = true;
  cudaStream_t stream2_ = nullptr;  // side branch (shared experts) forked from stream_
  cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
  bool record_taps_ = false;
  bool kv_f16_ = false;
  bool moe_stats_ = false;
  DevBuf moe_stats_dev_;
  bool w_tiled_ = getenv("DSOCR_NO_TILED") == nullptr;  // decoder weights in the pre-tiled streaming layout
  void retile_inplace(DevBuf& w, long long n, int k);
  bool streamk_ = getenv("DSOCR_NO_STREAMK") == nullptr;  // A/B switch: balanced static units instead
  // largest decode batch (pages per step) that takes the fused decode schedule (post_attn / stream-K experts / combine_norm)
  int fused_max_rows_ = getenv("DSOCR_FUSED_MAX_ROWS") ? atoi(getenv("DSOCR_FUSED_MAX_ROWS")) : 1024;
  DevBuf sk_ws_, sk_flags_;  // stream-K partial slots + hand-off flags of the decode-time expert GEMMs
  bool quantized_ = false;
  // DSQ engines: prefill and decode steps of > 4 pages run the dequant-fused tensor-core GEMM (linear_dq.cuh) through
  // decoder_forward; DSOCR_DSQ_GEMV=1 switches back to the per-row GEMVs of decoder_forward_dsq (A/B)
  bool dsq_gemm_ = getenv("DSOCR_DSQ_GEMV") == nullptr;
  // A/B switch: decode steps of <= 4 pages through the batched kernels (float engine) / per-linear GEMVs (DSQ)
  bool small_fused_ = getenv("DSOCR_DSQ_UNFUSED") == nullptr && getenv("DSOCR_NO_SMALL_FUSED") == nullptr;
  QuantWeight q_lm_head_;
  long long iota_n_ = 0;
  std::map<std::string, std::vector<float>> taps_;

  // weights
  DevBuf patch_w_, patch_b_;
  std::vector<float> sam_pos_host_;
  std::map<int, DevBuf> sam_pos_;
  std::vector<SamBlockW> sam_;
  DevBuf neck0_w_, neck1_w_, neck1_b_, neck2_w_, neck3_w_, neck3_b_, net2_w_, net3_w_;
  DevBuf clip_cls_, clip_preln_w_, clip_preln_b_;
  std::vector<float> clip_pos_host_;
  std::map<int, DevBuf> clip_pos_;
  std::vector<ClipBlockW> clip_;
  DevBuf proj_w_, proj_b_, newline_, separator_;
  DevBuf embed_, final_norm_, lm_head_;
  std::vector<DecLayerW> dec_;
  DevBuf rope_cos_, rope_sin_;
  int rope_len_ = 0;

  std::map<std::string, DevBuf> ws_;
  EpPeers ep_peers_;                       // world == 1: not expert parallel
  int ep_cap_ = 0;                         // rows per expert segment (pages per rank x world)
  DevBuf ep_counts_, ep_xperm_, ep_hperm_, ep_y_, ep_flags_, ep_gen_;
  std::vector<DevBuf> ep_gate_, ep_up_, ep_down_;  // per layer: [eloc + n_shared] weight groups of this rank
  std::vector<DevBuf> kcache_, vcache_;
  // KV cache of page p of the current pass lives at page (kv_page_base_ + p) of the call's cache (chunked prefill)
  size_t kv_page_bytes_ = 0;
  int kv_page_base_ = 0;
  void* kc_ptr(int l) const { return (char*)kcache_[l].p + (size_t)kv_page_base_ * kv_page_bytes_; }
  void* vc_ptr(int l) const { return (char*)vcache_[l].p + (size_t)kv_page_base_ * kv_page_bytes_; }
};

}  // namespace dsocr
