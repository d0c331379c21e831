// Internal launcher interface shared by the engine and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <stdexcept>
#include <string>

namespace dsocr {

enum class DType : int { F32 = 0, F16 = 1, BF16 = 2 };  // values match include/dsocr.h

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
// Every kernel launch of this library goes through launch_check: it surfaces launch errors and counts the
// launches (dsocr_launch_count).
std::atomic<long long>& launch_counter();
void launch_check(const char* what);
// One-time set-up that CUDA keeps PER DEVICE (cudaFuncSetAttribute): engines on several GPUs may live in one process
// and be driven from several threads, so "configured" is tracked per device ordinal under a lock.
struct PerDeviceOnce {
  std::mutex m;
  unsigned long long mask = 0;
  template <typename F>
  void run(F&& f) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(m);
    if (!((mask >> (dev & 63)) & 1ull)) { f(); mask |= 1ull << (dev & 63); }
  }
};
// Optional per-kernel timing: while enabled, launch_check records a CUDA event after every launch on the
// engine stream; consecutive event deltas are aggregated by kernel name (kernels on one stream serialise).
void kernel_timing_begin(cudaStream_t stream);
void kernel_timing_phase(const char* phase);  // prefix for subsequently recorded kernels (string literal)
std::string kernel_timing_end_json();

// ------------------------------------------------------------------------- tensor-core linear
struct LinearTile { int w_row0, x_row0, rows, n0, group, r0; };  // == lin::Tile

struct QuantWeight;
struct LinearCall {
  // weights: [w_rows, K] 16-bit, row pitch ldw (0 = K).  w1 != nullptr selects the dual (SwiGLU) kernel.
  const void* w0 = nullptr;
  const void* w1 = nullptr;
  long long w_rows = 0;  // 0 = N
  long long ldw = 0;
  bool w_tiled = false;  // w0/w1 are in the pre-tiled streaming layout (retile_weights); rows padded to 128
  // DSQ snapshot weights (Q8_0 / Q4_K / Q6_K / f32 fallback planes, dsq.h) instead of w0 / w1: the dequant-fused kernel
  // of linear_dq.cuh; needs x_parts == 2, no batched X and no device-scheduled groups
  const QuantWeight* q0 = nullptr;
  const QuantWeight* q1 = nullptr;
  // activations: [x_rows, K] 16-bit, row pitch ldx (0 = K); x_parts = 2 adds the lo part at row x_lo_row_off
  const void* x = nullptr;
  long long x_rows = 0;
  long long ldx = 0;
  int x_parts = 1;
  int x_lo_row_off = 0;
  // batched X (3-D): nbatch slices, x_batch_stride elements apart (e.g. attention heads inside a qkv row)
  int nbatch = 1;
  long long x_batch_stride = 0;
  long long out_batch_stride = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  void* out = nullptr;
  void* out_lo = nullptr;
  long long ldo = 0;
  const int* row_map = nullptr;
  int act = 0;       // lin::Act
  int out_mode = 0;  // lin::Out
  // grouped problems
  const LinearTile* tiles = nullptr;
  const int* num_tiles_dev = nullptr;
  const int* group_counts = nullptr;  // fixed-capacity groups: tile rows derived from device-side counts
  // fixed-capacity groups scheduled on the device: group g = weight rows [g*N, +N), token rows [g*dyn_cap, +count)
  int dyn_groups = 0;
  int dyn_cap = 0;
  // stream-K workspace for the device-scheduled groups: num_sms slots of 2*128*128 floats and 2*num_sms ints that are
  // zero before the first launch (the kernel hands them back zeroed); see linear_streamk_ws_bytes
  float* sk_ws = nullptr;
  int* sk_flags = nullptr;
  int max_tiles = 0;
  int tile_rows_hint = 0;  // typical rows per tile, used to pick the token tile
  int bn = 0;              // force the token tile (0 = auto)
  const char* tag = nullptr;  // kernel name for timing reports
  // deterministic split-K: k_splits partial f32 results at out + s*split_stride (out_mode must be OUT_F32 or
  // OUT_F32_DUAL); see linear_plan_splits
  int k_splits = 1;
  long long split_stride = 0;
  long long dual_stride = 0;
};
// number of k-splits that fills the GPU for a small-M problem (1 = do not split)
int linear_plan_splits(long long M, int N, int K, int num_sms, bool dual = false);

int linear_pick_bn(long long m, bool dual);
// Re-lays a row-major 16-bit weight [n, k] (k % 64 == 0) out as 128x64 tiles, each 16 KB contiguous and already in
// the 128B-swizzled shared-memory order (tile index = row_block * k/64 + k_block); rows are padded to a multiple of
// 128 with zeros.  dst needs retiled_bytes(n, k).
inline size_t retiled_bytes(long long n, long long k) { return (size_t)((n + 127) / 128) * 128 * (size_t)k * 2; }
void retile_weights(const void* src, void* dst, long long n, int k, cudaStream_t s);
inline size_t linear_streamk_ws_bytes(int num_sms) { return (size_t)num_sms * 2 * 128 * 128 * 4; }
void linear(const LinearCall& c, DType dt, int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------------- vision attention
// qkv: [rows = B*S, 3, H, 64] 16-bit.  grid > 0: tokens form a grid x grid raster per sequence and the
// decomposed rel-pos bias is added from Z = q . [rel_h ; rel_w]^T  ([rows, H, zw] f32, Zw half at zhalf).
struct VAttnCall {
  const void* qkv = nullptr;
  long long rows = 0;
  int B = 0, S = 0, H = 0;
  int grid = 0;
  const float* Z = nullptr;
  int zw = 0, zhalf = 0;
  void* out = nullptr;  // [rows, H*64] 16-bit
  float scale = 0.125f;
  const char* tag = nullptr;
};
bool vision_attention_supported(int grid);
void vision_attention(const VAttnCall& c, DType dt, cudaStream_t stream);
// Z = q . table^T through the tensor-core linear kernel.  table: [2*zhalf, 64] 16-bit (rel_h rows then rel_w rows).
void vision_relpos_products(const void* qkv, long long rows, int H, const void* table, int zhalf, float* Z, DType dt,
                            int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------------- vision SIMT kernels
void patchify_u8(const uint8_t* img, void* out, int B, int G, DType dt, cudaStream_t s);
void patchify_f32(const float* img, void* out, int B, int G, DType dt, cudaStream_t s);
void bcast_rows(const float* src, float* dst, long long rows_per_batch, int batch, int cols, cudaStream_t s);
void layernorm(const float* x, const float* w, const float* b, void* out16, float* out32, long long out_rows, int C,
               float eps, int win, int g, int nw, DType dt, cudaStream_t s);
void window_row_map(int* map, long long rows, int win, int g, int nw, cudaStream_t s);
void cast16(const float* x, void* out, long long n, DType dt, cudaStream_t s);
void im2col3x3(const void* in, void* out, int B, int Hin, int Win, int C, int stride, DType dt, cudaStream_t s);
void clip_embed(const float* sam, const float* cls, const float* pos, float* out, int B, int n, int C, cudaStream_t s);
void concat_clip_sam(const float* clip, const float* sam, void* out16, float* out32, int B, int n, int C, DType dt,
                     cudaStream_t s);
void scatter_tokens(const float* proj, const float* newline, const float* sep, const int* map, float* dst,
                    long long rows, int C, cudaStream_t s);

void resample_h(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, const int* start, const int* len,
                const int* coef, int ksize, cudaStream_t s);
void resample_v(const uint8_t* horiz, int dw, int dh, uint8_t* dst, const int* start, const int* len, const int* coef,
                int ksize, int canvas, int x_off, int y_off, int tile, int tiles_w, cudaStream_t s);

// ------------------------------------------------------------------------- decoder SIMT kernels
void embed_gather(const int* src, const void* table, const float* img_rows, float* out, long long rows, int H, DType dt,
                  cudaStream_t s);
void rmsnorm_split(float* x, const float* w, void* out16, long long lo_off_elems, float* out32,
                   const int* row_idx, long long rows, int H, float eps, const float* partials, int n_splits,
                   long long split_stride, DType dt, cudaStream_t s);
void swiglu_reduce(const float* part, int n_splits, long long split_stride, long long dual_stride, void* out16,
                   long long lo_off_elems, long long n, DType dt, cudaStream_t s);
void rope_kv(const float* qkv, const float* cos_t, const float* sin_t, const int* row_page, const int* row_pos,
             float* q_out, void* kc, void* vc, bool kv_f16, long long rows, int heads, int smax, int n_splits,
             long long split_stride, cudaStream_t s);
// page_spans: optional int[2 * n_pages] workspace; with it a prefill batch (rows > 256, rows ordered by page then
// position) runs the shared-memory tiled kernel
void kv_attention(const float* q, const void* kc, const void* vc, bool kv_f16, const int* row_page, const int* row_pos,
                  void* ctx, long long lo_off_elems, float* ctx32, long long rows, int heads, int smax, float scale,
                  DType dt, cudaStream_t s, int* page_spans = nullptr, int n_pages = 0);
void moe_router(const float* x, const float* wgt, int* topk_idx, float* topk_w, int* counts, long long rows, int H,
                int E, int topk, cudaStream_t s);
void moe_plan(int* counts, int* offsets, int* cursor, LinearTile* tiles1, int* ntiles1, LinearTile* tiles2,
              int* ntiles2, int E, int bn, int N1, int N2, cudaStream_t s);
void moe_dispatch(const int* topk_idx, const int* offsets, int* cursor, const void* xn, long long xn_lo_off,
                  void* xperm, long long xperm_lo_off, int* perm_pos, long long n_assign, int topk, int H, DType dt,
                  cudaStream_t s);
void moe_combine(const float* y, const int* perm_pos, const float* topk_w, float* x, long long rows, int topk, int H,
                 const float* partials, int n_splits, long long split_stride, cudaStream_t s);
void select_token(const float* logits, int V, int* hist, int hist_stride, int* hist_len, int* gen_count, int* finished,
                  int n_pages, int ngram, float penalty, int eos, int max_new, const int* forced, int forced_stride,
                  int* selected_out, int selected_stride, float* scratch, cudaStream_t s);
constexpr int kSelectScratchPerPage = 16 * 3 * 2 + 1;  // 32-bit words of select_token scratch per page
// host-selected tokens (sampling path): same bookkeeping as select_token's tail (append / EOS freeze / budget)
void append_tokens(const int* chosen, int* hist, int hist_stride, int* hist_len, int* gen_count, int* finished, int n_pages,
                   int eos, int max_new, cudaStream_t s);
bool kernel_timing_enabled();
void decode_rows(const int* hist, int hist_stride, const int* hist_len, int* src, int* row_pos, int n_pages,
                 cudaStream_t s);
void fill_i32(int* p, int v, long long n, cudaStream_t s);
void moe_active_stat(const int* counts, int n, unsigned long long* stats, cudaStream_t s);
// decode-step fusions (rows <= 256)
void rope_attn_decode(const float* qkv, int n_splits, long long split_stride, const float* cos_t, const float* sin_t,
                      void* kc, void* vc, bool kv_f16, const int* row_page, const int* row_pos, void* ctx,
                      long long lo_off_elems, long long rows, int heads, int smax, float scale, DType dt, cudaStream_t s);
// Expert-parallel decode (BASELINE configs[4]): `world` engine replicas, rank r owns the routed experts
// [r*eloc, (r+1)*eloc).  Every table entry is the address, valid on THIS device, of the named buffer of rank i
// (peer mappings over NVLink / NVSwitch; the rank's own buffers for i == rank).
//   counts[i] : int  [layers][eloc + n_shared]      rows that arrived in each local expert segment of rank i
//   xperm[i]  : T    [2 (hi, lo)][(eloc + n_shared) * cap][H]   token rows dispatched to rank i's experts
//   y[i]      : f32  [(eloc + n_shared) * cap][H]   expert outputs of rank i
//   flags[i]  : int  [world]                        arrival generation of every rank at rank i's barrier
struct EpPeers {
  int world = 1, rank = 0, eloc = 0, cap = 0;
  int* counts[8] = {};
  void* xperm[8] = {};
  float* y[8] = {};
  int* flags[8] = {};
};
// cross-GPU barrier of the EP group on the engines' streams: every rank's prior device work (peer stores included) is
// visible to every rank's subsequent kernels.  gen: device counter of this rank (one increment per barrier).
void ep_barrier(const EpPeers& ep, int* gen, cudaStream_t s);
void post_attn(float* x, const float* partials, int n_splits, long long split_stride, const float* w, const float* wgt,
               void* xn16, long long xn_lo_off, int* topk_idx, float* topk_w, int* counts, int* perm_pos, void* xperm,
               long long xperm_lo_off, int cap, long long rows, int H, int E, int topk, int n_shared, float eps, DType dt,
               cudaStream_t s, const EpPeers* ep = nullptr, int ep_counts_off = 0);
void combine_norm(float* x, const float* y, const int* perm_pos, const float* topk_w, int topk, const float* partials,
                  int n_splits, long long split_stride, const float* w_next, void* out16, long long lo_off_elems,
                  long long rows, int H, float eps, int n_shared, int shared_row0, int cap, DType dt, cudaStream_t s,
                  const EpPeers* ep = nullptr);

}  // namespace dsocr
