// DSQ snapshot container reader (crates/dsq/src/lib.rs:14-15, 60-110, 208-306, 314-391) and the device-side
// representation of a quantised linear layer.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "util.h"

namespace dsocr {

enum class DsqDType : uint32_t { F32 = 0, F16 = 1, Q8_0 = 8, Q4K = 12, Q6K = 14, BF16 = 16 };

struct DsqRecord {
  std::string name;
  uint32_t out_dim = 0, in_dim = 0;
  DsqDType q_dtype = DsqDType::Q8_0;
  uint64_t q_offset = 0, q_len = 0;
  uint64_t bias_offset = 0, bias_len = 0;
  uint32_t bias_dtype = 0;
  bool has_bias = false;
};

class DsqReader {
 public:
  explicit DsqReader(const std::string& path);
  ~DsqReader();
  DsqReader(const DsqReader&) = delete;
  const DsqRecord* find(const std::string& name) const;
  const uint8_t* bytes(const DsqRecord& r) const { return base_ + r.q_offset; }
  const std::vector<DsqRecord>& records() const { return records_; }
  DsqDType default_dtype() const { return default_dtype_; }
  uint32_t block_size() const { return block_size_; }
  size_t file_size() const { return size_; }
  std::string model_id, backend, candle_version;

 private:
  int fd_ = -1;
  size_t size_ = 0;
  const uint8_t* base_ = nullptr;
  DsqDType default_dtype_ = DsqDType::Q8_0;
  uint32_t block_size_ = 0;
  std::vector<DsqRecord> records_;
  std::map<std::string, size_t> index_;
};

// DsqWriter (crates/dsq-writer/src/lib.rs:93-527): collects tensors and emits header | records | payload.  Host only.
class DsqWriter {
 public:
  DsqWriter(const std::string& path, const std::string& candle_version, const std::string& model_id, const std::string& backend,
            DsqDType default_dtype);
  // quantise (Q8_0 / Q4_K / Q6_K) or convert (F32 / F16 / BF16) a row-major f32 matrix [out_dim, in_dim]; bias: out_dim f32 or null
  void add_tensor_f32(const std::string& name, uint32_t out_dim, uint32_t in_dim, DsqDType dt, const float* w, const float* bias);
  // blocks quantised elsewhere (add_quantized_bytes, :366-409)
  void add_quantized_bytes(const std::string& name, uint32_t out_dim, uint32_t in_dim, DsqDType dt, const uint8_t* q, size_t q_len,
                           const float* bias);
  void finalize();
  const std::string& path() const { return path_; }

 private:
  struct Pending {
    std::string name;
    uint32_t out_dim = 0, in_dim = 0;
    DsqDType dtype = DsqDType::Q8_0;
    uint64_t q_offset = 0, q_len = 0, bias_offset = 0, bias_len = 0;
    bool has_bias = false;
  };
  void check_new(const std::string& name, DsqDType dt, uint32_t in_dim) const;
  void append(const std::string& name, uint32_t out_dim, uint32_t in_dim, DsqDType dt, const uint8_t* q, size_t q_len, const float* bias);
  std::string path_, candle_version_, model_id_, backend_;
  DsqDType default_dtype_;
  std::vector<uint8_t> payload_;
  std::vector<Pending> records_;
};

inline int dsq_block_elems(DsqDType t) { return t == DsqDType::Q8_0 ? 32 : (t == DsqDType::Q4K || t == DsqDType::Q6K) ? 256 : 0; }
inline int dsq_block_bytes(DsqDType t) { return t == DsqDType::Q8_0 ? 34 : t == DsqDType::Q4K ? 144 : t == DsqDType::Q6K ? 210 : 0; }

// Device storage of one quantised weight matrix [N, K] (or E stacked matrices).  The on-disk ggml blocks are
// re-laid out at load time into 16-byte-friendly planes (same byte count):
//   Q8_0 : qs  int8 [N][K]          | d f16 [N][K/32]
//   Q4_K : blk [N][K/256][144]      (as on disk: f16 d, f16 dmin, u8 scales[12], u8 qs[128])
//   Q6_K : ql u8 [N][K/2] | qh u8 [N][K/4] | sc i8 [N][K/16] | d f16 [N][K/256]
//   F16/BF16/F32 records: converted to f32 [N][K] (float fallback of the exporter's chain)
struct QuantWeight {
  DsqDType fmt = DsqDType::Q8_0;
  long long N = 0;  // rows of one matrix (experts are stacked: total rows = N * count)
  int K = 0;
  int count = 1;
  DevBuf a, b, c, d;  // planes as listed above (a is the main plane)
  size_t bytes() const { return a.bytes + b.bytes + c.bytes + d.bytes; }
};

// Host-side repacking of `rows` rows of on-disk blocks into the planes of `dst` starting at row `row0`.
void dsq_upload_rows(QuantWeight& dst, long long row0, const uint8_t* src, DsqDType src_fmt, long long rows);
void dsq_alloc(QuantWeight& w, DsqDType fmt, long long N, int K, int count);

// out[r, n] (+)= sum_k x[xr(r), k] * dequant(W[e(r)])[n, k]   with xr(r) = r / x_row_div and
// e(r) = row_expert ? row_expert[r] : 0.  f32 activations, f32 accumulation.
struct DsqGemvCall {
  const QuantWeight* w = nullptr;
  const float* x = nullptr;
  long long ldx = 0;
  int x_row_div = 1;
  const int* row_expert = nullptr;
  float* out = nullptr;
  long long ldo = 0;
  long long rows = 0;
  bool accumulate = false;
  const char* tag = "dsq_gemv";
};
void dsq_gemv(const DsqGemvCall& c, cudaStream_t stream);

// ---- fused small-batch decode step (dsq_decode.cu)
// One launch runs up to 3 GEMV jobs.  A job's token rows come in `groups` of `rpg` (<= 4) rows that share one weight
// matrix (expert = row_expert[group] or 0); x row of (group g, m) = (g*rpg + m) / x_row_div; out row = g*rpg + m.
// w1 != nullptr: out = silu(x.w0^T) * (x.w1^T).
// Weight matrix of a job: a DSQ QuantWeight, or a 16-bit matrix in the engine's pre-tiled streaming layout
// (retile_weights: 128x64 tiles of 16 KB, rows 128-byte swizzled) - the float engine's decoder weights.
struct FusedWeight {
  int fmt = -1;  // 8 / 12 / 14 / 0 (f32): QuantWeight planes; 16 / 17: tiled f16 / bf16
  long long N = 0;
  int K = 0;
  const void* p[4] = {nullptr, nullptr, nullptr, nullptr};
  bool valid() const { return fmt >= 0; }
};
FusedWeight fused_weight(const QuantWeight& w);
FusedWeight fused_weight_tiled16(const void* tiled, long long N, int K, bool bf16);
struct DsqFusedJob {
  FusedWeight w0, w1;
  const float* x = nullptr;
  long long ldx = 0;
  int groups = 1, rpg = 1, x_row_div = 1;
  const int* row_expert = nullptr;
  bool expert_dep = false;  // row_expert is produced by the kernel launched immediately before this one
  float* out = nullptr;
  long long ldo = 0;
};
// How every block produces its activation rows before the GEMV (all jobs of the launch must then share x and K):
// x = base + (sum_j wmoe[r*topk+j] * ymoe[r*topk+j] + add1 + add2); block 0 stores x to write_back (a buffer no job reads);
// with norm_w the rows are RMS-normalised (rms_norm, f32) and scaled by norm_w.
struct DsqFusedStage {
  const float* add1 = nullptr;
  const float* add2 = nullptr;
  const float* ymoe = nullptr;
  const float* wmoe = nullptr;
  int topk = 0;
  float* write_back = nullptr;
  const float* norm_w = nullptr;
  float eps = 0.f;
};
void dsq_fused_gemv(const DsqFusedJob* jobs, int njobs, const DsqFusedStage& st, const char* tag, cudaStream_t stream);
// logits_ws: float[rows*E]; counters: int[rows], zero before the first launch (handed back zeroed)
void dsq_router(const float* base, const float* add1, float* xout, const float* w, const float* wgt, float* xn32,
                float* logits_ws, int* counters, int* topk_idx, float* topk_w, long long rows, int H, int E, int topk,
                float eps, cudaStream_t s);
void dsq_combine_norm(const float* base, const float* ymoe, const float* wmoe, int topk, const float* add1,
                      const float* w, float* out, long long rows, int H, float eps, cudaStream_t s);
int dsq_attn_splits(int smax);
size_t dsq_attn_ws_floats(long long rows, int heads, int nsplit);
// counters: int[rows*heads], zero before the first launch (the kernel hands them back zeroed)
void dsq_attn_split(const float* qkv, const float* cos_t, const float* sin_t, void* kc, void* vc, bool kv_f16,
                    const int* row_page, const int* row_pos, float* part, int* counters, float* ctx, long long rows,
                    int heads, int head_dim, int smax, float scale, int nsplit, cudaStream_t s);

// h[i] = silu(g[i]) * u[i]
void swiglu_f32(const float* g, const float* u, float* h, long long n, cudaStream_t stream);

}  // namespace dsocr
