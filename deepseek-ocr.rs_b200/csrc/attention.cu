// Host launchers for the vision attention kernel (attention_tc.cuh).
#include <cstdlib>

#include "attention_tc.cuh"
#include "kernels.h"
#include "linear_tc.cuh"
#include "tmap.h"

namespace dsocr {

namespace {

template <typename T, int GW, int RB, int KV, int MINB, int POLY>
void launch_poly(const VAttnCall& c, cudaStream_t stream) {
  using C = vattn::Cfg<KV>;
  auto kern = vattn::vattn_kernel<T, GW, RB, KV, MINB, POLY>;
  static PerDeviceOnce once;  // per instantiation
  once.run([&] {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes),
               "vattn: set max dynamic smem");
  });
  const long long ld = 3LL * c.H * vattn::D;
  CUtensorMap tq = tmap::make_2d_16bit(c.qkv, c.rows, ld, ld, vattn::BQ, vattn::D);
  CUtensorMap tkv = tmap::make_2d_16bit(c.qkv, c.rows, ld, ld, KV, vattn::D);
  vattn::Params p{};
  p.S = c.S; p.H = c.H; p.nblk = (c.S + KV - 1) / KV;
  p.scale_log2 = c.scale * 1.4426950408889634f;
  p.Z = c.Z; p.zw = c.zw; p.zhalf = c.zhalf; p.out = c.out;
  dim3 grid((c.S + vattn::BQ - 1) / vattn::BQ, c.B * c.H);
  kern<<<grid, vattn::kThreads, C::kSmemBytes, stream>>>(tq, tkv, p);
  launch_check(c.tag ? c.tag : "vision_attention");
}

// Exponentials per 16 that run as an FMA-pipe polynomial instead of MUFU.EX2 (DSOCR_VATTN_POLY = 0 or 3 is an A/B
// switch for the profiles; the default is what measured fastest on the 64-grid global attention).
int vattn_poly() {
  static const int v = [] {
    const char* e = getenv("DSOCR_VATTN_POLY");
    return e ? atoi(e) : 0;
  }();
  return v;
}

template <typename T, int GW, int RB, int KV, int MINB>
void launch(const VAttnCall& c, cudaStream_t stream) {
  if (vattn_poly() == 3) launch_poly<T, GW, RB, KV, MINB, 3>(c, stream);
  else launch_poly<T, GW, RB, KV, MINB, 0>(c, stream);
}

template <typename T>
void dispatch(const VAttnCall& c, cudaStream_t stream) {
  // DSOCR_VATTN_KV128=1: the single-score-stage configurations (128-key blocks, no QK / softmax overlap) - A/B switch
  static const bool kv128 = getenv("DSOCR_VATTN_KV128") != nullptr;
  if (kv128) {
    switch (c.grid) {
      case 0: launch<T, 0, 0, 128, 2>(c, stream); return;
      case 64: launch<T, 64, 2, 128, 2>(c, stream); return;
      case 32: launch<T, 32, 4, 128, 2>(c, stream); return;
      case 16: launch<T, 16, 8, 128, 2>(c, stream); return;
      default: break;
    }
  }
  switch (c.grid) {
    case 0: launch<T, 0, 0, 64, 2>(c, stream); break;
    case 64: launch<T, 64, 1, 64, 2>(c, stream); break;
    case 40: launch<T, 40, 2, 80, 2>(c, stream); break;
    case 32: launch<T, 32, 2, 64, 2>(c, stream); break;
    case 16: launch<T, 16, 4, 64, 2>(c, stream); break;
    case 14: launch<T, 14, 8, 112, 2>(c, stream); break;
    default: throw std::runtime_error("vision attention: unsupported token grid " + std::to_string(c.grid));
  }
}

}  // namespace

bool vision_attention_supported(int grid) {
  return grid == 0 || grid == 64 || grid == 40 || grid == 32 || grid == 16 || grid == 14;
}

void vision_attention(const VAttnCall& c, DType dt, cudaStream_t stream) {
  if (c.grid > 0 && c.grid * c.grid != c.S) throw std::runtime_error("vision attention: S must equal grid^2");
  if (dt == DType::BF16) dispatch<__nv_bfloat16>(c, stream);
  else dispatch<__half>(c, stream);
}

void vision_relpos_products(const void* qkv, long long rows, int H, const void* table, int zhalf, float* Z, DType dt,
                            int num_sms, cudaStream_t stream) {
  LinearCall c;
  c.w0 = table; c.N = 2 * zhalf; c.K = 64;
  c.x = qkv; c.x_rows = rows; c.ldx = 3LL * H * 64; c.nbatch = H; c.x_batch_stride = 64;
  c.M = (int)rows;
  c.out = Z; c.ldo = (long long)H * 2 * zhalf; c.out_batch_stride = 2 * zhalf;
  c.out_mode = lin::OUT_F32;
  c.tag = "sam_relpos_products";
  linear(c, dt, num_sms, stream);
}

}  // namespace dsocr
