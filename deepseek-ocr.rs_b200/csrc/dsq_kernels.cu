// DSQ dequant-fused GEMV (decode path of run_quantized_matmul, crates/infer-deepseek/src/quantization.rs:164-185):
//   out[r, n] (+)= sum_k x[r, k] * dequant(W)[n, k]
// for Q8_0 / Q4_K / Q6_K ggml blocks (and the exporter's float fallback).  Weights are streamed once with
// 128-bit loads straight from their quantised planes, dequantised in registers and multiplied with f32
// activations (no Q8_1 re-quantisation of the activations as candle's CUDA path does), f32 accumulation,
// warp-shuffle reduction.  One warp = R output features x MT token rows; lanes stride over 16-byte units of K.
#include <cuda_fp16.h>

#include "dsq.h"
#include "kernels.h"

namespace dsocr {

namespace {

constexpr int kWarps = 8;  // per block
constexpr int R = 2;       // output features per warp

struct GemvArgs {
  const uint8_t* a; const uint8_t* b; const uint8_t* c; const uint8_t* d;
  long long N;      // rows per expert matrix
  int K;
  const float* x; long long ldx; int x_row_div;
  const int* row_expert;
  float* out; long long ldo; long long rows;
  int accumulate;
};

__device__ __forceinline__ void load_x16(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <int FMT, int MT>
__global__ void __launch_bounds__(kWarps * 32)
dsq_gemv_kernel(const GemvArgs g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n0 = ((long long)blockIdx.x * kWarps + warp) * R;
  const long long r0 = (long long)blockIdx.y * MT;
  if (n0 >= g.N) return;
  const int K = g.K;
  const long long e = g.row_expert ? g.row_expert[r0] : 0;
  const long long wrow0 = e * g.N + n0;  // first weight row of this warp inside the stacked planes
  const float* xr[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const long long r = r0 + m < g.rows ? r0 + m : g.rows - 1;
    xr[m] = g.x + (r / g.x_row_div) * g.ldx;
  }
  float acc[R][MT];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;
  const int nrow = (int)min((long long)R, g.N - n0);

  if (FMT == 8) {  // Q8_0 planes: qs int8 [rows][K], d f16 [rows][K/32]
    for (int k = lane * 16; k < K; k += 512) {
      float xv[MT][16];
#pragma unroll
      for (int m = 0; m < MT; ++m) load_x16(xr[m] + k, xv[m]);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nrow) break;
        const long long row = wrow0 + r;
        const uint4 q = *reinterpret_cast<const uint4*>(g.a + row * K + k);
        const float d = __half2float(reinterpret_cast<const __half*>(g.b)[row * (K / 32) + (k >> 5)]);
        const int8_t* qb = reinterpret_cast<const int8_t*>(&q);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) s = fmaf((float)qb[i], xv[m][i], s);
          acc[r][m] = fmaf(d, s, acc[r][m]);
        }
      }
    }
  } else if (FMT == 12) {  // Q4_K blocks as on disk (144 B / 256 weights)
    const int units = (K / 256) * 8;
    for (int u = lane; u < units; u += 32) {
      const int sb = u >> 3, j = u & 7, gq = j >> 1, lo = (j & 1) * 16;
      const int k1 = sb * 256 + (2 * gq) * 32 + lo, k2 = k1 + 32;
      float x1[MT][16], x2[MT][16], sx1[MT], sx2[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        load_x16(xr[m] + k1, x1[m]);
        load_x16(xr[m] + k2, x2[m]);
        sx1[m] = 0.f; sx2[m] = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { sx1[m] += x1[m][i]; sx2[m] += x2[m][i]; }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nrow) break;
        const uint8_t* blk = g.a + ((wrow0 + r) * (K / 256) + sb) * 144;
        const uint4 hdr = *reinterpret_cast<const uint4*>(blk);  // d, dmin, scales[12]
        const __half2 dd = *reinterpret_cast<const __half2*>(&hdr.x);
        const float d = __low2float(dd), dmin = __high2float(dd);
        const uint8_t* s = reinterpret_cast<const uint8_t*>(&hdr) + 4;
        int sc1, m1, sc2, m2;  // get_scale_min_k4 for sub-blocks 2g and 2g+1
        {
          const int j1 = 2 * gq, j2 = 2 * gq + 1;
          if (j1 < 4) { sc1 = s[j1] & 63; m1 = s[j1 + 4] & 63; }
          else { sc1 = (s[j1 + 4] & 0xF) | ((s[j1 - 4] >> 6) << 4); m1 = (s[j1 + 4] >> 4) | ((s[j1] >> 6) << 4); }
          if (j2 < 4) { sc2 = s[j2] & 63; m2 = s[j2 + 4] & 63; }
          else { sc2 = (s[j2 + 4] & 0xF) | ((s[j2 - 4] >> 6) << 4); m2 = (s[j2 + 4] >> 4) | ((s[j2] >> 6) << 4); }
        }
        const uint4 q = *reinterpret_cast<const uint4*>(blk + 16 + gq * 32 + lo);
        const uint8_t* qb = reinterpret_cast<const uint8_t*>(&q);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            s1 = fmaf((float)(qb[i] & 0xF), x1[m][i], s1);
            s2 = fmaf((float)(qb[i] >> 4), x2[m][i], s2);
          }
          acc[r][m] += d * ((float)sc1 * s1 + (float)sc2 * s2) - dmin * ((float)m1 * sx1[m] + (float)m2 * sx2[m]);
        }
      }
    }
  } else if (FMT == 14) {  // Q6_K planes: ql [rows][K/2], qh [rows][K/4], sc i8 [rows][K/16], d f16 [rows][K/256]
    const int units = (K / 256) * 8;
    for (int u = lane; u < units; u += 32) {
      const int sb = u >> 3, j = u & 7, half = j >> 2, jj = j & 3, second = jj >> 1, l0 = (jj & 1) * 16;
      const int kb = sb * 256 + half * 128 + l0;
      const int k1 = kb + (second ? 32 : 0), k2 = kb + (second ? 96 : 64);
      float x1[MT][16], x2[MT][16];
#pragma unroll
      for (int m = 0; m < MT; ++m) { load_x16(xr[m] + k1, x1[m]); load_x16(xr[m] + k2, x2[m]); }
      const int is = half * 8 + (jj & 1) + (second ? 2 : 0);
      const int sh = second ? 2 : 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nrow) break;
        const long long row = wrow0 + r;
        const uint4 ql = *reinterpret_cast<const uint4*>(g.a + row * (K / 2) + sb * 128 + half * 64 + second * 32 + l0);
        const uint4 qh = *reinterpret_cast<const uint4*>(g.b + row * (K / 4) + sb * 64 + half * 32 + l0);
        const int8_t* sc = reinterpret_cast<const int8_t*>(g.c) + row * (K / 16) + sb * 16;
        const float d = __half2float(reinterpret_cast<const __half*>(g.d)[row * (K / 256) + sb]);
        const float d1 = d * (float)sc[is], d2 = d * (float)sc[is + 4];
        const uint8_t* lb = reinterpret_cast<const uint8_t*>(&ql);
        const uint8_t* hb = reinterpret_cast<const uint8_t*>(&qh);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int qa = ((lb[i] & 0xF) | (((hb[i] >> sh) & 3) << 4)) - 32;
            const int qb2 = ((lb[i] >> 4) | (((hb[i] >> (sh + 4)) & 3) << 4)) - 32;
            s1 = fmaf((float)qa, x1[m][i], s1);
            s2 = fmaf((float)qb2, x2[m][i], s2);
          }
          acc[r][m] += d1 * s1 + d2 * s2;
        }
      }
    }
  } else {  // f32 fallback plane [rows][K]
    for (int k = lane * 4; k < K; k += 128) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (r >= nrow) break;
        const float4 w = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g.a) + (wrow0 + r) * K + k);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(xr[m] + k);
          acc[r][m] += w.x * xv.x + w.y * xv.y + w.z * xv.z + w.w * xv.w;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float v = acc[r][m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && r < nrow && r0 + m < g.rows) {
        float* o = g.out + (r0 + m) * g.ldo + n0 + r;
        *o = g.accumulate ? *o + v : v;
      }
    }
}

__global__ void swiglu_f32_kernel(const float* __restrict__ gte, const float* __restrict__ up, float* __restrict__ h, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gv = gte[i];
  h[i] = gv / (1.f + __expf(-gv)) * up[i];
}

template <int FMT>
void launch_fmt(const GemvArgs& a, bool pairs, cudaStream_t s) {
  const unsigned gx = (unsigned)((a.N + kWarps * R - 1) / (kWarps * R));
  if (pairs) dsq_gemv_kernel<FMT, 2><<<dim3(gx, (unsigned)((a.rows + 1) / 2)), kWarps * 32, 0, s>>>(a);
  else dsq_gemv_kernel<FMT, 1><<<dim3(gx, (unsigned)a.rows), kWarps * 32, 0, s>>>(a);
}

}  // namespace

void dsq_gemv(const DsqGemvCall& c, cudaStream_t stream) {
  if (c.rows <= 0) return;
  const QuantWeight& w = *c.w;
  if (w.K % 16) throw std::runtime_error("dsq_gemv: K must be a multiple of 16");
  GemvArgs a{};
  a.a = (const uint8_t*)w.a.p; a.b = (const uint8_t*)w.b.p; a.c = (const uint8_t*)w.c.p; a.d = (const uint8_t*)w.d.p;
  a.N = w.N; a.K = w.K; a.x = c.x; a.ldx = c.ldx; a.x_row_div = c.x_row_div < 1 ? 1 : c.x_row_div;
  a.row_expert = c.row_expert; a.out = c.out; a.ldo = c.ldo; a.rows = c.rows; a.accumulate = c.accumulate ? 1 : 0;
  // token rows are processed in pairs (weights re-read once per pair) unless every row selects its own expert
  const bool pairs = !c.row_expert && c.x_row_div == 1 && c.rows > 1;
  switch (w.fmt) {
    case DsqDType::Q8_0: launch_fmt<8>(a, pairs, stream); break;
    case DsqDType::Q4K: launch_fmt<12>(a, pairs, stream); break;
    case DsqDType::Q6K: launch_fmt<14>(a, pairs, stream); break;
    default: launch_fmt<0>(a, pairs, stream); break;
  }
  launch_check(c.tag);
}

void swiglu_f32(const float* g, const float* u, float* h, long long n, cudaStream_t stream) {
  swiglu_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(g, u, h, n);
  launch_check("swiglu_f32");
}

}  // namespace dsocr
