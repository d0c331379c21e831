// Bandwidth-bound kernels of the DeepSeek-V2 MoE decoder path: embedding gather + image-token injection,
// RMSNorm (hi/lo split outputs), RoPE + KV-cache append, causal prefill attention, flash-decode over the KV
// cache, MoE router/top-k, dispatch/combine, n-gram ban + first-index argmax.  f32 math throughout (the
// reference keeps the whole decoder in f32); 128-bit accesses, warp-shuffle reductions.
#include "kernels.h"

#include <algorithm>
#include "ptx.cuh"
#include "attention_prefill_tc.cuh"

#include <cfloat>
#include <cstdlib>

namespace dsocr {

namespace {

template <typename T>
__device__ __forceinline__ uint32_t pack2f(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2f<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2f<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 4 floats -> 4 x 16-bit hi (uint2) and 4 x 16-bit lo (uint2) with lo = r16(x - hi)
template <typename T>
__device__ __forceinline__ void split4(const float* f, uint2& hi, uint2& lo) {
  float r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = f[i] - Elem<T>::to(Elem<T>::from(f[i]));
  hi.x = pack2f<T>(f[0], f[1]); hi.y = pack2f<T>(f[2], f[3]);
  lo.x = pack2f<T>(r[0], r[1]); lo.y = pack2f<T>(r[2], r[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------------
// embed_tokens + inject_image_tokens (transformer/model.rs:116-127, model/mod.rs:1760-1857):
// src[r] >= 0 -> embedding row src[r];  src[r] < 0 -> image row (-src[r] - 1).
template <typename T>
__global__ void embed_gather_kernel(const int* __restrict__ src, const T* __restrict__ table,
                                    const float* __restrict__ img_rows, float* __restrict__ out, long long rows,
                                    int H) {
  const int h4 = H / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * h4) return;
  const long long r = idx / h4;
  const int c = idx % h4;
  const int s = src[r];
  float4 v;
  if (s >= 0) {
    const uint2 raw = reinterpret_cast<const uint2*>(table + (long long)s * H)[c];
    const T* e = reinterpret_cast<const T*>(&raw);
    v = make_float4(Elem<T>::to(e[0]), Elem<T>::to(e[1]), Elem<T>::to(e[2]), Elem<T>::to(e[3]));
  } else {
    v = reinterpret_cast<const float4*>(img_rows + (long long)(-s - 1) * H)[c];
  }
  reinterpret_cast<float4*>(out)[idx] = v;
}

// rms_norm (block.rs:24-29): x * rsqrt(mean(x^2) + eps) * w in f32.  One warp per row.  Outputs: 16-bit hi/lo
// split [2][lo_off rows apart] for the tensor-core GEMMs and optionally the f32 row (router input).
// `row_idx` (optional) selects source rows (last-row-only final norm).
// When `partials` is given the row first absorbs the split-K partial sums of the preceding projection
// (x[src] += sum_s partials[s][src], fixed order, written back) - the residual add fused into the norm.
template <typename T>
__global__ void rmsnorm_kernel(float* __restrict__ x, const float* __restrict__ w, T* __restrict__ out16,
                               long long lo_off_elems, float* __restrict__ out32, const int* __restrict__ row_idx,
                               long long rows, int H, float eps, const float* __restrict__ partials, int n_splits,
                               long long split_stride) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long src = row_idx ? row_idx[row] : row;
  float4* xr = reinterpret_cast<float4*>(x + src * H);
  const int n4 = H / 4;
  float ss = 0.f;
  for (int i = lane; i < n4; i += 32) {
    float4 v = xr[i];
    if (partials) {
      for (int sidx = 0; sidx < n_splits; ++sidx) {
        const float4 pv = reinterpret_cast<const float4*>(partials + sidx * split_stride + src * H)[i];
        v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
      }
      xr[i] = v;
    }
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (partials) __syncwarp();
  const float inv = rsqrtf(warp_sum(ss) / (float)H + eps);
  for (int i = lane; i < n4; i += 32) {
    const float4 v = xr[i];
    const float4 ww = reinterpret_cast<const float4*>(w)[i];
    float o[4] = {v.x * inv * ww.x, v.y * inv * ww.y, v.z * inv * ww.z, v.w * inv * ww.w};
    uint2 hi, lo;
    split4<T>(o, hi, lo);
    reinterpret_cast<uint2*>(out16 + row * H)[i] = hi;
    reinterpret_cast<uint2*>(out16 + lo_off_elems + row * H)[i] = lo;
    if (out32) reinterpret_cast<float4*>(out32 + row * H)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Decode-sized variant: one 128-thread block per row, all partial loads of a thread in flight at once.
template <typename T, int MAXS>
__global__ void __launch_bounds__(128)
rmsnorm_row_kernel(float* __restrict__ x, const float* __restrict__ w, T* __restrict__ out16, long long lo_off_elems,
                   float* __restrict__ out32, const int* __restrict__ row_idx, int H, float eps,
                   const float* __restrict__ partials, int n_splits, long long split_stride) {
  const long long row = blockIdx.x;
  const long long src = row_idx ? row_idx[row] : row;
  float4* xr = reinterpret_cast<float4*>(x + src * H);
  const int n4 = H / 4;
  float4 v[3];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int i = threadIdx.x + it * 128;
    v[it] = make_float4(0, 0, 0, 0);
    if (i < n4) {
      float4 a = xr[i];
      if (partials) {
        float4 pv[MAXS];
#pragma unroll
        for (int sidx = 0; sidx < MAXS; ++sidx)
          pv[sidx] = sidx < n_splits ? reinterpret_cast<const float4*>(partials + sidx * split_stride + src * H)[i]
                                     : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int sidx = 0; sidx < MAXS; ++sidx) { a.x += pv[sidx].x; a.y += pv[sidx].y; a.z += pv[sidx].z; a.w += pv[sidx].w; }
        xr[i] = a;
      }
      v[it] = a;
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
  }
  __shared__ float red[4];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float inv = rsqrtf((red[0] + red[1] + red[2] + red[3]) / (float)H + eps);
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int i = threadIdx.x + it * 128;
    if (i < n4) {
      const float4 ww = reinterpret_cast<const float4*>(w)[i];
      float o[4] = {v[it].x * inv * ww.x, v[it].y * inv * ww.y, v[it].z * inv * ww.z, v[it].w * inv * ww.w};
      uint2 hi, lo;
      split4<T>(o, hi, lo);
      reinterpret_cast<uint2*>(out16 + row * H)[i] = hi;
      reinterpret_cast<uint2*>(out16 + lo_off_elems + row * H)[i] = lo;
      if (out32) reinterpret_cast<float4*>(out32 + row * H)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// RoPE (block.rs:1403-1471, NeoX rotate-half over all 128 dims, tables rope.rs:172-207) applied to the q and
// k thirds of the fused qkv projection [rows, 3, heads, 128] f32, then K/V appended to the per-page cache
// [page][head][S_max][128] and q written as hi/lo... q stays f32 ([rows, heads, 128]).
// One thread = 4 consecutive dims of the low half (and their partners in the high half).
__device__ __forceinline__ void store4(float* p, int t, float4 v) { reinterpret_cast<float4*>(p)[t] = v; }
__device__ __forceinline__ void store4(__half* p, int t, float4 v) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  reinterpret_cast<uint2*>(p)[t] = u;
}
template <typename TKV>
__global__ void rope_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ cos_t,
                               const float* __restrict__ sin_t, const int* __restrict__ row_page,
                               const int* __restrict__ row_pos, float* __restrict__ q_out, TKV* __restrict__ kc,
                               TKV* __restrict__ vc, long long rows, int heads, int smax, int n_splits,
                               long long split_stride) {
  constexpr int D = 128;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * heads * 16;  // 16 threads x 4 dims = low half (64)
  if (idx >= total) return;
  const int t = idx % 16;
  const int hd = (idx / 16) % heads;
  const long long r = idx / (16LL * heads);
  const int pos = row_pos[r];
  const int page = row_page[r];
  const float4 c = reinterpret_cast<const float4*>(cos_t + (long long)pos * 64)[t];
  const float4 s = reinterpret_cast<const float4*>(sin_t + (long long)pos * 64)[t];
  const float* base = qkv + r * 3 * heads * D;
  const long long cache_off = (((long long)page * heads + hd) * smax + pos) * D;
  auto load4 = [&](const float* p) {  // sum of the split-K partials of the qkv projection (fixed order)
    float4 a = reinterpret_cast<const float4*>(p)[t];
    for (int sidx = 1; sidx < n_splits; ++sidx) {
      const float4 b = reinterpret_cast<const float4*>(p + sidx * split_stride)[t];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    return a;
  };
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float* src = base + (which * heads + hd) * D;
    const float4 lo = load4(src);
    const float4 hi = load4(src + 64);
    // out = x*cos + rotate_half(x)*sin ; rotate_half = [-hi, lo]; cos/sin halves are duplicated
    float4 olo, ohi;
    olo.x = lo.x * c.x - hi.x * s.x; olo.y = lo.y * c.y - hi.y * s.y; olo.z = lo.z * c.z - hi.z * s.z; olo.w = lo.w * c.w - hi.w * s.w;
    ohi.x = hi.x * c.x + lo.x * s.x; ohi.y = hi.y * c.y + lo.y * s.y; ohi.z = hi.z * c.z + lo.z * s.z; ohi.w = hi.w * c.w + lo.w * s.w;
    if (which == 0) {
      float* dst = q_out + (r * heads + hd) * D;
      store4(dst, t, olo); store4(dst + 64, t, ohi);
    } else {
      store4(kc + cache_off, t, olo); store4(kc + cache_off + 64, t, ohi);
    }
  }
  const float* vsrc = base + (2 * heads + hd) * D;
  store4(vc + cache_off, t, load4(vsrc));
  store4(vc + cache_off + 64, t, load4(vsrc + 64));
}

// ---------------------------------------------------------------------------------------------------
// Attention over the KV cache, f32 (attention_forward, block.rs:446-804).  Used for both prefill (causal:
// query at position p attends keys 0..p) and decode (one query per page).  Grid: (query, head); 4 warps split
// the keys; each warp processes 4 keys per step with 8 lanes per key (16 dims per lane).  Output: hi/lo split
// 16-bit context rows [rows, heads*128] feeding the o_proj GEMM.
__device__ __forceinline__ void load16(const float* p, float* out) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(p)[i];
    out[4 * i] = t.x; out[4 * i + 1] = t.y; out[4 * i + 2] = t.z; out[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void load16(const __half* p, float* out) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint4 t = reinterpret_cast<const uint4*>(p)[i];
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); out[8 * i + 2 * j] = f.x; out[8 * i + 2 * j + 1] = f.y; }
  }
}
template <typename T, typename TKV>
__global__ void __launch_bounds__(128)
kv_attention_kernel(const float* __restrict__ q, const TKV* __restrict__ kc, const TKV* __restrict__ vc,
                    const int* __restrict__ row_page, const int* __restrict__ row_pos, T* __restrict__ ctx,
                    long long lo_off_elems, float* __restrict__ ctx32, int heads, int smax, float scale) {
  constexpr int D = 128;
  const long long r = blockIdx.x;
  const int hd = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 3, sub = lane & 7;  // 4 key groups x 8 lanes; lane covers dims [sub*16, sub*16+16)
  const int nkeys = row_pos[r] + 1;
  const int page = row_page[r];
  const TKV* kbase = kc + ((long long)page * heads + hd) * smax * D;
  const TKV* vbase = vc + ((long long)page * heads + hd) * smax * D;
  float qv[16];
  {
    const float4* qp = reinterpret_cast<const float4*>(q + (r * heads + hd) * D + sub * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = qp[i];
      qv[4 * i] = t.x * scale; qv[4 * i + 1] = t.y * scale; qv[4 * i + 2] = t.z * scale; qv[4 * i + 3] = t.w * scale;
    }
  }
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int k0 = warp * 4; k0 < nkeys; k0 += 16) {
    const int k = k0 + grp;
    const bool ok = k < nkeys;
    float s = 0.f;
    float vv[16];
    if (ok) {
      float kk[16];
      load16(kbase + (long long)k * D + sub * 16, kk);
      load16(vbase + (long long)k * D + sub * 16, vv);
#pragma unroll
      for (int i = 0; i < 16; ++i) s += qv[i] * kk[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok) {
      const float mn = fmaxf(m, s);
      const float a = __expf(m - mn);
      const float p = __expf(s - mn);
      l = l * a + p;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + p * vv[i];
      m = mn;
    }
  }
  // merge the 4 key groups of the warp, then the 4 warps
  __shared__ float sm_m[16], sm_l[16], sm_acc[16][D];
  const int slot = warp * 4 + grp;
  if (sub == 0) { sm_m[slot] = m; sm_l[slot] = l; }
#pragma unroll
  for (int i = 0; i < 16; ++i) sm_acc[slot][sub * 16 + i] = acc[i];
  __syncthreads();
  const int d = threadIdx.x;  // 128 threads = 128 dims
  float gm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) gm = fmaxf(gm, sm_m[i]);
  float num = 0.f, den = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float f = sm_m[i] == -INFINITY ? 0.f : __expf(sm_m[i] - gm);
    num += f * sm_acc[i][d];
    den += f * sm_l[i];
  }
  const float o = num / den;
  const T hi = Elem<T>::from(o);
  const long long oidx = (r * heads + hd) * D + d;
  if (ctx32) { ctx32[oidx] = o; return; }
  ctx[oidx] = hi;
  ctx[lo_off_elems + oidx] = Elem<T>::from(o - Elem<T>::to(hi));
}

// Prefill, tiled: one block = 64 consecutive query rows of one page x one head.  Q (pre-scaled), a 32-key K tile
// and V tile live in shared memory as f32 (rows padded to 132 floats: conflict-free float4 reads); each thread owns
// a 4x4 register tile of the scores and a 4-row x 16-dim tile of the output, so every shared-memory float4 feeds 8
// FMAs instead of one key row being re-read from L1 by every query row.  Online softmax per row across the 8 lanes
// that share it.  Causal: the row at position p attends keys 0..p; key tiles past the block's last row are skipped.
constexpr int kPfQ = 64, kPfK = 32, kPfLd = 132;
constexpr int kPfSmemBytes = (kPfQ * kPfLd + 2 * kPfK * kPfLd + kPfQ * (kPfK + 1)) * 4;

__device__ __forceinline__ void stage_row_chunk(const float* g, float* s) {  // 8 floats
  const float4 a = reinterpret_cast<const float4*>(g)[0], b = reinterpret_cast<const float4*>(g)[1];
  reinterpret_cast<float4*>(s)[0] = a; reinterpret_cast<float4*>(s)[1] = b;
}
__device__ __forceinline__ void stage_row_chunk(const __half* g, float* s) {  // 8 halves -> 8 floats
  const uint4 t = *reinterpret_cast<const uint4*>(g);
  const __half2* h = reinterpret_cast<const __half2*>(&t);
  const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]), f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
  reinterpret_cast<float4*>(s)[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
  reinterpret_cast<float4*>(s)[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
}

template <typename T, typename TKV>
__global__ void __launch_bounds__(128)
kv_attention_prefill_tiled_kernel(const float* __restrict__ q, const TKV* __restrict__ kc, const TKV* __restrict__ vc,
                                  const int* __restrict__ page_row0, const int* __restrict__ page_len,
                                  T* __restrict__ ctx, long long lo_off_elems, float* __restrict__ ctx32, int heads,
                                  int smax, float scale) {
  constexpr int D = 128;
  extern __shared__ float pf_smem[];
  float* Qs = pf_smem;                      // [64][132]
  float* Ks = Qs + kPfQ * kPfLd;            // [32][132]
  float* Vs = Ks + kPfK * kPfLd;            // [32][132]
  float* Ps = Vs + kPfK * kPfLd;            // [64][33]
  const int page = blockIdx.y, hd = blockIdx.z;
  const int len = page_len[page];
  const int q0 = blockIdx.x * kPfQ;
  if (q0 >= len) return;
  const int nq = min(kPfQ, len - q0);
  const long long r0 = (long long)page_row0[page] + q0;
  const int tid = threadIdx.x;
  const int tq = tid >> 3, tk = tid & 7;
  // Q tile, scaled; rows past the page end are zero
  for (int c = tid; c < kPfQ * (D / 4); c += 128) {
    const int i = c >> 5, d4 = c & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nq) {
      v = reinterpret_cast<const float4*>(q + ((r0 + i) * heads + hd) * D)[d4];
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    reinterpret_cast<float4*>(Qs + i * kPfLd)[d4] = v;
  }
  const TKV* kbase = kc + ((long long)page * heads + hd) * smax * D;
  const TKV* vbase = vc + ((long long)page * heads + hd) * smax * D;
  float m[4], l[4], o[4][16];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m[a] = -INFINITY; l[a] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[a][i] = 0.f;
  }
  const int kend = q0 + nq;  // keys 0 .. kend-1 are visible to at least one row of this block
  for (int k0 = 0; k0 < kend; k0 += kPfK) {
    __syncthreads();  // previous tile fully consumed (also orders the Q stores before the first use)
    for (int c = tid; c < kPfK * (D / 8); c += 128) {  // 32 rows x 16 chunks of 8
      const int i = c >> 4, d8 = c & 15;
      if (k0 + i < kend) {
        stage_row_chunk(kbase + (long long)(k0 + i) * D + d8 * 8, Ks + i * kPfLd + d8 * 8);
        stage_row_chunk(vbase + (long long)(k0 + i) * D + d8 * 8, Vs + i * kPfLd + d8 * 8);
      } else {
        // cache rows past the block's last position were never written by this call: stale bytes there may decode to
        // Inf / NaN, and a masked probability of 0 times NaN would still poison the output row
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(Ks + i * kPfLd + d8 * 8)[0] = z; reinterpret_cast<float4*>(Ks + i * kPfLd + d8 * 8)[1] = z;
        reinterpret_cast<float4*>(Vs + i * kPfLd + d8 * 8)[0] = z; reinterpret_cast<float4*>(Vs + i * kPfLd + d8 * 8)[1] = z;
      }
    }
    __syncthreads();
    float sc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) sc[a][b] = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < D / 4; ++d4) {
      float4 qa[4], kb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qa[a] = reinterpret_cast<const float4*>(Qs + (tq + 16 * a) * kPfLd)[d4];
#pragma unroll
      for (int b = 0; b < 4; ++b) kb[b] = reinterpret_cast<const float4*>(Ks + (tk + 8 * b) * kPfLd)[d4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          sc[a][b] += qa[a].x * kb[b].x + qa[a].y * kb[b].y + qa[a].z * kb[b].z + qa[a].w * kb[b].w;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int qpos = q0 + tq + 16 * a;  // position of this query row
      float mx = -INFINITY;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (k0 + tk + 8 * b > qpos) sc[a][b] = -INFINITY;
        mx = fmaxf(mx, sc[a][b]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float mn = fmaxf(m[a], mx);
      const float alpha = (m[a] == -INFINITY) ? 0.f : __expf(m[a] - mn);
      float ps = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float pe = (sc[a][b] == -INFINITY) ? 0.f : __expf(sc[a][b] - mn);
        Ps[(tq + 16 * a) * (kPfK + 1) + tk + 8 * b] = pe;
        ps += pe;
      }
      l[a] = l[a] * alpha + ps;  // per-lane partial sum; the 8 lanes of a row are added at the end
      m[a] = mn;
#pragma unroll
      for (int i = 0; i < 16; ++i) o[a][i] *= alpha;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kPfK; ++kk) {
      float pa[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pa[a] = Ps[(tq + 16 * a) * (kPfK + 1) + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = reinterpret_cast<const float4*>(Vs + kk * kPfLd + tk * 4 + 32 * j)[0];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          o[a][4 * j] += pa[a] * v.x; o[a][4 * j + 1] += pa[a] * v.y;
          o[a][4 * j + 2] += pa[a] * v.z; o[a][4 * j + 3] += pa[a] * v.w;
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float lt = l[a];
    lt += __shfl_xor_sync(0xffffffffu, lt, 1);
    lt += __shfl_xor_sync(0xffffffffu, lt, 2);
    lt += __shfl_xor_sync(0xffffffffu, lt, 4);
    const int i = tq + 16 * a;
    if (i >= nq) continue;
    const float inv = 1.f / lt;
    const long long obase = ((r0 + i) * heads + hd) * D;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ov[4] = {o[a][4 * j] * inv, o[a][4 * j + 1] * inv, o[a][4 * j + 2] * inv, o[a][4 * j + 3] * inv};
      const long long oidx = obase + tk * 4 + 32 * j;
      if (ctx32) { *reinterpret_cast<float4*>(ctx32 + oidx) = make_float4(ov[0], ov[1], ov[2], ov[3]); continue; }
      uint2 hi, lo;
      split4<T>(ov, hi, lo);
      *reinterpret_cast<uint2*>(ctx + oidx) = hi;
      *reinterpret_cast<uint2*>(ctx + lo_off_elems + oidx) = lo;
    }
  }
}

// first row and length of every page of a prefill batch (rows are ordered by page, then position)
__global__ void page_spans_kernel(const int* __restrict__ row_page, const int* __restrict__ row_pos, long long rows,
                                  int* __restrict__ page_row0, int* __restrict__ page_len) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int p = row_page[r];
  if (row_pos[r] == 0) page_row0[p] = (int)r;
  if (r + 1 == rows || row_page[r + 1] != p) page_len[p] = row_pos[r] + 1;
}

// Prefill variant: one warp per query row (4 rows per block share the K/V rows they read through L1), the
// 4 key groups of the warp are merged with shuffles.  Causal: row at position p attends keys 0..p.
template <typename T, typename TKV>
__global__ void __launch_bounds__(128)
kv_attention_prefill_kernel(const float* __restrict__ q, const TKV* __restrict__ kc, const TKV* __restrict__ vc,
                            const int* __restrict__ row_page, const int* __restrict__ row_pos, T* __restrict__ ctx,
                            long long lo_off_elems, float* __restrict__ ctx32, long long rows, int heads, int smax,
                            float scale) {
  constexpr int D = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 4 + warp;
  if (r >= rows) return;
  const int hd = blockIdx.y;
  const int grp = lane >> 3, sub = lane & 7;
  const int nkeys = row_pos[r] + 1;
  const int page = row_page[r];
  const TKV* kbase = kc + ((long long)page * heads + hd) * smax * D;
  const TKV* vbase = vc + ((long long)page * heads + hd) * smax * D;
  float qv[16];
  load16(q + (r * heads + hd) * D + sub * 16, qv);
#pragma unroll
  for (int i = 0; i < 16; ++i) qv[i] *= scale;
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int k0 = 0; k0 < nkeys; k0 += 4) {
    const int k = k0 + grp;
    const bool ok = k < nkeys;
    float s = 0.f;
    float vv[16];
    if (ok) {
      float kk[16];
      load16(kbase + (long long)k * D + sub * 16, kk);
      load16(vbase + (long long)k * D + sub * 16, vv);
#pragma unroll
      for (int i = 0; i < 16; ++i) s += qv[i] * kk[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok) {
      const float mn = fmaxf(m, s);
      const float a = __expf(m - mn);
      const float pe = __expf(s - mn);
      l = l * a + pe;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + pe * vv[i];
      m = mn;
    }
  }
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {  // merge the 4 key groups (lanes with equal `sub`)
    const float mo = __shfl_xor_sync(0xffffffffu, m, off);
    const float lo = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, mo);
    const float a = m == -INFINITY ? 0.f : __expf(m - mn);
    const float b = mo == -INFINITY ? 0.f : __expf(mo - mn);
    l = l * a + lo * b;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float ao = __shfl_xor_sync(0xffffffffu, acc[i], off);
      acc[i] = acc[i] * a + ao * b;
    }
    m = mn;
  }
  if (grp == 0) {
    const float inv = 1.f / l;
    const long long oidx = (r * heads + hd) * D + sub * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float o = acc[i] * inv;
      if (ctx32) { ctx32[oidx + i] = o; continue; }
      const T hi = Elem<T>::from(o);
      ctx[oidx + i] = hi;
      ctx[lo_off_elems + oidx + i] = Elem<T>::from(o - Elem<T>::to(hi));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Decode-step fusions (one query row per page; rows <= 256).  They remove ~56 of the ~175 launches of a step.

// (1) RoPE + KV append + attention: the block of (row, head) reduces the split-K partials of the qkv
// projection for its own head, rotates q/k, appends k/v to the cache and then attends over the cache.
template <typename T, typename TKV>
__global__ void __launch_bounds__(128)
rope_attn_decode_kernel(const float* __restrict__ qkv, int n_splits, long long split_stride,
                        const float* __restrict__ cos_t, const float* __restrict__ sin_t, TKV* __restrict__ kc,
                        TKV* __restrict__ vc, const int* __restrict__ row_page, const int* __restrict__ row_pos,
                        T* __restrict__ ctx, long long lo_off_elems, int heads, int smax, float scale) {
  constexpr int D = 128;
  const long long r = blockIdx.x;
  const int hd = blockIdx.y;
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int pos = row_pos[r];
  const int page = row_page[r];
  const int nkeys = pos + 1;
  TKV* kbase = kc + ((long long)page * heads + hd) * smax * D;
  TKV* vbase = vc + ((long long)page * heads + hd) * smax * D;
  __shared__ float q_s[D];
  constexpr int NS = 16;  // 4 warps x 4 key groups (8 warps measured slower: the smem merge outweighs the extra loads in flight)
  __shared__ float sm_m[NS], sm_l[NS], sm_acc[NS][D];
  auto part_sum = [&](const float* p) {
    float a = p[0];
    for (int sidx = 1; sidx < n_splits; ++sidx) a += p[sidx * split_stride];
    return a;
  };
  const float* base = qkv + r * 3 * heads * D;
  if (t < 64) {
    const float c = cos_t[(long long)pos * 64 + t], sn = sin_t[(long long)pos * 64 + t];
    const float qlo = part_sum(base + hd * D + t), qhi = part_sum(base + hd * D + 64 + t);
    const float klo = part_sum(base + (heads + hd) * D + t), khi = part_sum(base + (heads + hd) * D + 64 + t);
    q_s[t] = (qlo * c - qhi * sn) * scale;
    q_s[t + 64] = (qhi * c + qlo * sn) * scale;
    kbase[(long long)pos * D + t] = (TKV)(klo * c - khi * sn);
    kbase[(long long)pos * D + 64 + t] = (TKV)(khi * c + klo * sn);
  } else if (t < 128) {
    const int d = t - 64;
    vbase[(long long)pos * D + d] = (TKV)part_sum(base + (2 * heads + hd) * D + d);
    vbase[(long long)pos * D + 64 + d] = (TKV)part_sum(base + (2 * heads + hd) * D + 64 + d);
  }
  __syncthreads();  // q in smem, new K/V row visible to the whole block
  float qv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) qv[i] = q_s[sub * 16 + i];
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int k0 = warp * 4; k0 < nkeys; k0 += NS) {
    const int k = k0 + grp;
    const bool ok = k < nkeys;
    float s = 0.f;
    float vv[16];
    if (ok) {
      float kk[16];
      load16(kbase + (long long)k * D + sub * 16, kk);
      load16(vbase + (long long)k * D + sub * 16, vv);
#pragma unroll
      for (int i = 0; i < 16; ++i) s += qv[i] * kk[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok) {
      const float mn = fmaxf(m, s);
      const float a = __expf(m - mn);
      const float pe = __expf(s - mn);
      l = l * a + pe;
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + pe * vv[i];
      m = mn;
    }
  }
  const int slot = warp * 4 + grp;
  if (sub == 0) { sm_m[slot] = m; sm_l[slot] = l; }
#pragma unroll
  for (int i = 0; i < 16; ++i) sm_acc[slot][sub * 16 + i] = acc[i];
  __syncthreads();
  float gm = -INFINITY;
#pragma unroll
  for (int i = 0; i < NS; ++i) gm = fmaxf(gm, sm_m[i]);
  float num = 0.f, den = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const float f = sm_m[i] == -INFINITY ? 0.f : __expf(sm_m[i] - gm);
    num += f * sm_acc[i][t];
    den += f * sm_l[i];
  }
  const float o = num / den;
  const T hi = Elem<T>::from(o);
  const long long oidx = (r * heads + hd) * D + t;
  ctx[oidx] = hi;
  ctx[lo_off_elems + oidx] = Elem<T>::from(o - Elem<T>::to(hi));
}

// (1b) Same fusion, but the cached K/V rows of the (page, head) stream through shared memory with 1-D bulk
// copies (cp.async.bulk, 16 KB per stage, double-buffered per block) instead of per-lane register loads:
// the bytes in flight no longer depend on occupancy / registers.  The new token's k/v stay in shared memory
// and are folded in as the last key (the cache rows written by this block are not re-read).
template <typename T, typename TKV>
__global__ void __launch_bounds__(128)
rope_attn_decode_bulk_kernel(const float* __restrict__ qkv, int n_splits, long long split_stride,
                             const float* __restrict__ cos_t, const float* __restrict__ sin_t, TKV* __restrict__ kc,
                             TKV* __restrict__ vc, const int* __restrict__ row_page, const int* __restrict__ row_pos,
                             T* __restrict__ ctx, long long lo_off_elems, int heads, int smax, float scale) {
  constexpr int D = 128;
  constexpr int KT = 8192 / (D * (int)sizeof(TKV));  // keys per stage tile: 32 (f16) / 16 (f32) -> 8 KB per K or V tile
  constexpr int NST = 2;  // 2 stages x 16 KB: 5 blocks per SM -> ~70 KB of K/V in flight per SM, one wave for 64 pages x 10 heads
  extern __shared__ __align__(128) uint8_t dyn[];
  TKV* tiles = reinterpret_cast<TKV*>(dyn);  // [NST][2][KT][D]
  __shared__ uint64_t full[NST];
  __shared__ float q_s[D], knew[D], vnew[D];
  __shared__ float sm_m[16], sm_l[16], sm_acc[16][D];
  const long long r = blockIdx.x;
  const int hd = blockIdx.y;
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int grp = lane >> 3, sub = lane & 7;
  const int pos = row_pos[r];
  const int page = row_page[r];
  TKV* kbase = kc + ((long long)page * heads + hd) * smax * D;
  TKV* vbase = vc + ((long long)page * heads + hd) * smax * D;
  const int ntiles = (pos + KT - 1) / KT;  // cached keys 0..pos-1
  if (t == 0) {
    for (int s = 0; s < NST; ++s) ptx::mbar_init(&full[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int tile) {
    const int s = tile % NST;
    const int nk = min(KT, pos - tile * KT);
    const uint32_t bytes = (uint32_t)nk * D * sizeof(TKV);
    ptx::mbar_expect_tx(&full[s], 2 * bytes);
    ptx::bulk_load(tiles + ((size_t)s * 2 + 0) * KT * D, kbase + (long long)tile * KT * D, bytes, &full[s]);
    ptx::bulk_load(tiles + ((size_t)s * 2 + 1) * KT * D, vbase + (long long)tile * KT * D, bytes, &full[s]);
  };
  if (t == 0) for (int i = 0; i < NST - 1 && i < ntiles; ++i) issue(i);

  auto part_sum = [&](const float* p) {
    float a = p[0];
    for (int sidx = 1; sidx < n_splits; ++sidx) a += p[sidx * split_stride];
    return a;
  };
  const float* base = qkv + r * 3 * heads * D;
  if (t < 64) {
    const float c = cos_t[(long long)pos * 64 + t], sn = sin_t[(long long)pos * 64 + t];
    const float qlo = part_sum(base + hd * D + t), qhi = part_sum(base + hd * D + 64 + t);
    const float klo = part_sum(base + (heads + hd) * D + t), khi = part_sum(base + (heads + hd) * D + 64 + t);
    q_s[t] = (qlo * c - qhi * sn) * scale;
    q_s[t + 64] = (qhi * c + qlo * sn) * scale;
    const TKV k0 = (TKV)(klo * c - khi * sn), k1 = (TKV)(khi * c + klo * sn);
    kbase[(long long)pos * D + t] = k0;
    kbase[(long long)pos * D + 64 + t] = k1;
    knew[t] = (float)k0; knew[t + 64] = (float)k1;  // what later steps will read back from the cache
  } else {
    const int d = t - 64;
    const TKV v0 = (TKV)part_sum(base + (2 * heads + hd) * D + d), v1 = (TKV)part_sum(base + (2 * heads + hd) * D + 64 + d);
    vbase[(long long)pos * D + d] = v0;
    vbase[(long long)pos * D + 64 + d] = v1;
    vnew[d] = (float)v0; vnew[d + 64] = (float)v1;
  }
  __syncthreads();
  float qv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) qv[i] = q_s[sub * 16 + i];
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  auto fold = [&](float s, const float* vv) {
    const float mn = fmaxf(m, s);
    const float a = __expf(m - mn);
    const float pe = __expf(s - mn);
    l = l * a + pe;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = acc[i] * a + pe * vv[i];
    m = mn;
  };
  for (int tile = 0; tile < ntiles; ++tile) {
    const int s = tile % NST;
    if (t == 0 && tile + NST - 1 < ntiles) issue(tile + NST - 1);  // its stage was released by the sync below
    ptx::mbar_wait(&full[s], (tile / NST) & 1);
    const TKV* kt = tiles + ((size_t)s * 2 + 0) * KT * D;
    const TKV* vt = tiles + ((size_t)s * 2 + 1) * KT * D;
    const int nk = min(KT, pos - tile * KT);
#pragma unroll
    for (int pass = 0; pass < KT / 16; ++pass) {
      const int k = pass * 16 + warp * 4 + grp;
      const bool ok = k < nk;
      float sc = 0.f, vv[16];
      if (ok) {
        float kk[16];
        load16(kt + k * D + sub * 16, kk);
        load16(vt + k * D + sub * 16, vv);
#pragma unroll
        for (int i = 0; i < 16; ++i) sc += qv[i] * kk[i];
      }
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      if (ok) fold(sc, vv);
    }
    __syncthreads();  // everyone is done with stage s before it is refilled
  }
  if (warp == 0 && grp == 0) {  // the new token itself (key index pos)
    float sc = 0.f, vv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { sc += qv[i] * knew[sub * 16 + i]; vv[i] = vnew[sub * 16 + i]; }
    sc += __shfl_xor_sync(0x000000ffu, sc, 1);
    sc += __shfl_xor_sync(0x000000ffu, sc, 2);
    sc += __shfl_xor_sync(0x000000ffu, sc, 4);
    fold(sc, vv);
  }
  const int slot = warp * 4 + grp;
  if (sub == 0) { sm_m[slot] = m; sm_l[slot] = l; }
#pragma unroll
  for (int i = 0; i < 16; ++i) sm_acc[slot][sub * 16 + i] = acc[i];
  __syncthreads();
  float gm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) gm = fmaxf(gm, sm_m[i]);
  float num = 0.f, den = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float f = sm_m[i] == -INFINITY ? 0.f : __expf(sm_m[i] - gm);
    num += f * sm_acc[i][t];
    den += f * sm_l[i];
  }
  const float o = num / den;
  const T hi = Elem<T>::from(o);
  const long long oidx = (r * heads + hd) * D + t;
  ctx[oidx] = hi;
  ctx[lo_off_elems + oidx] = Elem<T>::from(o - Elem<T>::to(hi));
}

// (2) o_proj split-K reduce + residual add + RMSNorm(ln2) + router + top-k + dispatch into fixed-capacity
// expert segments (slot = atomic counter per expert; the grouped GEMM reads the counters).  One block handles R
// consecutive rows: with hundreds of rows per step the router gate (H x E f32, 327 KB) is then read from L2 once per R
// rows instead of once per row and the launch fits one wave of blocks (R = 8: 1024 rows -> 128 blocks; one row per block
// was 7 waves of ~10 us latency chains, a 1024-thread block being alone on its SM).  Per-row arithmetic and summation order do not depend on R.
template <typename T, int E, int R>
__global__ void __launch_bounds__(1024)
post_attn_kernel(float* __restrict__ x, const float* __restrict__ partials, int n_splits, long long split_stride,
                 const float* __restrict__ w, const float* __restrict__ wgt, T* __restrict__ xn16, long long xn_lo_off,
                 int* __restrict__ topk_idx, float* __restrict__ topk_w, int* __restrict__ counts,
                 int* __restrict__ perm_pos, T* __restrict__ xperm, long long xperm_lo_off, int cap, int rows, int H, int topk,
                 int n_shared, float eps, const EpPeers ep, int ep_counts_off) {
  constexpr int KS = 1024 / E;
  constexpr int PER = (E + 31) / 32;
  extern __shared__ float sm[];
  float* xn_s = sm;                 // [R][H]   the rows: first x + partials, then the normalised values
  float* part = sm + R * H;         // [R][1024] router partial sums
  __shared__ float red[R][32];      // sums of squares per (row, warp-sized slice of the row)
  __shared__ int sel_pos[R][16];
  const long long row0 = (long long)blockIdx.x * R;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n4 = H / 4;             // a multiple of 32: a warp never straddles two rows
  const int nslice = n4 / 32;

  for (int idx = t; idx < R * n4; idx += 1024) {
    const int rr = idx / n4, c4 = idx - rr * n4;
    const long long row = row0 + rr;
    float4 v = make_float4(0, 0, 0, 0);
    if (row < rows) {
      v = reinterpret_cast<float4*>(x + row * H)[c4];
      for (int sidx = 0; sidx < n_splits; ++sidx) {
        const float4 pv = reinterpret_cast<const float4*>(partials + sidx * split_stride + row * H)[c4];
        v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
      }
      reinterpret_cast<float4*>(x + row * H)[c4] = v;
    }
    reinterpret_cast<float4*>(xn_s + rr * H)[c4] = v;
    const float ss = warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
    if (lane == 0) red[rr][c4 >> 5] = ss;
  }
  __syncthreads();
  for (int idx = t; idx < R * n4; idx += 1024) {
    const int rr = idx / n4, c4 = idx - rr * n4;
    const long long row = row0 + rr;
    float tot = 0.f;
    for (int i = 0; i < nslice; ++i) tot += red[rr][i];
    const float inv = rsqrtf(tot / (float)H + eps);
    const float4 v = reinterpret_cast<const float4*>(xn_s + rr * H)[c4];
    const float4 ww = reinterpret_cast<const float4*>(w)[c4];
    float o[4] = {v.x * inv * ww.x, v.y * inv * ww.y, v.z * inv * ww.z, v.w * inv * ww.w};
    reinterpret_cast<float4*>(xn_s + rr * H)[c4] = make_float4(o[0], o[1], o[2], o[3]);
    if (row < rows) {
      uint2 hi, lo;
      split4<T>(o, hi, lo);
      reinterpret_cast<uint2*>(xn16 + row * H)[c4] = hi;
      reinterpret_cast<uint2*>(xn16 + xn_lo_off + row * H)[c4] = lo;
    }
  }
  __syncthreads();
  {  // router logits: thread = (k-slice, expert); a gate weight is loaded once and used for the block's R rows
    const int e = t % E, ks = t / E;
    const int kper = H / KS;
    const float* wp = wgt + (long long)(ks * kper) * E + e;
    const float* xp = xn_s + ks * kper;
    float acc[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) acc[rr] = 0.f;
#pragma unroll 16  // 16 independent L2 loads in flight per thread: the gate GEMV is latency bound
    for (int k = 0; k < kper; ++k) {
      const float wv = wp[(long long)k * E];
#pragma unroll
      for (int rr = 0; rr < R; ++rr) acc[rr] = fmaf(xp[rr * H + k], wv, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) part[rr * 1024 + ks * E + e] = acc[rr];
  }
  __syncthreads();
  if (warp < R && row0 + warp < rows) {  // one warp per row: softmax + top-k + slot reservation
    const int rr = warp;
    const long long row = row0 + rr;
    const float* prt = part + rr * 1024;
    float p[PER];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int ee = j * 32 + lane;
      float s = -INFINITY;
      if (ee < E) {
        s = 0.f;
#pragma unroll
        for (int q = 0; q < KS; ++q) s += prt[q * E + ee];
      }
      p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      p[j] = (j * 32 + lane < E) ? expf(p[j] - mx) : 0.f;
      sum += p[j];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int j = 0; j < PER; ++j) p[j] = (j * 32 + lane < E) ? p[j] / sum : -1.f;
    int my_e = 0; float my_w = 0.f;
    for (int k = 0; k < topk; ++k) {
      float bv = -1.f; int bi = 1 << 30;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int ee = j * 32 + lane;
        if (ee < E && (p[j] > bv || (p[j] == bv && ee < bi))) { bv = p[j]; bi = ee; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == k) { my_e = bi; my_w = bv; }
#pragma unroll
      for (int j = 0; j < PER; ++j) if (j * 32 + lane == bi) p[j] = -2.f;
    }
    // one slot reservation per selected expert, all in flight at once (lane k owns choice k)
    if (lane < topk) {
      int pos;
      if (ep.world > 1) {
        // expert parallel: the slot is reserved in the OWNER rank's segment counter (system-scope atomic through the
        // peer mapping); perm_pos = owner << 24 | row inside the owner's segment buffer
        const int owner = my_e / ep.eloc, le = my_e - owner * ep.eloc;
        const int slot = atomicAdd_system(ep.counts[owner] + ep_counts_off + le, 1);
        pos = (owner << 24) | (le * cap + slot);
      } else {
        pos = my_e * cap + atomicAdd(&counts[my_e], 1);
      }
      topk_idx[row * topk + lane] = my_e;
      topk_w[row * topk + lane] = my_w;
      perm_pos[row * topk + lane] = pos;
      sel_pos[rr][lane] = pos;
    }
  }
  __syncthreads();
  // the shared experts ride in the same grouped GEMMs as groups E, E+1, ... (EP: eloc, eloc+1, ... of the token's own
  // rank): every token, slot = its row
  const int g_shared = ep.world > 1 ? ep.eloc : E;
  if (blockIdx.x == 0 && t < n_shared) counts[g_shared + t] = rows;
  const int per_row = (topk + n_shared) * n4;
  for (int idx = t; idx < R * per_row; idx += 1024) {  // copy the normed rows into their expert slots
    const int rr = idx / per_row, rem = idx - rr * per_row;
    const long long row = row0 + rr;
    if (row >= rows) break;
    const int k = rem / n4, c4 = rem % n4;
    const float4 o4 = reinterpret_cast<const float4*>(xn_s + rr * H)[c4];
    const float o[4] = {o4.x, o4.y, o4.z, o4.w};
    uint2 hi, lo;
    split4<T>(o, hi, lo);
    T* base = xperm;
    long long dst;
    if (k >= topk) dst = ((long long)(g_shared + k - topk) * cap + row) * H;
    else if (ep.world > 1) { base = reinterpret_cast<T*>(ep.xperm[sel_pos[rr][k] >> 24]); dst = (long long)(sel_pos[rr][k] & 0xFFFFFF) * H; }
    else dst = (long long)sel_pos[rr][k] * H;
    reinterpret_cast<uint2*>(base + dst)[c4] = hi;
    reinterpret_cast<uint2*>(base + xperm_lo_off + dst)[c4] = lo;
  }
  if (ep.world > 1) __threadfence_system();  // peer stores are performed before the kernel (and the barrier after it) completes
}

// Cross-GPU barrier of an expert-parallel group (one warp): publish this rank's arrival generation in every peer's
// flag row, then wait until every rank has arrived at this rank.  Bounded spin: a protocol error traps instead of
// hanging the box.
__global__ void ep_barrier_kernel(const EpPeers ep, int* __restrict__ gen_ptr) {
  __shared__ int s_gen;
  if (threadIdx.x == 0) { s_gen = *gen_ptr + 1; *gen_ptr = s_gen; }
  __syncthreads();
  const int gen = s_gen;
  __threadfence_system();
  if ((int)threadIdx.x < ep.world) {
    const int peer = threadIdx.x;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ep.flags[peer] + ep.rank), "r"(gen) : "memory");
    const int* mine = ep.flags[ep.rank] + peer;
    unsigned long long spins = 0;
    for (;;) {
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v >= gen) break;
      if (++spins > (1ull << 25)) {  // tens of seconds; a legitimate wait is microseconds to milliseconds
        printf("ep_barrier timeout: rank %d waiting for rank %d at generation %d (has %d)\n", ep.rank, peer, gen, v); __trap(); }
    }
  }
  __syncthreads();
  __threadfence_system();
}

// (3) MoE combine + shared-experts split-K reduce + residual add + the NEXT RMSNorm (next layer's ln1 or the
// final norm), one 128-thread block per row.
template <typename T, int MAXS, int THREADS, int ITERS>
__global__ void __launch_bounds__(THREADS)
combine_norm_kernel(float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ perm_pos,
                    const float* __restrict__ topk_w, int topk, const float* __restrict__ partials, int n_splits,
                    long long split_stride, const float* __restrict__ w_next, T* __restrict__ out16,
                    long long lo_off_elems, int H, float eps, int n_shared, int shared_row0, int cap, const EpPeers ep) {
  const long long row = blockIdx.x;
  const int n4 = H / 4;
  float4 v[ITERS];
  float ss = 0.f;
  float wk[8]; int pk[8];
  const float* yk[8];  // expert-output buffer holding row pk[k]: this rank's, or (expert parallel) the owner rank's over the peer mapping
#pragma unroll
  for (int k = 0; k < 8; ++k) { wk[k] = k < topk ? topk_w[row * topk + k] : 0.f; pk[k] = k < topk ? perm_pos[row * topk + k] : 0; yk[k] = y; }
  if (ep.world > 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < topk) { yk[k] = ep.y[pk[k] >> 24]; pk[k] &= 0xFFFFFF; }
  }
  // shared experts computed as extra groups of the grouped GEMM: rows shared_row0 + s*cap + row, weight 1
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k >= topk && k < topk + n_shared) { wk[k] = 1.f; pk[k] = shared_row0 + (k - topk) * cap + (int)row; }
  topk += n_shared;
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = threadIdx.x + it * THREADS;
    v[it] = make_float4(0, 0, 0, 0);
    if (i < n4) {
      float4 a = reinterpret_cast<float4*>(x + row * H)[i];
      float4 pv[MAXS];
#pragma unroll
      for (int sidx = 0; sidx < MAXS; ++sidx)
        pv[sidx] = sidx < n_splits ? reinterpret_cast<const float4*>(partials + sidx * split_stride + row * H)[i] : make_float4(0, 0, 0, 0);
      float4 yv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) yv[k] = k < topk ? __ldcv(reinterpret_cast<const float4*>(yk[k] + (long long)pk[k] * H) + i) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int sidx = 0; sidx < MAXS; ++sidx) { a.x += pv[sidx].x; a.y += pv[sidx].y; a.z += pv[sidx].z; a.w += pv[sidx].w; }
      float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) { acc.x += wk[k] * yv[k].x; acc.y += wk[k] * yv[k].y; acc.z += wk[k] * yv[k].z; acc.w += wk[k] * yv[k].w; }
      a.x += acc.x; a.y += acc.y; a.z += acc.z; a.w += acc.w;
      reinterpret_cast<float4*>(x + row * H)[i] = a;
      v[it] = a;
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
  }
  __shared__ float red[THREADS / 32];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; ++w) tot += red[w];
  const float inv = rsqrtf(tot / (float)H + eps);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int i = threadIdx.x + it * THREADS;
    if (i < n4) {
      const float4 ww = reinterpret_cast<const float4*>(w_next)[i];
      float o[4] = {v[it].x * inv * ww.x, v[it].y * inv * ww.y, v[it].z * inv * ww.z, v[it].w * inv * ww.w};
      uint2 hi, lo;
      split4<T>(o, hi, lo);
      reinterpret_cast<uint2*>(out16 + row * H)[i] = hi;
      reinterpret_cast<uint2*>(out16 + lo_off_elems + row * H)[i] = lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// MoE router (run_moe, block.rs:1263-1301): logits = x . Wg^T in f32 -> softmax over the experts -> top-k by
// value (ties -> lowest index, the CPU reference's stable descending sort).  One block = TOK tokens; the
// reduction dimension is split over KS thread groups of E threads (thread = expert, coalesced reads of the
// transposed gate weight WgT [H, E]); partial sums meet in shared memory, then one warp per token does
// softmax + top-k with shuffles.  Also counts tokens per expert.
template <int E, int TOK, int THREADS>
__global__ void __launch_bounds__(THREADS)
router_kernel(const float* __restrict__ x, const float* __restrict__ wgt, int* __restrict__ topk_idx,
              float* __restrict__ topk_w, int* __restrict__ counts, long long rows, int H, int topk) {
  constexpr int KS = THREADS / E;      // k-slices
  constexpr int PER = (E + 31) / 32;   // experts per lane in the top-k phase
  extern __shared__ float sm[];        // x tile [TOK][H] | partials [KS][TOK][E]
  float* xs = sm;
  float* part = sm + TOK * H;
  const long long row0 = (long long)blockIdx.x * TOK;
  const int ntok = (int)min((long long)TOK, rows - row0);
  if constexpr (TOK >= 16) {
    // token rows global -> shared with cp.async (zero-filled past the last row): all 16-byte pieces of a thread are in
    // flight at once instead of one register round trip per piece (22 % of the kernel's stall samples were the stores
    // of this loop waiting for their loads)
    for (int i = threadIdx.x; i < TOK * H / 4; i += THREADS) {
      const int t = i / (H / 4);
      const float* src = x + (row0 + (t < ntok ? t : 0)) * H + (size_t)(i % (H / 4)) * 4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xs + (size_t)i * 4);
      const int bytes = t < ntok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  } else {
    for (int i = threadIdx.x; i < TOK * H / 4; i += THREADS) {
      const int t = i / (H / 4);
      float4 v = make_float4(0, 0, 0, 0);
      if (t < ntok) v = reinterpret_cast<const float4*>(x + (row0 + t) * H)[i % (H / 4)];
      reinterpret_cast<float4*>(xs)[i] = v;
    }
  }
  __syncthreads();
  const int e = threadIdx.x % E, ks = threadIdx.x / E;
  const int kper = H / KS;
  float acc[TOK];
#pragma unroll
  for (int t = 0; t < TOK; ++t) acc[t] = 0.f;
  const float* wp = wgt + (long long)(ks * kper) * E + e;
  const float* xp = xs + ks * kper;
  if constexpr (TOK >= 16) {
    // Many tokens per block (prefill).  One expert x all tokens per thread needs one shared-memory load per FMA, and a
    // warp-wide load returns 128 B per clock whether or not it is a broadcast: that, not the FMAs, bounded the kernel
    // (369 us per 31 k rows).  Here a thread owns 4 experts x 4 tokens of its k-slice: per 4 k it loads four float4 of gate
    // weights (coalesced over the expert groups) and four float4 of activations for 64 FMAs.  Per (token, expert) the
    // products are still added in ascending k inside the slice, so the logits are bit-identical to the scalar loop.
    static_assert(TOK == 16 && E % 4 == 0, "router: 4 x 4 register tile");
    constexpr int EG = E / 4;
    const int slot = threadIdx.x % E;
    const int eg = slot % EG, tg = slot / EG;   // expert group (4 experts), token group (4 tokens)
    float a4[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) a4[i][jx] = 0.f;
    const float* wq = wgt + (long long)(ks * kper) * E + 4 * eg;
    const float* xq = xs + (4 * tg) * H + ks * kper;
    // The gate weights come from L2 (~700 cycles) and a thread's FMAs of one 4-k step take ~300: the weights of the next
    // three steps are kept in flight in a register ring of four (the capture of the form without it: 38 % of all stall
    // samples on the first FMA that consumes a freshly loaded weight).  kper / 4 is a multiple of 4 for every supported shape.
    constexpr int PD = 4;
    float4 wr[PD][4];
    const int nstep = kper / 4;
#pragma unroll
    for (int d = 0; d < PD - 1; ++d)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        wr[d][kk] = d < nstep ? *reinterpret_cast<const float4*>(wq + (long long)(4 * d + kk) * E) : make_float4(0, 0, 0, 0);
    for (int s0 = 0; s0 < nstep; s0 += PD) {
#pragma unroll
      for (int d = 0; d < PD; ++d) {
        const int st = s0 + d, k = 4 * st;
        const int pre = st + PD - 1;  // step whose weights are requested now, into the slot used last iteration
        if (pre < nstep) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) wr[(d + PD - 1) % PD][kk] = *reinterpret_cast<const float4*>(wq + (long long)(4 * pre + kk) * E);
        }
        float4 xv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) xv[t] = *reinterpret_cast<const float4*>(xq + t * H + k);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float xk[4] = {xv[t].x, xv[t].y, xv[t].z, xv[t].w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            a4[t][0] = fmaf(xk[kk], wr[d][kk].x, a4[t][0]);
            a4[t][1] = fmaf(xk[kk], wr[d][kk].y, a4[t][1]);
            a4[t][2] = fmaf(xk[kk], wr[d][kk].z, a4[t][2]);
            a4[t][3] = fmaf(xk[kk], wr[d][kk].w, a4[t][3]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      *reinterpret_cast<float4*>(part + (ks * TOK + 4 * tg + t) * E + 4 * eg) = make_float4(a4[t][0], a4[t][1], a4[t][2], a4[t][3]);
  } else {
#pragma unroll 8
    for (int k = 0; k < kper; ++k) {
      const float w = wp[(long long)k * E];
#pragma unroll
      for (int t = 0; t < TOK; ++t) acc[t] = fmaf(xp[t * H + k], w, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < TOK; ++t) part[(ks * TOK + t) * E + e] = acc[t];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < ntok; t += THREADS / 32) {
    float p[PER];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int ee = j * 32 + lane;
      float v = -INFINITY;
      if (ee < E) {
        v = 0.f;
#pragma unroll
        for (int q = 0; q < KS; ++q) v += part[(q * TOK + t) * E + ee];
      }
      p[j] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      p[j] = (j * 32 + lane < E) ? expf(p[j] - mx) : 0.f;
      sum += p[j];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int j = 0; j < PER; ++j) p[j] = (j * 32 + lane < E) ? p[j] / sum : -1.f;
    const long long row = row0 + t;
    for (int k = 0; k < topk; ++k) {
      float bv = -1.f; int bi = 1 << 30;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int ee = j * 32 + lane;
        if (ee < E && (p[j] > bv || (p[j] == bv && ee < bi))) { bv = p[j]; bi = ee; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        topk_idx[row * topk + k] = bi;
        topk_w[row * topk + k] = bv;
        atomicAdd(&counts[bi], 1);
      }
#pragma unroll
      for (int j = 0; j < PER; ++j) if (j * 32 + lane == bi) p[j] = -2.f;
    }
  }
}

// Exclusive scan of the expert counts + tile tables for the two grouped GEMMs (one block, one thread per
// expert).  tiles1: N1 (= moe intermediate) output features per expert; tiles2: N2 (= hidden).
__global__ void moe_plan_kernel(int* __restrict__ counts, int* __restrict__ offsets, int* __restrict__ cursor,
                                LinearTile* __restrict__ tiles1, int* __restrict__ ntiles1,
                                LinearTile* __restrict__ tiles2, int* __restrict__ ntiles2, int E, int bn, int N1,
                                int N2) {
  __shared__ int s_off[257], s_chunk[257];
  const int e = threadIdx.x;
  if (e == 0) {
    int off = 0, ch = 0;
    for (int i = 0; i < E; ++i) {
      s_off[i] = off; s_chunk[i] = ch;
      const int c = counts[i];
      off += c; ch += (c + bn - 1) / bn;
    }
    s_off[E] = off; s_chunk[E] = ch;
    *ntiles1 = ch * (N1 / 128);
    *ntiles2 = ch * (N2 / 128);
  }
  __syncthreads();
  if (e >= E) return;
  const int c = counts[e], off = s_off[e];
  counts[e] = 0;  // ready for the next layer's router (the buffer is zero-initialised once per call)
  offsets[e] = off;
  cursor[e] = 0;
  int ch = s_chunk[e];
  const int w1 = N1 / 128, w2 = N2 / 128;
  for (int r0 = 0; r0 < c; r0 += bn, ++ch) {
    const int rows = min(bn, c - r0);
    for (int wb = 0; wb < w1; ++wb) tiles1[ch * w1 + wb] = LinearTile{e * N1 + wb * 128, off + r0, rows, wb * 128};
    for (int wb = 0; wb < w2; ++wb) tiles2[ch * w2 + wb] = LinearTile{e * N2 + wb * 128, off + r0, rows, wb * 128};
  }
}

// Dispatch: each (token, slot) claims a row in its expert's segment and copies the token's normed hi/lo
// activations there.  One warp per assignment.  (The reference sorts assignments on the HOST:
// block.rs:1303-1313; row order inside an expert does not affect any value.)
template <typename T>
__global__ void moe_dispatch_kernel(const int* __restrict__ topk_idx, const int* __restrict__ offsets,
                                    int* __restrict__ cursor, const T* __restrict__ xn, long long xn_lo_off,
                                    T* __restrict__ xperm, long long xperm_lo_off, int* __restrict__ perm_pos,
                                    long long n_assign, int topk, int H) {
  const long long a = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n_assign) return;
  const int lane = threadIdx.x & 31;
  const int e = topk_idx[a];
  int pos = 0;
  if (lane == 0) pos = offsets[e] + atomicAdd(&cursor[e], 1);
  pos = __shfl_sync(0xffffffffu, pos, 0);
  if (lane == 0) perm_pos[a] = pos;
  const long long tok = a / topk;
  const uint4* shi = reinterpret_cast<const uint4*>(xn + tok * H);
  const uint4* slo = reinterpret_cast<const uint4*>(xn + xn_lo_off + tok * H);
  uint4* dhi = reinterpret_cast<uint4*>(xperm + (long long)pos * H);
  uint4* dlo = reinterpret_cast<uint4*>(xperm + xperm_lo_off + (long long)pos * H);
  for (int i = lane; i < H / 8; i += 32) { dhi[i] = shi[i]; dlo[i] = slo[i]; }
}

// Combine (block.rs:1363-1381): x[token] += sum_k w_k * y[perm_pos[token,k]] in slot order (deterministic).
__global__ void moe_combine_kernel(const float* __restrict__ y, const int* __restrict__ perm_pos,
                                   const float* __restrict__ topk_w, float* __restrict__ x, long long rows, int topk,
                                   int H, const float* __restrict__ partials, int n_splits, long long split_stride) {
  const int h4 = H / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * h4) return;
  const long long r = idx / h4;
  const int c = idx % h4;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int k = 0; k < topk; ++k) {
    const float w = topk_w[r * topk + k];
    const float4 v = reinterpret_cast<const float4*>(y + (long long)perm_pos[r * topk + k] * H)[c];
    acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
  }
  float4 o = reinterpret_cast<float4*>(x)[idx];
  for (int sidx = 0; sidx < n_splits; ++sidx) {  // split-K partials of the shared-experts down projection
    const float4 pv = reinterpret_cast<const float4*>(partials + sidx * split_stride)[idx];
    o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
  }
  o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
  reinterpret_cast<float4*>(x)[idx] = o;
}

// SwiGLU over split-K partials of a fused gate/up projection: g = sum_s part[s][0], u = sum_s part[s][1],
// out = silu(g) * u as hi/lo 16-bit.  part layout: [n_splits][2][rows][N] (dual_stride = rows*N).
template <typename T>
__global__ void swiglu_reduce_kernel(const float* __restrict__ part, int n_splits, long long split_stride,
                                     long long dual_stride, T* __restrict__ out16, long long lo_off_elems,
                                     long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = make_float4(0, 0, 0, 0), u = make_float4(0, 0, 0, 0);
  for (int sidx = 0; sidx < n_splits; ++sidx) {
    const float4 a = reinterpret_cast<const float4*>(part + sidx * split_stride)[i];
    const float4 b = reinterpret_cast<const float4*>(part + sidx * split_stride + dual_stride)[i];
    g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
    u.x += b.x; u.y += b.y; u.z += b.z; u.w += b.w;
  }
  float o[4] = {g.x / (1.f + __expf(-g.x)) * u.x, g.y / (1.f + __expf(-g.y)) * u.y, g.z / (1.f + __expf(-g.z)) * u.z,
                g.w / (1.f + __expf(-g.w)) * u.w};
  uint2 hi, lo;
  split4<T>(o, hi, lo);
  reinterpret_cast<uint2*>(out16)[i] = hi;
  reinterpret_cast<uint2*>(out16 + lo_off_elems)[i] = lo;
}

// ---------------------------------------------------------------------------------------------------
// Token selection (crates/core/src/sampling.rs:34-158, greedy path): repetition penalty over the distinct tokens
// of the context (:120-139), ban of every token that would complete an n-gram already present in the page's
// context (prompt + generated, :141-158), then first-index argmax over the finite values with the reference's
// fall-backs (filtered -> penalised -> raw logits -> 0, :86-95).  kSelChunks blocks per page, each owning one
// slice of the vocabulary: the banned / seen sets are bitmaps of that slice in shared memory, so there is no
// cap on the number of bans (the reference's sets are unbounded HashSets).  The block that finishes last
// reduces the slice winners, appends the new token to the history and freezes pages that emitted EOS (or
// reached their budget).
constexpr int kSelChunks = 16;   // vocabulary slices (blocks) per page
constexpr int kSelThreads = 256;
struct SelBest {
  float v; int i;
  __device__ void init() { v = -INFINITY; i = INT_MAX; }
  __device__ void take(float ov, int oi) { if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; } }
};
__device__ __forceinline__ SelBest sel_block_reduce(SelBest b, float* s_val, int* s_idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    b.take(ov, oi);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = b.v; s_idx[threadIdx.x >> 5] = b.i; }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int w = 1; w < kSelThreads / 32; ++w) b.take(s_val[w], s_idx[w]);
  return b;  // valid in thread 0
}
__global__ void __launch_bounds__(kSelThreads)
select_token_kernel(const float* __restrict__ logits, int V, int* __restrict__ hist, int hist_stride,
                    int* __restrict__ hist_len, int* __restrict__ gen_count, int* __restrict__ finished, int ngram,
                    float penalty, int eos, int max_new, const int* __restrict__ forced, int forced_stride,
                    int* __restrict__ selected_out, int selected_stride, float* __restrict__ part_val,
                    int* __restrict__ part_idx, int* __restrict__ tickets, int per, int words) {
  const int page = blockIdx.y;
  const int chunk = blockIdx.x;
  extern __shared__ unsigned s_bits[];  // [words] banned | [words] seen, bit t - c0
  __shared__ float s_val[kSelThreads / 32];
  __shared__ int s_idx[kSelThreads / 32];
  __shared__ int s_last;
  if (finished[page]) return;
  const int step = gen_count[page];  // tokens accepted so far == index of this selection (graph-replay safe)
  int* h = hist + (long long)page * hist_stride;
  const int L = hist_len[page];
  const int c0 = chunk * per, c1 = min(V, c0 + per);
  unsigned* s_ban = s_bits;
  unsigned* s_seen = s_bits + words;
  for (int i = threadIdx.x; i < 2 * words; i += blockDim.x) s_bits[i] = 0u;
  __syncthreads();
  if (ngram > 1 && L >= ngram - 1) {
    const int pre = ngram - 1;
    for (int i = threadIdx.x; i + ngram <= L; i += blockDim.x) {
      bool eq = true;
      for (int j = 0; j < pre && eq; ++j) eq = h[i + j] == h[L - pre + j];
      if (eq) {
        const int t = h[i + pre];
        if (t >= c0 && t < c1) atomicOr(&s_ban[(t - c0) >> 5], 1u << ((t - c0) & 31));
      }
    }
  }
  const bool use_pen = penalty > 0.f && fabsf(penalty - 1.0f) > FLT_EPSILON;  // sampling.rs:121-123
  if (use_pen) {
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const int t = h[i];
      if (t >= c0 && t < c1) atomicOr(&s_seen[(t - c0) >> 5], 1u << ((t - c0) & 31));
    }
  }
  __syncthreads();
  const float* lg = logits + (long long)page * V;
  SelBest bf, ba, br;  // best of: filtered (penalised + bans) | adjusted (penalised) | raw logits
  bf.init(); ba.init(); br.init();
  for (int i = c0 + threadIdx.x * 4; i < c1; i += kSelThreads * 4) {
    float v4[4];
    if (i + 3 < c1 && ((reinterpret_cast<uintptr_t>(lg + i) & 15) == 0)) {
      const float4 t = *reinterpret_cast<const float4*>(lg + i);
      v4[0] = t.x; v4[1] = t.y; v4[2] = t.z; v4[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v4[k] = (i + k < c1) ? lg[i + k] : NAN;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float raw = v4[k];
      const int t = i + k, bit = t - c0;
      if (t >= c1) continue;
      // ascending scan + strict > keeps the first index per thread; non-finite values are skipped like the
      // reference's is_finite filter (sampling.rs:104-118)
      if (fabsf(raw) <= FLT_MAX && raw > br.v) { br.v = raw; br.i = t; }
      float adj = raw;
      if (use_pen && ((s_seen[bit >> 5] >> (bit & 31)) & 1u)) adj = raw > 0.f ? __fdiv_rn(raw, penalty) : raw * penalty;
      if (!(fabsf(adj) <= FLT_MAX)) continue;
      if (adj > ba.v) { ba.v = adj; ba.i = t; }
      if (!((s_ban[bit >> 5] >> (bit & 31)) & 1u) && adj > bf.v) { bf.v = adj; bf.i = t; }
    }
  }
  bf = sel_block_reduce(bf, s_val, s_idx);
  ba = sel_block_reduce(ba, s_val, s_idx);
  br = sel_block_reduce(br, s_val, s_idx);
  if (threadIdx.x == 0) {
    const int o = (page * kSelChunks + chunk) * 3;
    part_val[o] = bf.v; part_idx[o] = bf.i;
    part_val[o + 1] = ba.v; part_idx[o + 1] = ba.i;
    part_val[o + 2] = br.v; part_idx[o + 2] = br.i;
    __threadfence();
    s_last = atomicAdd(&tickets[page], 1) == kSelChunks - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  // last block of this page: reduce the slice winners and do the bookkeeping
  __threadfence();
  tickets[page] = 0;
  SelBest best[3];
  for (int q = 0; q < 3; ++q) best[q].init();
  for (int c = 0; c < kSelChunks; ++c)
    for (int q = 0; q < 3; ++q) {
      const int o = (page * kSelChunks + c) * 3 + q;
      best[q].take(reinterpret_cast<volatile float*>(part_val)[o], reinterpret_cast<volatile int*>(part_idx)[o]);
    }
  int tok = 0;
  for (int q = 2; q >= 0; --q)
    if (best[q].i != INT_MAX) tok = best[q].i;  // filtered, else penalised, else raw, else 0
  if (selected_out) selected_out[(long long)page * selected_stride + step] = tok;
  if (forced) {
    tok = forced[(long long)page * forced_stride + step];
  } else if (eos >= 0 && tok == eos) {
    finished[page] = 1;  // EOS is not appended (model/mod.rs:2029-2033)
    return;
  }
  h[L] = tok;
  hist_len[page] = L + 1;
  const int g = gen_count[page] + 1;
  gen_count[page] = g;
  if (g >= max_new) finished[page] = 1;
}

// Tokens chosen on the host (sampling path): append / EOS freeze / budget exactly like select_token_kernel's tail.
__global__ void append_tokens_kernel(const int* __restrict__ chosen, int* __restrict__ hist, int hist_stride,
                                     int* __restrict__ hist_len, int* __restrict__ gen_count, int* __restrict__ finished,
                                     int n_pages, int eos, int max_new) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pages || finished[p]) return;
  const int tok = chosen[p];
  if (eos >= 0 && tok == eos) { finished[p] = 1; return; }
  const int L = hist_len[p];
  hist[(long long)p * hist_stride + L] = tok;
  hist_len[p] = L + 1;
  const int g = gen_count[p] + 1;
  gen_count[p] = g;
  if (g >= max_new) finished[p] = 1;
}

// Decode-step bookkeeping: for every page, the row of the next forward is its last history token at
// position hist_len-1 (frozen pages keep recomputing their last position; their results are ignored).
__global__ void decode_rows_kernel(const int* __restrict__ hist, int hist_stride, const int* __restrict__ hist_len,
                                   int* __restrict__ src, int* __restrict__ row_pos, int n_pages) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pages) return;
  const int L = hist_len[p];
  src[p] = hist[(long long)p * hist_stride + L - 1];
  row_pos[p] = L - 1;
}

// diagnostics: stats[0] += number of non-empty (layer, expert) segments of this step, stats[1] += 1
__global__ void moe_active_stat_kernel(const int* __restrict__ counts, int n, unsigned long long* stats) {
  int c = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += counts[i] > 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    stats[0] += (unsigned long long)t;
    stats[1] += 1ull;
  }
}

// row-major [n, k] 16-bit -> 128x64 tiles in 128B-swizzled order (one thread per 16-byte chunk of the output)
__global__ void retile_weights_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n, int k,
                                      long long n_chunks) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_chunks) return;
  const int num_kb = k / 64;
  const long long tile = idx >> 10;          // 1024 chunks of 16 B per tile
  const int within = (int)(idx & 1023);
  const int r = within >> 3, qs = within & 7;  // row inside the tile, swizzled chunk position
  const int q = qs ^ (r & 7);                  // logical 16-byte chunk of that row stored at position qs
  const long long nb = tile / num_kb;
  const int kb = (int)(tile - nb * num_kb);
  const long long row = nb * 128 + r;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (row < n) v = src[(row * k + kb * 64 + q * 8) >> 3];
  dst[idx] = v;
}

__global__ void fill_i32_kernel(int* p, int v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

#define DISPATCH_T(dt, ...)                                              \
  do {                                                                   \
    if ((dt) == DType::BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else { using T = __half; __VA_ARGS__; }                              \
  } while (0)

void embed_gather(const int* src, const void* table, const float* img_rows, float* out, long long rows, int H, DType dt,
                  cudaStream_t s) {
  DISPATCH_T(dt, (embed_gather_kernel<T><<<blocks_for(rows * (H / 4), 256), 256, 0, s>>>(src, (const T*)table, img_rows, out, rows, H)));
  launch_check("embed_gather");
}
void rmsnorm_split(float* x, const float* w, void* out16, long long lo_off_elems, float* out32,
                   const int* row_idx, long long rows, int H, float eps, const float* partials, int n_splits,
                   long long split_stride, DType dt, cudaStream_t s) {
  if (rows <= 256 && H <= 1536 && n_splits <= 16) {
    DISPATCH_T(dt, (rmsnorm_row_kernel<T, 16><<<(unsigned)rows, 128, 0, s>>>(x, w, (T*)out16, lo_off_elems, out32, row_idx, H, eps, partials, n_splits, split_stride)));
  } else {
    DISPATCH_T(dt, (rmsnorm_kernel<T><<<blocks_for(rows, 8), 256, 0, s>>>(x, w, (T*)out16, lo_off_elems, out32, row_idx, rows, H, eps, partials, n_splits, split_stride)));
  }
  launch_check("rmsnorm");
}
void rope_kv(const float* qkv, const float* cos_t, const float* sin_t, const int* row_page, const int* row_pos,
             float* q_out, void* kc, void* vc, bool kv_f16, long long rows, int heads, int smax, int n_splits,
             long long split_stride, cudaStream_t s) {
  const int ns = n_splits < 1 ? 1 : n_splits;
  if (kv_f16) rope_kv_kernel<__half><<<blocks_for(rows * heads * 16, 128), 128, 0, s>>>(qkv, cos_t, sin_t, row_page, row_pos, q_out, (__half*)kc, (__half*)vc, rows, heads, smax, ns, split_stride);
  else rope_kv_kernel<float><<<blocks_for(rows * heads * 16, 128), 128, 0, s>>>(qkv, cos_t, sin_t, row_page, row_pos, q_out, (float*)kc, (float*)vc, rows, heads, smax, ns, split_stride);
  launch_check("rope_kv");
}
void kv_attention(const float* q, const void* kc, const void* vc, bool kv_f16, const int* row_page, const int* row_pos,
                  void* ctx, long long lo_off_elems, float* ctx32, long long rows, int heads, int smax, float scale,
                  DType dt, cudaStream_t s, int* page_spans, int n_pages) {
  if (rows > 256 && page_spans && n_pages > 0) {
    int* row0 = page_spans;
    int* plen = page_spans + n_pages;
    page_spans_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(row_page, row_pos, rows, row0, plen);
    launch_check("page_spans");
    static const bool simt = getenv("DSOCR_PREFILL_SIMT") != nullptr;  // A/B switch: the f32 shared-memory-tile kernel
    if (!simt && !ctx32) {  // prefill on the tensor cores: 128 queries x 64-key blocks, hi/lo split f16 operands
      const long long max_len = std::min<long long>(smax, rows);
      dim3 pgrid((unsigned)((max_len + pattn::BQ - 1) / pattn::BQ), (unsigned)n_pages, (unsigned)heads);
      const float scale_log2 = scale * 1.4426950408889634f;
      if (kv_f16) {
        DISPATCH_T(dt, {
          auto kern = pattn::pattn_kernel<T, __half>;
          static PerDeviceOnce once;  // per instantiation
          once.run([&] { cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pattn::Cfg<1>::kSmemBytes), "prefill attention smem"); });
          kern<<<pgrid, pattn::kThreads, pattn::Cfg<1>::kSmemBytes, s>>>(q, (const __half*)kc, (const __half*)vc, row0, plen, (T*)ctx, lo_off_elems, heads, smax, scale_log2);
        });
      } else {
        DISPATCH_T(dt, {
          auto kern = pattn::pattn_kernel<T, float>;
          static PerDeviceOnce once;  // per instantiation
          once.run([&] { cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pattn::Cfg<2>::kSmemBytes), "prefill attention smem"); });
          kern<<<pgrid, pattn::kThreads, pattn::Cfg<2>::kSmemBytes, s>>>(q, (const float*)kc, (const float*)vc, row0, plen, (T*)ctx, lo_off_elems, heads, smax, scale_log2);
        });
      }
      launch_check("kv_attention");
      return;
    }
    // 64-query x 32-key shared-memory tiles, f32 FMA
    static PerDeviceOnce once;
    once.run([&] {
      cuda_check(cudaFuncSetAttribute(kv_attention_prefill_tiled_kernel<__half, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmemBytes), "prefill attention smem");
      cuda_check(cudaFuncSetAttribute(kv_attention_prefill_tiled_kernel<__half, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmemBytes), "prefill attention smem");
      cuda_check(cudaFuncSetAttribute(kv_attention_prefill_tiled_kernel<__nv_bfloat16, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmemBytes), "prefill attention smem");
      cuda_check(cudaFuncSetAttribute(kv_attention_prefill_tiled_kernel<__nv_bfloat16, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPfSmemBytes), "prefill attention smem");
    });
    const long long max_len = std::min<long long>(smax, rows);  // blocks past a page's end exit at once
    dim3 tgrid((unsigned)((max_len + kPfQ - 1) / kPfQ), (unsigned)n_pages, (unsigned)heads);
    if (kv_f16) {
      DISPATCH_T(dt, (kv_attention_prefill_tiled_kernel<T, __half><<<tgrid, 128, kPfSmemBytes, s>>>(q, (const __half*)kc, (const __half*)vc, row0, plen, (T*)ctx, lo_off_elems, ctx32, heads, smax, scale)));
    } else {
      DISPATCH_T(dt, (kv_attention_prefill_tiled_kernel<T, float><<<tgrid, 128, kPfSmemBytes, s>>>(q, (const float*)kc, (const float*)vc, row0, plen, (T*)ctx, lo_off_elems, ctx32, heads, smax, scale)));
    }
    launch_check("kv_attention");
    return;
  }
  if (rows > 256) {  // prefill without page spans: warp-per-row variant
    dim3 pgrid((unsigned)((rows + 3) / 4), heads);
    if (kv_f16) {
      DISPATCH_T(dt, (kv_attention_prefill_kernel<T, __half><<<pgrid, 128, 0, s>>>(q, (const __half*)kc, (const __half*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, ctx32, rows, heads, smax, scale)));
    } else {
      DISPATCH_T(dt, (kv_attention_prefill_kernel<T, float><<<pgrid, 128, 0, s>>>(q, (const float*)kc, (const float*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, ctx32, rows, heads, smax, scale)));
    }
    launch_check("kv_attention");
    return;
  }
  dim3 grid((unsigned)rows, heads);
  if (kv_f16) {
    DISPATCH_T(dt, (kv_attention_kernel<T, __half><<<grid, 128, 0, s>>>(q, (const __half*)kc, (const __half*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, ctx32, heads, smax, scale)));
  } else {
    DISPATCH_T(dt, (kv_attention_kernel<T, float><<<grid, 128, 0, s>>>(q, (const float*)kc, (const float*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, ctx32, heads, smax, scale)));
  }
  launch_check("kv_attention");
}
template <int E>
static void launch_router(const float* x, const float* wgt, int* topk_idx, float* topk_w, int* counts, long long rows,
                          int H, int topk, cudaStream_t s) {
  if (rows <= 512) {  // decode: one token per block, 1024 threads = 1024/E k-slices for latency hiding
    const size_t smem = (size_t)(1 * H + 1024 * 1) * 4;
    router_kernel<E, 1, 1024><<<(unsigned)rows, 1024, smem, s>>>(x, wgt, topk_idx, topk_w, counts, rows, H, topk);
  } else {            // prefill: 16 tokens per block reuse every gate-weight load 16x (the gate is re-read from L2 by every block:
                      // 8 tokens per block moved 1.2 GB per launch of 30 k rows); same k-slices, same summation order
    const size_t smem = (size_t)(16 * H + 256 * 16) * 4;
    if (smem > 110 * 1024 || (H / (256 / E)) % 16) throw std::runtime_error("router: unsupported hidden size");
    static PerDeviceOnce once;  // per instantiation
    once.run([&] { cuda_check(cudaFuncSetAttribute(router_kernel<E, 16, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024), "router smem"); });
    router_kernel<E, 16, 256><<<(unsigned)((rows + 15) / 16), 256, smem, s>>>(x, wgt, topk_idx, topk_w, counts, rows, H, topk);
  }
}
void moe_router(const float* x, const float* wgt, int* topk_idx, float* topk_w, int* counts, long long rows, int H,
                int E, int topk, cudaStream_t s) {
  if (H % 256) throw std::runtime_error("router: hidden size must be a multiple of 256");
  if (E == 64) launch_router<64>(x, wgt, topk_idx, topk_w, counts, rows, H, topk, s);
  else if (E == 16) launch_router<16>(x, wgt, topk_idx, topk_w, counts, rows, H, topk, s);
  else if (E == 32) launch_router<32>(x, wgt, topk_idx, topk_w, counts, rows, H, topk, s);
  else throw std::runtime_error("router: unsupported expert count " + std::to_string(E));
  launch_check("moe_router");
}
void moe_plan(int* counts, int* offsets, int* cursor, LinearTile* tiles1, int* ntiles1, LinearTile* tiles2,
              int* ntiles2, int E, int bn, int N1, int N2, cudaStream_t s) {
  moe_plan_kernel<<<1, 256, 0, s>>>(counts, offsets, cursor, tiles1, ntiles1, tiles2, ntiles2, E, bn, N1, N2);
  launch_check("moe_plan");
}
void moe_dispatch(const int* topk_idx, const int* offsets, int* cursor, const void* xn, long long xn_lo_off,
                  void* xperm, long long xperm_lo_off, int* perm_pos, long long n_assign, int topk, int H, DType dt,
                  cudaStream_t s) {
  DISPATCH_T(dt, (moe_dispatch_kernel<T><<<blocks_for(n_assign, 8), 256, 0, s>>>(topk_idx, offsets, cursor, (const T*)xn, xn_lo_off, (T*)xperm, xperm_lo_off, perm_pos, n_assign, topk, H)));
  launch_check("moe_dispatch");
}
void moe_combine(const float* y, const int* perm_pos, const float* topk_w, float* x, long long rows, int topk, int H,
                 const float* partials, int n_splits, long long split_stride, cudaStream_t s) {
  moe_combine_kernel<<<blocks_for(rows * (H / 4), 128), 128, 0, s>>>(y, perm_pos, topk_w, x, rows, topk, H, partials, partials ? n_splits : 0, split_stride);
  launch_check("moe_combine");
}
void swiglu_reduce(const float* part, int n_splits, long long split_stride, long long dual_stride, void* out16,
                   long long lo_off_elems, long long n, DType dt, cudaStream_t s) {
  DISPATCH_T(dt, (swiglu_reduce_kernel<T><<<blocks_for(n / 4, 128), 128, 0, s>>>(part, n_splits, split_stride, dual_stride, (T*)out16, lo_off_elems, n / 4)));
  launch_check("swiglu_reduce");
}
void select_token(const float* logits, int V, int* hist, int hist_stride, int* hist_len, int* gen_count, int* finished,
                  int n_pages, int ngram, float penalty, int eos, int max_new, const int* forced, int forced_stride,
                  int* selected_out, int selected_stride, float* scratch, cudaStream_t s) {
  // scratch: [n_pages*16*3] f32 values | [n_pages*16*3] i32 indices | [n_pages] i32 tickets (zero-initialised)
  float* part_val = scratch;
  int* part_idx = reinterpret_cast<int*>(scratch + (size_t)n_pages * kSelChunks * 3);
  int* tickets = part_idx + (size_t)n_pages * kSelChunks * 3;
  const int per = ((V + kSelChunks - 1) / kSelChunks + 3) & ~3;  // slice length, multiple of 4
  const int words = (per + 31) / 32;
  const size_t smem = (size_t)2 * words * 4;
  if (smem > 48 * 1024) throw std::runtime_error("select_token: vocabulary too large for the per-slice bitmaps");
  select_token_kernel<<<dim3(kSelChunks, n_pages), kSelThreads, smem, s>>>(logits, V, hist, hist_stride, hist_len, gen_count,
                                                                         finished, ngram, penalty, eos, max_new, forced,
                                                                         forced_stride, selected_out, selected_stride,
                                                                         part_val, part_idx, tickets, per, words);
  launch_check("select_token");
}
void append_tokens(const int* chosen, int* hist, int hist_stride, int* hist_len, int* gen_count, int* finished, int n_pages,
                   int eos, int max_new, cudaStream_t s) {
  append_tokens_kernel<<<blocks_for(n_pages, 128), 128, 0, s>>>(chosen, hist, hist_stride, hist_len, gen_count, finished,
                                                               n_pages, eos, max_new);
  launch_check("append_tokens");
}
void decode_rows(const int* hist, int hist_stride, const int* hist_len, int* src, int* row_pos, int n_pages,
                 cudaStream_t s) {
  decode_rows_kernel<<<blocks_for(n_pages, 128), 128, 0, s>>>(hist, hist_stride, hist_len, src, row_pos, n_pages);
  launch_check("decode_rows");
}
void moe_active_stat(const int* counts, int n, unsigned long long* stats, cudaStream_t s) {
  moe_active_stat_kernel<<<1, 256, 0, s>>>(counts, n, stats);
  launch_check("moe_active_stat");
}

void retile_weights(const void* src, void* dst, long long n, int k, cudaStream_t s) {
  if (k % 64) throw std::runtime_error("retile_weights: k must be a multiple of 64");
  const long long n_chunks = (long long)(retiled_bytes(n, k) / 16);
  retile_weights_kernel<<<blocks_for(n_chunks, 256), 256, 0, s>>>((const uint4*)src, (uint4*)dst, n, k, n_chunks);
  launch_check("retile_weights");
}

void fill_i32(int* p, int v, long long n, cudaStream_t s) {
  fill_i32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, v, n);
  launch_check("fill_i32");
}

void rope_attn_decode(const float* qkv, int n_splits, long long split_stride, const float* cos_t, const float* sin_t,
                      void* kc, void* vc, bool kv_f16, const int* row_page, const int* row_pos, void* ctx,
                      long long lo_off_elems, long long rows, int heads, int smax, float scale, DType dt, cudaStream_t s) {
  dim3 grid((unsigned)rows, heads);
  const int ns = n_splits < 1 ? 1 : n_splits;
  static const bool use_bulk = getenv("DSOCR_ATTN_NO_BULK") == nullptr;
  constexpr int kDyn = 2 * 2 * 8192;  // NST stages x (K, V) x 8 KB
  if (use_bulk) {
    if (kv_f16) {
      DISPATCH_T(dt, (rope_attn_decode_bulk_kernel<T, __half><<<grid, 128, kDyn, s>>>(qkv, ns, split_stride, cos_t, sin_t, (__half*)kc, (__half*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, heads, smax, scale)));
    } else {
      DISPATCH_T(dt, (rope_attn_decode_bulk_kernel<T, float><<<grid, 128, kDyn, s>>>(qkv, ns, split_stride, cos_t, sin_t, (float*)kc, (float*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, heads, smax, scale)));
    }
  } else if (kv_f16) {
    DISPATCH_T(dt, (rope_attn_decode_kernel<T, __half><<<grid, 128, 0, s>>>(qkv, ns, split_stride, cos_t, sin_t, (__half*)kc, (__half*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, heads, smax, scale)));
  } else {
    DISPATCH_T(dt, (rope_attn_decode_kernel<T, float><<<grid, 128, 0, s>>>(qkv, ns, split_stride, cos_t, sin_t, (float*)kc, (float*)vc, row_page, row_pos, (T*)ctx, lo_off_elems, heads, smax, scale)));
  }
  launch_check("rope_attn_decode");
}
void ep_barrier(const EpPeers& ep, int* gen, cudaStream_t s) {
  ep_barrier_kernel<<<1, 32, 0, s>>>(ep, gen);
  launch_check("ep_barrier");
}
void post_attn(float* x, const float* partials, int n_splits, long long split_stride, const float* w, const float* wgt,
               void* xn16, long long xn_lo_off, int* topk_idx, float* topk_w, int* counts, int* perm_pos, void* xperm,
               long long xperm_lo_off, int cap, long long rows, int H, int E, int topk, int n_shared, float eps, DType dt,
               cudaStream_t s, const EpPeers* epp, int ep_counts_off) {
  if (H % 256 || H > 1536 || topk > 16 || rows > cap) throw std::runtime_error("post_attn: unsupported shape");
  EpPeers ep;
  if (epp) ep = *epp;
  if (ep.world > 1 && (E % ep.world || ep.eloc != E / ep.world || (ep.eloc + n_shared) * (long long)cap >= (1 << 24)))
    throw std::runtime_error("post_attn: unsupported expert-parallel layout");
  // rows per block: the smallest of 1 / 2 / 4 / 8 that puts the step into ONE wave of blocks (a 1024-thread block at
  // ~55 registers per thread is alone on its SM); the gate is then read from L2 once per R rows
  const char* r_str = getenv("DSOCR_POST_ATTN_ROWS");  // A/B switch, also how the tests reach R > 1 with few pages
  const int r_env = r_str ? atoi(r_str) : 0;
  int R = 1;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  while (R < 8 && (rows + R - 1) / R > sms) R *= 2;
  if (r_env == 1 || r_env == 2 || r_env == 4 || r_env == 8) R = r_env;
  const size_t smem = (size_t)R * (H + 1024) * 4;
  const unsigned blocks = (unsigned)((rows + R - 1) / R);
#define POST_ATTN_LAUNCH(EE, RR)                                                                                          \
  do {                                                                                                                    \
    auto kern = post_attn_kernel<T, EE, RR>;                                                                              \
    if (smem > 48 * 1024) {                                                                                               \
      static PerDeviceOnce once;                                                                                          \
      once.run([&] { cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1536 + 1024) * 4), "post_attn smem"); }); \
    }                                                                                                                     \
    kern<<<blocks, 1024, smem, s>>>(x, partials, n_splits, split_stride, w, wgt, (T*)xn16, xn_lo_off, topk_idx, topk_w,    \
                                    counts, perm_pos, (T*)xperm, xperm_lo_off, cap, (int)rows, H, topk, n_shared, eps, ep, \
                                    ep_counts_off);                                                                       \
  } while (0)
#define POST_ATTN_R(EE)                                                                                                   \
  do {                                                                                                                    \
    if (R == 8) POST_ATTN_LAUNCH(EE, 8); else if (R == 4) POST_ATTN_LAUNCH(EE, 4);                                        \
    else if (R == 2) POST_ATTN_LAUNCH(EE, 2); else POST_ATTN_LAUNCH(EE, 1);                                               \
  } while (0)
  DISPATCH_T(dt, {
    if (E == 64) POST_ATTN_R(64);
    else if (E == 32) POST_ATTN_R(32);
    else if (E == 16) POST_ATTN_R(16);
    else throw std::runtime_error("post_attn: unsupported expert count " + std::to_string(E));
  });
#undef POST_ATTN_R
#undef POST_ATTN_LAUNCH
  launch_check("post_attn_norm_router_dispatch");
}
void combine_norm(float* x, const float* y, const int* perm_pos, const float* topk_w, int topk, const float* partials,
                  int n_splits, long long split_stride, const float* w_next, void* out16, long long lo_off_elems,
                  long long rows, int H, float eps, int n_shared, int shared_row0, int cap, DType dt, cudaStream_t s,
                  const EpPeers* epp) {
  if (H > 1536 || n_splits > 16 || topk + n_shared > 8) throw std::runtime_error("combine_norm: unsupported shape");
  EpPeers ep;
  if (epp) ep = *epp;
  if (!partials && H <= 1280) {  // decode: one float4 per thread, a single round of loads
    DISPATCH_T(dt, (combine_norm_kernel<T, 1, 320, 1><<<(unsigned)rows, 320, 0, s>>>(x, y, perm_pos, topk_w, topk, nullptr, 0, split_stride, w_next, (T*)out16, lo_off_elems, H, eps, n_shared, shared_row0, cap, ep)));
  } else {
    DISPATCH_T(dt, (combine_norm_kernel<T, 16, 128, 3><<<(unsigned)rows, 128, 0, s>>>(x, y, perm_pos, topk_w, topk, partials, partials ? n_splits : 0, split_stride, w_next, (T*)out16, lo_off_elems, H, eps, n_shared, shared_row0, cap, ep)));
  }
  launch_check("moe_combine_norm");
}

}  // namespace dsocr
