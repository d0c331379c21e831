// Fused small-batch decode step over a DSQ snapshot (BASELINE.json configs[3]: batch-1 q4k / q8_0 decode with long
// outputs).  The unfused path (decoder_forward_dsq) spends ~18 launches per layer on 6-14 us kernels; at batch 1 the
// whole quantised model is only ~320 MB per token, so the step is bound by launch count and by how many bytes each
// launch keeps in flight.  Here a layer is 6 launches:
//   qkv       : [pending residual adds + RMSNorm(ln1)] staged in shared memory by every block -> q|k|v GEMVs (3 jobs)
//   attention : RoPE + KV append + split-key attention over the cache, merged by the last block of a (row, head)
//   o_proj    : GEMV
//   router    : residual add + RMSNorm(ln2) + gate GEMV + softmax + top-k           (dense layer: folded into gate/up)
//   gate/up   : routed experts (one job row per (token, slot)) + shared experts, SwiGLU in the epilogue
//   down      : routed + shared experts; the weighted combine is done by the next layer's qkv staging
// All activations stay f32 (run_quantized_matmul semantics, quantization.rs:164-185: y = x . dequant(W)^T).
// GEMV mapping: 8 lanes per output feature (16-byte units of one weight row, 128 contiguous bytes per row and
// iteration), 4 x R features per warp, the next unit's weight bytes are loaded before the current one is consumed and
// the first unit before the activations are staged.
#include <cuda_fp16.h>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <utility>

#include "dsq.h"
#include "kernels.h"
#include "ptx.cuh"

namespace dsocr {

namespace {

constexpr int kThreads = 256, kWarps = 8;
constexpr int kMaxJobs = 3;
constexpr int kMaxSmem = 220 * 1024;

struct Job {
  const uint8_t* p[2][4];  // planes a..d of w0 (and w1 for the dual gate/up job)
  int fmt, K, dual, R, lpr;
  long long N;
  const float* x; long long ldx;
  int groups, rpg, x_row_div;
  const int* row_expert;
  int expert_dep;  // row_expert is written by the kernel launched just before this one
  float* out; long long ldo;
  int block0, fblocks;
};
struct Stage {
  const float* add1; const float* add2;
  const float* ymoe; const float* wmoe; int topk;
  float* write_back;
  const float* norm_w; float eps;
};
struct Launch { Job job[kMaxJobs]; int njobs; int w_off; Stage st; };

// Programmatic dependent launch: a kernel of the step is launched while its predecessor still runs; everything before
// pdl_wait() may only touch data no kernel of the step writes (weights, tables) or data at least two kernels old.
// Dependents are released after the wait, so at most two kernels of the chain overlap (main part of N, prologue of N+1).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  static const bool pdl = getenv("DSOCR_DSQ_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cuda_check(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...), "cudaLaunchKernelEx");
}

// shared-memory index of activation k: 4 floats of padding per 64 keep the 8 lanes of a feature on distinct banks
__device__ __forceinline__ int xpad(int k) { return k + ((k >> 6) << 2); }

// the bytes one lane consumes per step: a 16-byte unit of quants plus its scales
template <int FMT> struct Unit;
template <> struct Unit<8> { uint4 q; __half d; };
template <> struct Unit<12> { uint4 hdr; uint4 q; };
template <> struct Unit<14> { uint4 ql; uint4 qh; int8_t s1, s2; __half d; };
template <> struct Unit<0> { float4 w; };
template <> struct Unit<16> { uint4 w; };  // 8 f16 weights
template <> struct Unit<17> { uint4 w; };  // 8 bf16 weights

// bytes of one weight row in plane p (the layouts of QuantWeight, dsq.h)
template <int FMT>
__device__ __host__ __forceinline__ int plane_row_bytes(int K, int p) {
  if (FMT == 8) return p == 0 ? K : (p == 1 ? K / 16 : 0);
  if (FMT == 12) return p == 0 ? (K / 256) * 144 : 0;
  if (FMT == 14) return p == 0 ? K / 2 : (p == 1 ? K / 4 : (p == 2 ? K / 16 : K / 128));
  if (FMT == 16 || FMT == 17) return p == 0 ? K * 2 : 0;
  return p == 0 ? K * 4 : 0;
}
template <int FMT> struct NPlanes { static constexpr int value = FMT == 8 ? 2 : (FMT == 14 ? 4 : 1); };

__device__ __forceinline__ uint4 lds128(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

// unit u of local row `row` from the block's shared-memory slabs pl[0..3] (same layouts as the global planes)
template <int FMT>
__device__ __forceinline__ void load_unit(Unit<FMT>& o, const uint8_t* const* pl, int row, int K, int u, int rows_pb) {
  if constexpr (FMT == 8) {  // planes: qs int8 [rows][K], d f16 [rows][K/32]
    const int k = u * 16;
    o.q = lds128(pl[0] + row * K + k);
    o.d = reinterpret_cast<const __half*>(pl[1])[row * (K / 32) + (k >> 5)];
  } else if constexpr (FMT == 12) {  // 144-byte blocks as on disk
    const int sb = u >> 3, j = u & 7;
    const uint8_t* blk = pl[0] + (row * (K / 256) + sb) * 144;
    o.hdr = lds128(blk);
    o.q = lds128(blk + 16 + (j >> 1) * 32 + (j & 1) * 16);
  } else if constexpr (FMT == 14) {  // planes: ql [rows][K/2], qh [rows][K/4], sc i8 [rows][K/16], d f16 [rows][K/256]
    const int sb = u >> 3, j = u & 7, half = j >> 2, jj = j & 3, second = jj >> 1, l0 = (jj & 1) * 16;
    o.ql = lds128(pl[0] + row * (K / 2) + sb * 128 + half * 64 + second * 32 + l0);
    o.qh = lds128(pl[1] + row * (K / 4) + sb * 64 + half * 32 + l0);
    const int8_t* sc = reinterpret_cast<const int8_t*>(pl[2]) + row * (K / 16) + sb * 16;
    const int is = half * 8 + (jj & 1) + (second ? 2 : 0);
    o.s1 = sc[is]; o.s2 = sc[is + 4];
    o.d = reinterpret_cast<const __half*>(pl[3])[row * (K / 256) + sb];
  } else if constexpr (FMT == 16 || FMT == 17) {
    // slab = [k-block][row][128 B]; the 16-byte chunk q of a row sits at position q ^ (row & 7) (128B swizzle of the
    // tiled layout; the block's first row is a multiple of 8)
    const int kb = u >> 3, q = u & 7;
    o.w = lds128(pl[0] + ((size_t)kb * rows_pb + row) * 128 + ((q ^ (row & 7)) << 4));
  } else {
    o.w = *reinterpret_cast<const float4*>(pl[0] + ((size_t)row * K + (size_t)u * 4) * 4);
  }
}

__device__ __forceinline__ void lds16(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(p)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

// Byte -> float without the conversion unit (I2F runs at a quarter of the FMA rate and would bound the lm_head GEMV):
// PRMT builds the bit pattern of 2^23 + byte, one FADD removes the offset exactly.
__device__ __forceinline__ void bytes_to_f32(uint32_t w, float bias, float* out) {
  out[0] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440)) - bias;
  out[1] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441)) - bias;
  out[2] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442)) - bias;
  out[3] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7443)) - bias;
}
constexpr float kMagic = 8388608.f;  // 2^23
// Small unsigned values (4 or 6 bits, one per byte): PRMT drops the byte into mantissa bits 16..23 of 128.0f, giving
// 128 + v exactly with no further instruction; the offset leaves through the per-segment activation sums
// (sum (128 + v_i) x_i - 128 sum x_i), which the K-quant min term needs anyway.
__device__ __forceinline__ void small_to_f32_plus128(uint32_t w, float* out) {
  out[0] = __uint_as_float(__byte_perm(w, 0x43000000u, 0x7044));
  out[1] = __uint_as_float(__byte_perm(w, 0x43000000u, 0x7144));
  out[2] = __uint_as_float(__byte_perm(w, 0x43000000u, 0x7244));
  out[3] = __uint_as_float(__byte_perm(w, 0x43000000u, 0x7344));
}

// 6-bit scale / min of sub-block jx of a Q4_K super-block (ggml get_scale_min_k4) from the header words
// (hdr.y|z|w = scales[0..3|4..7|8..11]); shifts instead of byte indexing keep the header in registers.
__device__ __forceinline__ void q4k_scale_min(const uint4& hdr, int jx, int& sc, int& mn) {
  const int sh = 8 * (jx & 3);
  const uint32_t a = (hdr.y >> sh) & 0xFFu, b = (hdr.z >> sh) & 0xFFu, c = (hdr.w >> sh) & 0xFFu;
  if (jx < 4) { sc = (int)(a & 63u); mn = (int)(b & 63u); }
  else { sc = (int)((c & 0xFu) | ((a >> 6) << 4)); mn = (int)((c >> 4) | ((b >> 6) << 4)); }
}

// contribution of one unit of one weight row to the MT token rows staged in xs (pitch Kp)
template <int FMT, int MT>
__device__ __forceinline__ void dot_unit(const Unit<FMT>& w, const float* xs, int Kp, const float* xsum, int Sp, int u, float* acc) {
  if constexpr (FMT == 8) {
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(&w.q);
    float wq[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) bytes_to_f32(qw[i] ^ 0x80808080u, kMagic + 128.f, wq + 4 * i);  // int8 + 128 as a byte
    const float d = __half2float(w.d);
    const int xo = xpad(u * 16);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float xv[16];
      lds16(xs + m * Kp + xo, xv);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s = fmaf(wq[i], xv[i], s);
      acc[m] = fmaf(d, s, acc[m]);
    }
  } else if constexpr (FMT == 12) {
    const int sb = u >> 3, j = u & 7, gq = j >> 1, lo = (j & 1) * 16;
    const int k1 = sb * 256 + gq * 64 + lo, k2 = k1 + 32;
    const __half2 dd = *reinterpret_cast<const __half2*>(&w.hdr.x);
    const float d = __low2float(dd), dmin = __high2float(dd);
    int sc1, m1, sc2, m2;  // get_scale_min_k4 for sub-blocks 2*gq and 2*gq + 1
    q4k_scale_min(w.hdr, 2 * gq, sc1, m1);
    q4k_scale_min(w.hdr, 2 * gq + 1, sc2, m2);
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(&w.q);
    float w1[16], w2[16];  // 128 + nibble
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      small_to_f32_plus128(qw[i] & 0x0F0F0F0Fu, w1 + 4 * i);
      small_to_f32_plus128((qw[i] >> 4) & 0x0F0F0F0Fu, w2 + 4 * i);
    }
    const float ds1 = d * (float)sc1, ds2 = d * (float)sc2;
    const float c1 = 128.f * ds1 + dmin * (float)m1, c2 = 128.f * ds2 + dmin * (float)m2;
    const int xo1 = xpad(k1), xo2 = xpad(k2);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float x1[16], x2[16];
      lds16(xs + m * Kp + xo1, x1);
      lds16(xs + m * Kp + xo2, x2);
      const float sx1 = xsum[m * Sp + (k1 >> 4)], sx2 = xsum[m * Sp + (k2 >> 4)];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { s1 = fmaf(w1[i], x1[i], s1); s2 = fmaf(w2[i], x2[i], s2); }
      acc[m] += (ds1 * s1 + ds2 * s2) - (c1 * sx1 + c2 * sx2);
    }
  } else if constexpr (FMT == 14) {
    const int sb = u >> 3, j = u & 7, half = j >> 2, jj = j & 3, second = jj >> 1, l0 = (jj & 1) * 16;
    const int kb = sb * 256 + half * 128 + l0;
    const int k1 = kb + (second ? 32 : 0), k2 = kb + (second ? 96 : 64);
    const int sh = second ? 2 : 0;
    const float d = __half2float(w.d);
    const float d1 = d * (float)w.s1, d2 = d * (float)w.s2;
    const uint32_t* lw = reinterpret_cast<const uint32_t*>(&w.ql);
    const uint32_t* hw = reinterpret_cast<const uint32_t*>(&w.qh);
    float w1[16], w2[16];  // 128 + (low nibble | two high bits << 4); the stored value is that minus 32
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t a = (lw[i] & 0x0F0F0F0Fu) | (((hw[i] >> sh) & 0x03030303u) << 4);
      const uint32_t b = ((lw[i] >> 4) & 0x0F0F0F0Fu) | (((hw[i] >> (sh + 4)) & 0x03030303u) << 4);
      small_to_f32_plus128(a, w1 + 4 * i);
      small_to_f32_plus128(b, w2 + 4 * i);
    }
    const int xo1 = xpad(k1), xo2 = xpad(k2);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float x1[16], x2[16];
      lds16(xs + m * Kp + xo1, x1);
      lds16(xs + m * Kp + xo2, x2);
      const float sx1 = xsum[m * Sp + (k1 >> 4)], sx2 = xsum[m * Sp + (k2 >> 4)];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { s1 = fmaf(w1[i], x1[i], s1); s2 = fmaf(w2[i], x2[i], s2); }
      acc[m] += d1 * (s1 - 160.f * sx1) + d2 * (s2 - 160.f * sx2);
    }
  } else if constexpr (FMT == 16 || FMT == 17) {
    const uint32_t* ww = reinterpret_cast<const uint32_t*>(&w.w);
    float wf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if constexpr (FMT == 16) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ww[i]));
        wf[2 * i] = f.x; wf[2 * i + 1] = f.y;
      } else {
        wf[2 * i] = __uint_as_float(ww[i] << 16); wf[2 * i + 1] = __uint_as_float(ww[i] & 0xFFFF0000u);
      }
    }
    const int xo = xpad(u * 8);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float4 a = *reinterpret_cast<const float4*>(xs + m * Kp + xo);
      const float4 b = *reinterpret_cast<const float4*>(xs + m * Kp + xo + 4);
      float s0 = wf[0] * a.x;
      s0 = fmaf(wf[1], a.y, s0); s0 = fmaf(wf[2], a.z, s0); s0 = fmaf(wf[3], a.w, s0);
      s0 = fmaf(wf[4], b.x, s0); s0 = fmaf(wf[5], b.y, s0); s0 = fmaf(wf[6], b.z, s0); s0 = fmaf(wf[7], b.w, s0);
      acc[m] += s0;
    }
  } else {
    const int xo = xpad(u * 4);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + m * Kp + xo);
      acc[m] += w.w.x * xv.x + w.w.y * xv.y + w.w.z * xv.z + w.w.w * xv.w;
    }
  }
}

__device__ __forceinline__ void add4(float4& a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// LPR = lanes per weight row: 8 for the wide lm_head (few shuffles, R = 2 rows share the staged activations), 32 for the
// small projections of a batch-1 step (4x the warps and a quarter of the serial work per lane).
template <int FMT, int MT, int R, int NW, int LPR>
__device__ __forceinline__ void run_job(const Job& J, const Stage& st, float* xs, uint8_t* wsm, uint64_t* bar, float* red,
                                        int local) {
  constexpr int RPW = 32 / LPR;  // weight-row groups per warp
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, grp = lane / LPR, sub = lane % LPR;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const int K = J.K, Kp = xpad(K), rpg = J.rpg;
  // (integer division by a run-time value costs ~25 instructions; most jobs have one group)
  const int g = J.groups == 1 ? 0 : local / J.fblocks, fb = local - g * J.fblocks;
  const int rows_pb = nwarps * RPW * R;                   // weight rows (output features) of a block
  const long long nb0 = (long long)fb * rows_pb;
  const int rows_here = (int)min((long long)rows_pb, J.N - nb0);
  const int units = FMT == 8 ? K / 16 : (FMT == 0 ? K / 4 : ((FMT == 16 || FMT == 17) ? K / 8 : K / 32));
  constexpr int NP = NPlanes<FMT>::value;

  // ---- the block's weight rows are contiguous in every plane: bulk copies (TMA) bring the whole slab into shared
  // memory while the threads stage the activations; no registers are tied up by bytes in flight
  const uint8_t* pl[NW][4];
  {
    uint32_t off = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        pl[w][p] = wsm + off;
        if (p < NP) off += (uint32_t)rows_pb * plane_row_bytes<FMT>(K, p);
      }
    auto issue = [&]() {
      const long long e = J.row_expert ? J.row_expert[g] : 0;
      if constexpr (FMT == 16 || FMT == 17) {  // one copy per 64-wide k-block: rows_here x 128 B contiguous inside a tile
        const long long grow = e * J.N + nb0;
        const int num_kb = K / 64;
        ptx::mbar_expect_tx(bar, (uint32_t)rows_here * (uint32_t)K * 2u * NW);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const uint8_t* src = J.p[w][0] + ((size_t)(grow >> 7) * num_kb) * 16384 + (size_t)(grow & 127) * 128;
          uint8_t* dst = const_cast<uint8_t*>(pl[w][0]);
          for (int kb = 0; kb < num_kb; ++kb)
            ptx::bulk_load(dst + (size_t)kb * rows_pb * 128, src + (size_t)kb * 16384, (uint32_t)rows_here * 128u, bar);
        }
        return;
      }
      uint32_t total = 0;
#pragma unroll
      for (int p = 0; p < NP; ++p) total += (uint32_t)rows_here * plane_row_bytes<FMT>(K, p);
      ptx::mbar_expect_tx(bar, total * NW);
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t rb = plane_row_bytes<FMT>(K, p);
          const uint8_t* src = J.p[w][p] + (size_t)(e * J.N + nb0) * rb;
          uint8_t* dst = const_cast<uint8_t*>(pl[w][p]);
          const uint32_t bytes = (uint32_t)rows_here * rb;
          for (uint32_t o = 0; o < bytes; o += 16384) ptx::bulk_load(dst + o, src + o, min(16384u, bytes - o), bar);
        }
    };
    const bool early = !(J.row_expert && J.expert_dep);  // the weight stream does not depend on the predecessor
    if (t == 0) {
      ptx::mbar_init(bar, 1);
      ptx::fence_barrier_init();
      if (early) issue();
    }
    pdl_wait();
    pdl_release();
    if (t == 0 && !early) issue();
  }

  // ---- stage the token rows of this group: x = base + (sum_j w_j y_j + add1 + add2), optional RMSNorm weight
  float ss[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) ss[m] = 0.f;
  const bool wb = st.write_back != nullptr && blockIdx.x == 0;
  const int n4 = K / 4, Sp = K / 16;
  float* xsum = xs + MT * Kp;  // [MT][K/16] sums of the staged activations over 16-element segments
  for (int i0 = 0; i0 < n4; i0 += nthreads) {  // whole warps iterate together (segment sums use shuffles)
    const int i = i0 + t;
    const bool in = i < n4;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m >= rpg) break;
      const long long xr = J.x_row_div == 1 ? (long long)g * rpg + m : ((long long)g * rpg + m) / J.x_row_div;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) {
        v = reinterpret_cast<const float4*>(J.x + xr * J.ldx)[i];
        if (st.ymoe || st.add1 || st.add2) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (st.ymoe) {
            float wj[8]; float4 y[8];  // all expert rows are requested before the first one is consumed
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const bool on = j < st.topk;
              wj[j] = on ? st.wmoe[xr * st.topk + j] : 0.f;
              y[j] = on ? reinterpret_cast<const float4*>(st.ymoe + (xr * st.topk + j) * K)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              a.x = fmaf(wj[j], y[j].x, a.x); a.y = fmaf(wj[j], y[j].y, a.y); a.z = fmaf(wj[j], y[j].z, a.z); a.w = fmaf(wj[j], y[j].w, a.w);
            }
          }
          if (st.add1) add4(a, reinterpret_cast<const float4*>(st.add1 + xr * K)[i]);
          if (st.add2) add4(a, reinterpret_cast<const float4*>(st.add2 + xr * K)[i]);
          add4(v, a);
        }
        if (wb) reinterpret_cast<float4*>(st.write_back + xr * K)[i] = v;
        if (st.norm_w) {
          ss[m] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          const float4 nw = reinterpret_cast<const float4*>(st.norm_w)[i];
          v.x *= nw.x; v.y *= nw.y; v.z *= nw.z; v.w *= nw.w;
        }
        *reinterpret_cast<float4*>(xs + m * Kp + xpad(4 * i)) = v;
      }
      if constexpr (FMT == 12 || FMT == 14) {
        float sg = (v.x + v.y) + (v.z + v.w);
        sg += __shfl_xor_sync(0xffffffffu, sg, 1);
        sg += __shfl_xor_sync(0xffffffffu, sg, 2);
        if (in && (i & 3) == 0) xsum[m * Sp + (i >> 2)] = sg;
      }
    }
  }
  if (st.norm_w) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float v = ss[m];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[m * kWarps + warp] = v;
    }
  }
  __syncthreads();  // activations staged; the barrier initialised by thread 0 is visible
  float rs[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    rs[m] = 1.f;
    if (st.norm_w) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += w < nwarps ? red[m * kWarps + w] : 0.f;
      rs[m] = rsqrtf(tot / (float)K + st.eps);
    }
  }

  float acc[NW][R][MT];
#pragma unroll
  for (int w = 0; w < NW; ++w)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[w][r][m] = 0.f;

  const int lr0 = (warp * RPW + grp) * R;  // first local weight row of this lane group
  int lrow[R];
#pragma unroll
  for (int r = 0; r < R; ++r) lrow[r] = min(lr0 + r, rows_here - 1);
  ptx::mbar_wait(bar, 0);  // weight slab landed
  // not unrolled: a step's kernels run for a few microseconds each and ncu attributes ~49 % of their stall samples to
  // instruction fetch (stall_no_inst); the executed path is kept short instead
#pragma unroll 1
  for (int u = sub; u < units; u += LPR) {
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        Unit<FMT> un;
        load_unit<FMT>(un, pl[w], lrow[r], K, u, rows_pb);
        dot_unit<FMT, MT>(un, xs, Kp, xsum, Sp, u, acc[w][r]);
      }
  }
#pragma unroll
  for (int w = 0; w < NW; ++w)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float v = acc[w][r][m];
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[w][r][m] = v;
      }
  if (sub == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (lr0 + r >= rows_here) break;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        if (m >= rpg) break;
        float a = acc[0][r][m] * rs[m];
        if constexpr (NW == 2) {
          const float b = acc[1][r][m] * rs[m];
          a = a / (1.f + __expf(-a)) * b;  // SiLU(gate) * up (run_dense_mlp, block.rs:1166-1177)
        }
        J.out[((long long)g * rpg + m) * J.ldo + nb0 + lr0 + r] = a;
      }
    }
  }
}

template <int FMT, int MT>
__device__ __forceinline__ void run_fmt(const Job& J, const Stage& st, float* xs, uint8_t* wsm, uint64_t* bar, float* red,
                                        int local) {
  // lanes per weight row: the K-quants have 40 16-byte units per 1280-wide row (8 lanes: 5 steps each, no idle lane),
  // Q8_0 has 80 (16 lanes: 5 steps; 1821 vs 1670 tok/s with 32), 16-bit rows 160 (32 lanes: 5 steps)
  constexpr int LPR = (FMT == 12 || FMT == 14) ? 8 : ((FMT == 16 || FMT == 17) ? 32 : 16);
  if (J.R == 2) run_job<FMT, MT, 2, 1, 8>(J, st, xs, wsm, bar, red, local);
  else if (J.dual) run_job<FMT, MT, 1, 2, LPR>(J, st, xs, wsm, bar, red, local);
  else run_job<FMT, MT, 1, 1, LPR>(J, st, xs, wsm, bar, red, local);
}

// One kernel per (format pair, MT): the jobs of a launch use at most two block formats (the routed down projection
// falls back to Q8_0 because 896 % 256 != 0 while the shared one keeps the K-quant).
template <int FA, int FB, int MT>
__global__ void __launch_bounds__(kThreads)
dsq_fused_gemv_kernel(const __grid_constant__ Launch L) {
  extern __shared__ __align__(128) float xs[];  // [MT][Kp] activations | weight slabs at byte offset L.w_off
  __shared__ float red[MT * kWarps];
  __shared__ uint64_t bar;
  uint8_t* wsm = reinterpret_cast<uint8_t*>(xs) + L.w_off;
  int j = 0;
  for (int i = 1; i < L.njobs; ++i)
    if ((int)blockIdx.x >= L.job[i].block0) j = i;
  const Job& J = L.job[j];
  const int local = (int)blockIdx.x - J.block0;
  if (FA == FB || J.fmt == FA) run_fmt<FA, MT>(J, L.st, xs, wsm, &bar, red, local);
  else run_fmt<FB, MT>(J, L.st, xs, wsm, &bar, red, local);
}

template <int FA, int FB>
void launch_pair(const Launch& L, int blocks, int threads, size_t smem, int max_rpg, cudaStream_t stream) {
  static PerDeviceOnce once;  // per instantiation
  once.run([&] {
    cuda_check(cudaFuncSetAttribute(dsq_fused_gemv_kernel<FA, FB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem), "smem attr");
    cuda_check(cudaFuncSetAttribute(dsq_fused_gemv_kernel<FA, FB, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem), "smem attr");
  });
  if (max_rpg == 1) launch_pdl(dsq_fused_gemv_kernel<FA, FB, 1>, dim3(blocks), dim3(threads), smem, stream, L);
  else launch_pdl(dsq_fused_gemv_kernel<FA, FB, 4>, dim3(blocks), dim3(threads), smem, stream, L);
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

// residual add + RMSNorm(ln2) + router (run_moe, block.rs:1263-1301: f32 gate logits -> softmax -> top-k, ties to
// the lowest index).  NB blocks per token row: each normalises the row on its own and computes the logits of E/NB
// experts (one block pulling the whole 328 KB gate weight through one SM took ~9 us); the block that arrives last
// does softmax + top-k.  Block 0 writes the new residual and the normalised row.
template <int E, int NB>
__global__ void __launch_bounds__(256)
dsq_router_kernel(const float* __restrict__ base, const float* __restrict__ add1, float* __restrict__ xout,
                  const float* __restrict__ w, const float* __restrict__ wgt, float* __restrict__ xn32,
                  float* __restrict__ logits_ws, int* __restrict__ counters, int* __restrict__ topk_idx,
                  float* __restrict__ topk_w, int H, int topk, float eps) {
  constexpr int EPB = E / NB;
  constexpr int KS = 256 / EPB;
  constexpr int PER = (E + 31) / 32;
  extern __shared__ float sm[];
  float* xn_s = sm;      // [H]
  float* part = sm + H;  // [256]
  __shared__ float red[8];
  __shared__ int s_last;
  const long long row = blockIdx.x;
  const int blk = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int n4 = H / 4;
  // the gate weight (f32 [H][E], shared by the NB blocks of a row) is pulled towards L2 while o_proj still runs
  for (int line = blk * 256 + t; line * 32 < H * E; line += NB * 256) prefetch_l2(wgt + (size_t)line * 32);
  pdl_wait();
  pdl_release();
  float ss = 0.f;
  for (int i = t; i < n4; i += 256) {
    float4 v = reinterpret_cast<const float4*>(base + row * H)[i];
    add4(v, reinterpret_cast<const float4*>(add1 + row * H)[i]);
    if (blk == 0) reinterpret_cast<float4*>(xout + row * H)[i] = v;
    reinterpret_cast<float4*>(xn_s)[i] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = rsqrtf(tot / (float)H + eps);
  for (int i = t; i < n4; i += 256) {
    float4 v = reinterpret_cast<float4*>(xn_s)[i];
    const float4 ww = reinterpret_cast<const float4*>(w)[i];
    v.x = v.x * inv * ww.x; v.y = v.y * inv * ww.y; v.z = v.z * inv * ww.z; v.w = v.w * inv * ww.w;
    reinterpret_cast<float4*>(xn_s)[i] = v;
    if (blk == 0) reinterpret_cast<float4*>(xn32 + row * H)[i] = v;
  }
  __syncthreads();
  {
    const int e = blk * EPB + t % EPB, ks = t / EPB;
    const int kper = H / KS;
    const float* wp = wgt + (long long)(ks * kper) * E + e;
    const float* xp = xn_s + ks * kper;
    float acc = 0.f;
#pragma unroll 20
    for (int k = 0; k < kper; ++k) acc = fmaf(xp[k], wp[(long long)k * E], acc);
    part[t] = acc;
  }
  __syncthreads();
  if (t < EPB) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < KS; ++q) s += part[q * EPB + t];
    logits_ws[row * E + blk * EPB + t] = s;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(&counters[row], 1) == NB - 1;
  __syncthreads();
  if (!s_last || warp != 0) return;
  __threadfence();
  float p[PER];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int ee = j * 32 + lane;
    p[j] = ee < E ? __ldcg(logits_ws + row * E + ee) : -INFINITY;
    mx = fmaxf(mx, p[j]);
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
  if (lane < topk) {
    topk_idx[row * topk + lane] = my_e;
    topk_w[row * topk + lane] = my_w;
  }
  if (lane == 0) counters[row] = 0;  // ready for the next launch (graph replay)
}

// x = base + (sum_j w_j y_j + add1);  out = RMSNorm(x) * w   (last layer's MoE combine + final norm), block per row
__global__ void __launch_bounds__(256)
dsq_combine_norm_kernel(const float* __restrict__ base, const float* __restrict__ ymoe, const float* __restrict__ wmoe,
                        int topk, const float* __restrict__ add1, const float* __restrict__ w, float* __restrict__ out,
                        int H, float eps) {
  extern __shared__ float sm[];  // [H]
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const int t = threadIdx.x;
  pdl_wait();
  pdl_release();
  float ss = 0.f;
  for (int i = t; i < H / 4; i += 256) {
    float4 v = reinterpret_cast<const float4*>(base + row * H)[i];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ymoe) {
      float wj[8]; float4 y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool on = j < topk;
        wj[j] = on ? wmoe[row * topk + j] : 0.f;
        y[j] = on ? reinterpret_cast<const float4*>(ymoe + (row * topk + j) * H)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a.x = fmaf(wj[j], y[j].x, a.x); a.y = fmaf(wj[j], y[j].y, a.y); a.z = fmaf(wj[j], y[j].z, a.z); a.w = fmaf(wj[j], y[j].w, a.w);
      }
    }
    if (add1) add4(a, reinterpret_cast<const float4*>(add1 + row * H)[i]);
    add4(v, a);
    reinterpret_cast<float4*>(sm)[i] = v;
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = warp_sum(ss);
  if ((t & 31) == 0) red[t >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = rsqrtf(tot / (float)H + eps);
  for (int i = t; i < H / 4; i += 256) {
    float4 v = reinterpret_cast<float4*>(sm)[i];
    const float4 ww = reinterpret_cast<const float4*>(w)[i];
    v.x = v.x * inv * ww.x; v.y = v.y * inv * ww.y; v.z = v.z * inv * ww.z; v.w = v.w * inv * ww.w;
    reinterpret_cast<float4*>(out + row * H)[i] = v;
  }
}

__device__ __forceinline__ void ld16f(const float* p, float* out) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(p)[i];
    out[4 * i] = t.x; out[4 * i + 1] = t.y; out[4 * i + 2] = t.z; out[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ld16f(const __half* p, float* out) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint4 t = reinterpret_cast<const uint4*>(p)[i];
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); out[8 * i + 2 * j] = f.x; out[8 * i + 2 * j + 1] = f.y; }
  }
}

// RoPE + KV append + attention for a handful of query rows (attention_forward, block.rs:446-804, q = 1): the keys of
// a (row, head) are cut into `nsplit` contiguous ranges, one block each (one block per (row, head) would leave a
// batch-1 step with 10 blocks streaming the whole cache); each block leaves (max, sum, weighted V) in a workspace and
// the block that arrives last merges them in split order (deterministic) and resets the arrival counter.
constexpr int kPartStride = 136;  // m, l, pad, acc[128]
template <typename TKV>
__global__ void __launch_bounds__(128)
dsq_attn_split_kernel(const float* __restrict__ qkv, const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                      TKV* __restrict__ kc, TKV* __restrict__ vc, const int* __restrict__ row_page,
                      const int* __restrict__ row_pos, float* __restrict__ part, int* __restrict__ counters,
                      float* __restrict__ ctx, int heads, int smax, float scale, int nsplit) {
  constexpr int D = 128;
  const int rh = blockIdx.x, r = rh / heads, hd = rh % heads, sp = blockIdx.y;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, grp = lane >> 3, sub = lane & 7;
  const int pos = row_pos[r], page = row_page[r];
  TKV* kbase = kc + ((long long)page * heads + hd) * smax * D;
  TKV* vbase = vc + ((long long)page * heads + hd) * smax * D;
  __shared__ float q_s[D], knew[D], vnew[D];
  __shared__ float sm_m[16], sm_l[16], sm_acc[16][D];
  __shared__ int s_last;
  const bool tail = sp == nsplit - 1;  // this block also owns the new token
  const int chunk = (((pos + nsplit - 1) / nsplit) + 31) & ~31;  // cached keys 0..pos-1, 32 per block iteration
  const int kb = sp * chunk, ke = min(pos, kb + chunk);
  // this block's K/V range was written by earlier steps: pull it towards L2 while the qkv projection still runs
  for (int line = t; line * (128 / (int)sizeof(TKV)) < (ke - kb) * D; line += 128) {
    prefetch_l2(kbase + (long long)kb * D + (long long)line * (128 / (int)sizeof(TKV)));
    prefetch_l2(vbase + (long long)kb * D + (long long)line * (128 / (int)sizeof(TKV)));
  }
  pdl_wait();
  pdl_release();
  const float* base = qkv + (long long)r * 3 * heads * D;
  if (t < 64) {
    const float c = cos_t[(long long)pos * 64 + t], sn = sin_t[(long long)pos * 64 + t];
    const float qlo = base[hd * D + t], qhi = base[hd * D + 64 + t];
    q_s[t] = (qlo * c - qhi * sn) * scale;
    q_s[t + 64] = (qhi * c + qlo * sn) * scale;
    if (tail) {
      const float klo = base[(heads + hd) * D + t], khi = base[(heads + hd) * D + 64 + t];
      const TKV k0 = (TKV)(klo * c - khi * sn), k1 = (TKV)(khi * c + klo * sn);
      kbase[(long long)pos * D + t] = k0;
      kbase[(long long)pos * D + 64 + t] = k1;
      knew[t] = (float)k0; knew[t + 64] = (float)k1;  // what later steps read back from the cache
    }
  } else if (tail) {
    const int d = t - 64;
    const TKV v0 = (TKV)base[(2 * heads + hd) * D + d], v1 = (TKV)base[(2 * heads + hd) * D + 64 + d];
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
  const TKV* kr = kbase;
  const TKV* vr = vbase;
  for (int ka = kb + warp * 4 + grp; ka < ke; ka += 32) {  // two keys per lane group and iteration
    const int kb2 = ka + 16;
    const bool okb = kb2 < ke;
    float k1[16], v1[16], k2[16], v2[16];
    ld16f(kr + (long long)ka * D + sub * 16, k1);
    ld16f(vr + (long long)ka * D + sub * 16, v1);
    if (okb) { ld16f(kr + (long long)kb2 * D + sub * 16, k2); ld16f(vr + (long long)kb2 * D + sub * 16, v2); }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s1 += qv[i] * k1[i];
    if (okb) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s2 += qv[i] * k2[i];
    }
    const unsigned gmask = 0xffu << (grp * 8);  // the 8 lanes of a group share ka
    s1 += __shfl_xor_sync(gmask, s1, 1); s1 += __shfl_xor_sync(gmask, s1, 2); s1 += __shfl_xor_sync(gmask, s1, 4);
    s2 += __shfl_xor_sync(gmask, s2, 1); s2 += __shfl_xor_sync(gmask, s2, 2); s2 += __shfl_xor_sync(gmask, s2, 4);
    fold(s1, v1);
    if (okb) fold(s2, v2);
  }
  if (tail && warp == 0 && grp == 0) {  // the new token itself (key index pos)
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
  if (nsplit == 1) {
    ctx[((long long)r * heads + hd) * D + t] = num / den;
    return;
  }
  float* my = part + ((long long)rh * nsplit + sp) * kPartStride;
  if (t == 0) { my[0] = gm; my[1] = den; }
  my[8 + t] = num;
  __threadfence();
  __syncthreads();
  if (t == 0) s_last = atomicAdd(&counters[rh], 1) == nsplit - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* all = part + (long long)rh * nsplit * kPartStride;
  // merge in split order: the per-split (max, sum) pairs meet in shared memory first, then every thread sums its own
  // output dimension with the loads of 8 splits in flight at a time
  float* mm = &sm_acc[0][0];      // [nsplit] max   (the per-block merge above is done with sm_acc)
  float* dd = &sm_acc[1][0];      // [nsplit] sum
  if (t < nsplit) { mm[t] = __ldcg(all + t * kPartStride); dd[t] = __ldcg(all + t * kPartStride + 1); }
  __syncthreads();
  float tm = -INFINITY;
  for (int s = 0; s < nsplit; ++s) tm = fmaxf(tm, mm[s]);
  float tn = 0.f, td = 0.f;
  for (int s0 = 0; s0 < nsplit; s0 += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = s0 + i < nsplit ? __ldcg(all + (s0 + i) * kPartStride + 8 + t) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (s0 + i < nsplit) {
        const float f = mm[s0 + i] == -INFINITY ? 0.f : __expf(mm[s0 + i] - tm);
        tn += f * v[i];
        td += f * dd[s0 + i];
      }
    }
  }
  ctx[((long long)r * heads + hd) * D + t] = tn / td;
  if (t == 0) counters[rh] = 0;  // ready for the next step (graph replay)
}

}  // namespace

FusedWeight fused_weight(const QuantWeight& w) {
  FusedWeight f;
  switch (w.fmt) {
    case DsqDType::Q8_0: f.fmt = 8; break;
    case DsqDType::Q4K: f.fmt = 12; break;
    case DsqDType::Q6K: f.fmt = 14; break;
    default: f.fmt = 0; break;
  }
  f.N = w.N; f.K = w.K;
  f.p[0] = w.a.p; f.p[1] = w.b.p; f.p[2] = w.c.p; f.p[3] = w.d.p;
  return f;
}
FusedWeight fused_weight_tiled16(const void* tiled, long long N, int K, bool bf16) {
  FusedWeight f;
  f.fmt = bf16 ? 17 : 16; f.N = N; f.K = K; f.p[0] = tiled;
  return f;
}

namespace {

}  // namespace

void dsq_fused_gemv(const DsqFusedJob* jobs, int njobs, const DsqFusedStage& st, const char* tag, cudaStream_t stream) {
  if (njobs < 1 || njobs > kMaxJobs) throw std::runtime_error("dsq_fused_gemv: 1..3 jobs");
  Launch L{};
  L.njobs = njobs;
  int blocks = 0, max_rpg = 1;
  size_t smem = 0;
  for (int i = 0; i < njobs; ++i) {
    const DsqFusedJob& s = jobs[i];
    const FusedWeight& w = s.w0;
    Job& J = L.job[i];
    J.fmt = w.fmt; J.K = w.K; J.N = w.N; J.dual = s.w1.valid() ? 1 : 0;
    const int gran = J.fmt == 8 ? 32 : (J.fmt == 0 ? 16 : (J.fmt >= 16 ? 64 : 256));
    if (!w.valid() || w.K % gran) throw std::runtime_error("dsq_fused_gemv: unsupported K for this weight format");
    if (J.dual && (s.w1.fmt != w.fmt || s.w1.K != w.K || s.w1.N != w.N)) throw std::runtime_error("dsq_fused_gemv: gate/up formats differ");
    if (s.rpg < 1 || s.rpg > 4 || s.groups < 1) throw std::runtime_error("dsq_fused_gemv: 1..4 rows per group");
    for (int q = 0; q < 4; ++q) { J.p[0][q] = (const uint8_t*)w.p[q]; J.p[1][q] = (const uint8_t*)s.w1.p[q]; }
    J.R = (!J.dual && w.N >= 16384 && w.N % 64 == 0) ? 2 : 1;  // wide layers (lm_head): 8 lanes per row, two rows per lane group
    J.lpr = J.R == 2 ? 8 : ((J.fmt == 12 || J.fmt == 14) ? 8 : (J.fmt >= 16 ? 32 : 16));  // == run_fmt's LPR
    J.x = s.x; J.ldx = s.ldx; J.groups = s.groups; J.rpg = s.rpg; J.x_row_div = s.x_row_div < 1 ? 1 : s.x_row_div;
    J.row_expert = s.row_expert; J.expert_dep = s.expert_dep ? 1 : 0; J.out = s.out; J.ldo = s.ldo;
    J.block0 = 0; J.fblocks = 0;
    max_rpg = std::max(max_rpg, s.rpg);
    if (w.N % 8) throw std::runtime_error("dsq_fused_gemv: weight rows must be a multiple of 8");
    const size_t kp = (size_t)w.K + ((size_t)w.K >> 6) * 4 + 4 + (size_t)w.K / 16;  // padded row + its segment sums
    smem = std::max(smem, kp * 4);
  }
  // shared memory: [MT][Kp] activations, then the block's weight slabs (rows_pb rows of every plane, x2 for gate/up)
  const int MT = max_rpg == 1 ? 1 : 4;
  const size_t x_bytes = (smem * MT + 127) / 128 * 128;
  auto row_bytes = [](const Job& J) {
    size_t rb = 0;
    for (int p = 0; p < 4; ++p)
      rb += J.fmt == 8 ? plane_row_bytes<8>(J.K, p) : J.fmt == 12 ? plane_row_bytes<12>(J.K, p)
          : J.fmt == 14 ? plane_row_bytes<14>(J.K, p) : J.fmt >= 16 ? plane_row_bytes<16>(J.K, p) : plane_row_bytes<0>(J.K, p);
    return rb * (J.dual ? 2 : 1);
  };
  // Weight rows per 256-thread block: 8 (one warp per row) or 64 (lm_head: 8 lanes per row, 2 rows per lane group).
  int threads = kThreads;
  for (int i = 0; i < njobs; ++i)
    if (L.job[i].R == 2 && L.job[i].fmt >= 16) threads = 128;  // 16-bit lm_head: 32 rows (80 KB) per block, two blocks per SM
  size_t slab = 0;
  for (;; threads /= 2) {  // long rows (K = 6848) with 4 token rows staged: fewer weight rows per block
    slab = 0; blocks = 0;
    int min_rows = 1 << 30;
    for (int i = 0; i < njobs; ++i) {
      Job& J = L.job[i];
      const int per_block = (threads / 32) * (32 / J.lpr) * J.R;
      J.fblocks = (int)((J.N + per_block - 1) / per_block);
      J.block0 = blocks;
      blocks += J.fblocks * J.groups;
      slab = std::max(slab, row_bytes(J) * per_block);
      min_rows = std::min(min_rows, per_block);
    }
    if (x_bytes + slab <= (size_t)kMaxSmem || min_rows < 16) break;  // a block keeps >= 8 rows (slab alignment, swizzle phase)
  }
  if (x_bytes + slab > (size_t)kMaxSmem) {
    // long rows (the dense down projection, K = 6848) with several token rows staged: one launch per token row
    const bool plain = !st.add1 && !st.add2 && !st.ymoe && !st.write_back && !st.norm_w;
    if (njobs == 1 && jobs[0].groups == 1 && jobs[0].rpg > 1 && jobs[0].x_row_div <= 1 && !jobs[0].row_expert && plain) {
      for (int m = 0; m < jobs[0].rpg; ++m) {
        DsqFusedJob one = jobs[0];
        one.rpg = 1;
        one.x = jobs[0].x + (long long)m * jobs[0].ldx;
        one.out = jobs[0].out + (long long)m * jobs[0].ldo;
        dsq_fused_gemv(&one, 1, st, tag, stream);
      }
      return;
    }
    throw std::runtime_error("dsq_fused_gemv: weight rows do not fit in shared memory");
  }
  L.w_off = (int)x_bytes;
  smem = x_bytes + slab;
  L.st.add1 = st.add1; L.st.add2 = st.add2; L.st.ymoe = st.ymoe; L.st.wmoe = st.wmoe; L.st.topk = st.topk;
  L.st.write_back = st.write_back; L.st.norm_w = st.norm_w; L.st.eps = st.eps;
  int fa = L.job[0].fmt, fb = fa;
  for (int i = 1; i < njobs; ++i) {
    if (L.job[i].fmt == fa || L.job[i].fmt == fb) continue;
    if (fb != fa) throw std::runtime_error("dsq_fused_gemv: more than two block formats in one launch");
    fb = L.job[i].fmt;
  }
  if (fa > fb) std::swap(fa, fb);
  const int key = fa * 16 + fb;
  switch (key) {
    case 8 * 16 + 8: launch_pair<8, 8>(L, blocks, threads, smem, max_rpg, stream); break;
    case 12 * 16 + 12: launch_pair<12, 12>(L, blocks, threads, smem, max_rpg, stream); break;
    case 14 * 16 + 14: launch_pair<14, 14>(L, blocks, threads, smem, max_rpg, stream); break;
    case 0: launch_pair<0, 0>(L, blocks, threads, smem, max_rpg, stream); break;
    case 8 * 16 + 12: launch_pair<8, 12>(L, blocks, threads, smem, max_rpg, stream); break;
    case 8 * 16 + 14: launch_pair<8, 14>(L, blocks, threads, smem, max_rpg, stream); break;
    case 0 * 16 + 8: launch_pair<0, 8>(L, blocks, threads, smem, max_rpg, stream); break;
    case 16 * 16 + 16: launch_pair<16, 16>(L, blocks, threads, smem, max_rpg, stream); break;
    case 17 * 16 + 17: launch_pair<17, 17>(L, blocks, threads, smem, max_rpg, stream); break;
    default: throw std::runtime_error("dsq_fused_gemv: unsupported block format pair");
  }
  launch_check(tag);
}

void dsq_router(const float* base, const float* add1, float* xout, const float* w, const float* wgt, float* xn32,
                float* logits_ws, int* counters, int* topk_idx, float* topk_w, long long rows, int H, int E, int topk,
                float eps, cudaStream_t s) {
  const size_t smem = (size_t)(H + 256) * 4;
  if (H % 64 || topk > 32) throw std::runtime_error("dsq_router: unsupported shape");
  if (E == 64) launch_pdl(dsq_router_kernel<64, 16>, dim3((unsigned)rows, 16), dim3(256), smem, s, base, add1, xout, w, wgt, xn32, logits_ws, counters, topk_idx, topk_w, H, topk, eps);
  else if (E == 16) launch_pdl(dsq_router_kernel<16, 4>, dim3((unsigned)rows, 4), dim3(256), smem, s, base, add1, xout, w, wgt, xn32, logits_ws, counters, topk_idx, topk_w, H, topk, eps);
  else throw std::runtime_error("dsq_router: unsupported expert count");
  launch_check("dsq_router");
}

void dsq_combine_norm(const float* base, const float* ymoe, const float* wmoe, int topk, const float* add1,
                      const float* w, float* out, long long rows, int H, float eps, cudaStream_t s) {
  launch_pdl(dsq_combine_norm_kernel, dim3((unsigned)rows), dim3(256), (size_t)H * 4, s, base, ymoe, wmoe, topk, add1, w, out, H, eps);
  launch_check("dsq_combine_norm");
}

int dsq_attn_splits(int smax) { return std::max(1, std::min(32, (smax + 127) / 128)); }
size_t dsq_attn_ws_floats(long long rows, int heads, int nsplit) { return (size_t)rows * heads * nsplit * kPartStride; }

void dsq_attn_split(const float* qkv, const float* cos_t, const float* sin_t, void* kc, void* vc, bool kv_f16,
                    const int* row_page, const int* row_pos, float* part, int* counters, float* ctx, long long rows,
                    int heads, int head_dim, int smax, float scale, int nsplit, cudaStream_t s) {
  if (head_dim != 128) throw std::runtime_error("dsq_attn_split: head_dim must be 128");
  const dim3 grid((unsigned)(rows * heads), (unsigned)nsplit);
  if (kv_f16)
    launch_pdl(dsq_attn_split_kernel<__half>, grid, dim3(128), 0, s, qkv, cos_t, sin_t, (__half*)kc, (__half*)vc, row_page, row_pos, part, counters, ctx, heads, smax, scale, nsplit);
  else
    launch_pdl(dsq_attn_split_kernel<float>, grid, dim3(128), 0, s, qkv, cos_t, sin_t, (float*)kc, (float*)vc, row_page, row_pos, part, counters, ctx, heads, smax, scale, nsplit);
  launch_check("dsq_attn_split");
}

}  // namespace dsocr
