// Causal attention of the decoder PREFILL on tcgen05 / TMEM, head_dim 128 (attention_forward, block.rs:446-804 with the
// additive causal bias of build_attention_bias :1504-1561), over the per-page KV cache the RoPE kernel has just filled.
//
// One CTA = 128 consecutive query positions of one (page, head).  It walks the keys 0 .. last query of the tile in
// blocks of 64; blocks entirely above the diagonal are never visited, the diagonal ones are masked in the softmax.
// Precision model of the decoder (DESIGN.md "Precision model"): every tensor-core operand is an exact or hi+lo split
// 16-bit value and accumulation is f32, so the result matches the f32 oracle to ~1e-5 instead of one 16-bit rounding:
//   Q (f32 from the RoPE kernel)          -> f16 hi + lo
//   K, V from an f16 cache                -> used as they are (exact);  from an f32 cache -> f16 hi + lo
//   P = 2^(s - m) (f32 in registers)      -> f16 hi + lo
//   S = Qhi.Khi + Qlo.Khi (+ Qhi.Klo),  O += Phi.Vhi + Plo.Vhi (+ Phi.Vlo)     (the lo.lo terms are below f32 rounding)
// Operands are staged by four loader warps with plain loads (the cache rows of a (page, head) are contiguous) that
// convert and write the 128-byte-swizzled K-major / MN-major tiles the MMA descriptors expect, then publish them with
// fence.proxy.async + an mbarrier; one thread issues the MMAs; four softmax warps own one query row each (online
// softmax in the log2 domain with the lazy 2^8 rescale of attention_tc.cuh).  The output leaves as the hi/lo split
// 16-bit context rows the o_proj GEMM consumes.  Two K/V stages, one S and one P buffer, O resident in TMEM.
#pragma once
#include "ptx.cuh"

namespace pattn {

constexpr int BQ = 128;   // queries per CTA
constexpr int KV = 64;    // keys per block
constexpr int D = 128;    // head dim
constexpr int kThreads = 288;  // warp 0: MMA issue + TMEM; warps 1-4: loaders; warps 5-8: softmax

template <int NKV>  // 16-bit parts per cached K / V value: 1 (f16 cache) or 2 (f32 cache)
struct Cfg {
  static constexpr int kQPart = BQ * D * 2;            // 32 KB: [2 k-atoms][128 rows][128 B]
  static constexpr int kKPart = KV * D * 2;            // 16 KB: [2 atoms][64 rows][128 B]
  static constexpr int kStage = 2 * NKV * kKPart;      // K parts then V parts
  static constexpr int kPPart = BQ * KV * 2;           // 16 KB: [128 rows][128 B]
  static constexpr int kOffQ = 0;
  static constexpr int kOffStage = 2 * kQPart;
  static constexpr int kOffP = kOffStage + 2 * kStage;
  static constexpr int kOffBar = kOffP + 2 * kPPart;
  static constexpr int kSmemBytes = kOffBar + 128;
  static_assert(kSmemBytes <= 232448, "shared memory");
};

__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
  __align__(16) __half2 h[4];
  __align__(16) __half2 l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    const float2 back = __half22float2(h[i]);
    l[i] = __floats2half2_rn(f[2 * i] - back.x, f[2 * i + 1] - back.y);
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

// one 64-dim half row (one 128-byte swizzle-atom row) of a cached K or V tile -> hi (and lo) parts in shared memory
template <int NKV>
__device__ __forceinline__ void stage_half_row(const __half* g, bool valid, uint8_t* hi_row, uint8_t* lo_row, int rsw) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (valid) v = reinterpret_cast<const uint4*>(g)[c];
    *reinterpret_cast<uint4*>(hi_row + ((c ^ rsw) << 4)) = v;
  }
}
template <int NKV>
__device__ __forceinline__ void stage_half_row(const float* g, bool valid, uint8_t* hi_row, uint8_t* lo_row, int rsw) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float4 a = reinterpret_cast<const float4*>(g)[2 * c], b = reinterpret_cast<const float4*>(g)[2 * c + 1];
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    uint4 hi, lo;
    split8(f, hi, lo);
    *reinterpret_cast<uint4*>(hi_row + ((c ^ rsw) << 4)) = hi;
    *reinterpret_cast<uint4*>(lo_row + ((c ^ rsw) << 4)) = lo;
  }
}

template <typename T, typename TKV>
__global__ void __launch_bounds__(kThreads, 1)
pattn_kernel(const float* __restrict__ q, const TKV* __restrict__ kc, const TKV* __restrict__ vc,
             const int* __restrict__ page_row0, const int* __restrict__ page_len, T* __restrict__ ctx,
             long long lo_off_elems, int heads, int smax, float scale_log2) {
  constexpr int NKV = sizeof(TKV) == 2 ? 1 : 2;
  using C = Cfg<NKV>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int page = blockIdx.y, hd = blockIdx.z;
  const int len = page_len[page];
  const int q0 = blockIdx.x * BQ;
  if (q0 >= len) return;  // whole CTA, before any barrier / TMEM state exists
  const int nq = min(BQ, len - q0);
  const int kend = q0 + nq;                     // keys 0 .. kend-1 are visible to at least one row of the tile
  const int nblk = (kend + KV - 1) / KV;
  const long long r0 = (long long)page_row0[page] + q0;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars;        // 128 loader arrivals
  uint64_t* k_full = bars + 1;    // [2] 128 loader arrivals
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2] MMA commit
  uint64_t* s_full = bars + 7;    // MMA commit
  uint64_t* p_full = bars + 8;    // 128 softmax arrivals
  uint64_t* o_full = bars + 9;    // MMA commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) { printf("pattn: smem base not 1024-aligned\n"); __trap(); }
    ptx::mbar_init(q_full, 128);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&k_full[i], 128); ptx::mbar_init(&v_full[i], 128); ptx::mbar_init(&kv_empty[i], 1); }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem;        // S: columns [0, 64)
  const uint32_t tmem_o = tmem + 128;  // O: columns [128, 256), two 64-column halves

  if (warp == 0) {
    // ------------------------------------------------------------------ MMA issue
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_qk = ptx::idesc_f16(0, BQ, KV);        // f16 x f16, A and B K-major
      constexpr uint32_t idesc_pv = ptx::idesc_f16(0, BQ, 64, 0, 1);  // B (= V) MN-major, one 64-dim half per MMA
      const uint32_t sq = ptx::smem_u32(smem + C::kOffQ);
      const uint32_t sp = ptx::smem_u32(smem + C::kOffP);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait(&k_full[s], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t sk = ptx::smem_u32(smem + C::kOffStage + s * C::kStage);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t aoff = (kk >> 2) * (BQ * 128) + (kk & 3) * 32;
          const uint32_t boff = (kk >> 2) * (KV * 128) + (kk & 3) * 32;
          const uint64_t a_hi = ptx::smem_desc_sw128(sq + aoff, 16, 1024), a_lo = ptx::smem_desc_sw128(sq + C::kQPart + aoff, 16, 1024);
          const uint64_t b_hi = ptx::smem_desc_sw128(sk + boff, 16, 1024);
          ptx::mma_f16_ss(tmem_s, a_hi, b_hi, idesc_qk, kk ? 1u : 0u);
          ptx::mma_f16_ss(tmem_s, a_lo, b_hi, idesc_qk, 1u);
          if (NKV == 2) ptx::mma_f16_ss(tmem_s, a_hi, ptx::smem_desc_sw128(sk + C::kKPart + boff, 16, 1024), idesc_qk, 1u);
        }
        ptx::mma_commit(s_full);
      };
      ptx::mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(p_full, j & 1);  // P_j in smem, S_j fully read
        ptx::mbar_wait(&v_full[s], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t sv = ptx::smem_u32(smem + C::kOffStage + s * C::kStage + NKV * C::kKPart);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int kk = 0; kk < KV / 16; ++kk) {
            const uint64_t p_hi = ptx::smem_desc_sw128(sp + kk * 32, 16, 1024), p_lo = ptx::smem_desc_sw128(sp + C::kPPart + kk * 32, 16, 1024);
            const uint32_t voff = half * (KV * 128) + kk * 2048;  // 16 key rows down inside the 64-dim atom `half`
            const uint64_t v_hi = ptx::smem_desc_sw128(sv + voff, KV * 128, 1024);
            ptx::mma_f16_ss(tmem_o + half * 64, p_hi, v_hi, idesc_pv, (j | kk) ? 1u : 0u);
            ptx::mma_f16_ss(tmem_o + half * 64, p_lo, v_hi, idesc_pv, 1u);
            if (NKV == 2) ptx::mma_f16_ss(tmem_o + half * 64, p_hi, ptx::smem_desc_sw128(sv + C::kKPart + voff, KV * 128, 1024), idesc_pv, 1u);
          }
        }
        ptx::mma_commit(o_full);
        ptx::mma_commit(&kv_empty[s]);
        if (j + 1 < nblk) issue_qk(j + 1);
      }
    }
  } else if (warp <= 4) {
    // ------------------------------------------------------------------ loaders: Q once, then the K / V blocks
    const int lt = threadIdx.x - 32;  // 0..127
    {
      const int row = lt, rsw = row & 7;
      const bool valid = row < nq;
      const float* src = q + ((r0 + row) * heads + hd) * D;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        uint8_t* hi_row = smem + C::kOffQ + a * (BQ * 128) + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (valid) {
            const float4 x = reinterpret_cast<const float4*>(src + a * 64)[2 * c], y = reinterpret_cast<const float4*>(src + a * 64)[2 * c + 1];
            f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w; f[4] = y.x; f[5] = y.y; f[6] = y.z; f[7] = y.w;
          }
          uint4 hi, lo;
          split8(f, hi, lo);
          *reinterpret_cast<uint4*>(hi_row + ((c ^ rsw) << 4)) = hi;
          *reinterpret_cast<uint4*>(hi_row + C::kQPart + ((c ^ rsw) << 4)) = lo;
        }
      }
      ptx::fence_proxy_async();
      ptx::mbar_arrive(q_full);
    }
    const int key = lt >> 1, a = lt & 1, rsw = key & 7;
    const TKV* kbase = kc + ((long long)page * heads + hd) * smax * D + a * 64;
    const TKV* vbase = vc + ((long long)page * heads + hd) * smax * D + a * 64;
    for (int j = 0; j < nblk; ++j) {
      const int s = j & 1;
      ptx::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      const int kpos = j * KV + key;
      // rows past the tile's last position were never written by this call: zero them (stale bytes may decode to NaN)
      const bool valid = kpos < kend;
      uint8_t* st = smem + C::kOffStage + s * C::kStage;
      uint8_t* krow = st + a * (KV * 128) + key * 128;
      stage_half_row<NKV>(kbase + (long long)kpos * D, valid, krow, krow + C::kKPart, rsw);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&k_full[s]);
      uint8_t* vrow = st + NKV * C::kKPart + a * (KV * 128) + key * 128;
      stage_half_row<NKV>(vbase + (long long)kpos * D, valid, vrow, vrow + C::kKPart, rsw);
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&v_full[s]);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps (one query row per thread)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int qpos = q0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = smem + C::kOffP + r * 128;
    const int rsw = r & 7;
    constexpr float kRescaleThreshold = 8.0f;
    for (int j = 0; j < nblk; ++j) {
      const bool diag = (j + 1) * KV - 1 > q0;  // some (row, key) pairs of this block are above the diagonal
      const int kmax = qpos - j * KV;           // columns c <= kmax are visible to this row
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after();
      float bmax = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < KV; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x16(tmem_s + lane_off + c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float t = __uint_as_float(v[i]) * scale_log2;
          if (diag && c0 + i > kmax) t = -INFINITY;
          bmax = fmaxf(bmax, t);
        }
      }
      const bool need = bmax > m + kRescaleThreshold;
      float alpha = 1.f;
      if (need) {
        alpha = ptx::ex2_approx(m - bmax);  // first block: ex2(-inf) = 0
        l *= alpha;
        m = bmax;
      }
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        ptx::mbar_wait(o_full, (j - 1) & 1);  // P.V of the previous block has landed in O
        ptx::tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(tmem_o + lane_off + c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          ptx::tmem_st_32x16(tmem_o + lane_off + c0, v);
        }
        ptx::tmem_st_wait();
      }
      float rowsum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < KV; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x16(tmem_s + lane_off + c0, v);
        ptx::tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pe = ptx::ex2_approx(__uint_as_float(v[i]) * scale_log2 - m);
          if (diag && c0 + i > kmax) pe = 0.f;
          e[i] = pe;
          rowsum += pe;
        }
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint4 hi, lo;
          split8(e + g8 * 8, hi, lo);
          const int chunk = (((c0 >> 3) + g8) ^ rsw) << 4;
          *reinterpret_cast<uint4*>(prow + chunk) = hi;
          *reinterpret_cast<uint4*>(prow + C::kPPart + chunk) = lo;
        }
      }
      l += rowsum;
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }
    // epilogue: O / l as hi/lo split 16-bit context rows
    ptx::mbar_wait(o_full, (nblk - 1) & 1);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    const bool q_ok = r < nq;
    T* orow = ctx + ((r0 + r) * heads + hd) * D;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x16(tmem_o + lane_off + c0, v);
      ptx::tmem_ld_wait();
      if (q_ok) {
        __align__(16) T hi[16];
        __align__(16) T lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float o = __uint_as_float(v[i]) * inv;
          hi[i] = Elem<T>::from(o);
          lo[i] = Elem<T>::from(o - Elem<T>::to(hi[i]));
        }
        reinterpret_cast<uint4*>(orow + c0)[0] = reinterpret_cast<uint4*>(hi)[0];
        reinterpret_cast<uint4*>(orow + c0)[1] = reinterpret_cast<uint4*>(hi)[1];
        reinterpret_cast<uint4*>(orow + lo_off_elems + c0)[0] = reinterpret_cast<uint4*>(lo)[0];
        reinterpret_cast<uint4*>(orow + lo_off_elems + c0)[1] = reinterpret_cast<uint4*>(lo)[1];
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace pattn
