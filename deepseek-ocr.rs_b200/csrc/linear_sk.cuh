// Decode-time MoE expert GEMM: fixed-capacity expert segments, scheduled on the device, stream-K balanced.
//
// Expert g owns weight rows [g*w_rows, +w_rows) and token rows [g*cap, +counts[g]) of X / out; counts[] is produced
// on the device by the router kernel of the same step.  Every CTA derives the same compact list of non-empty
// (expert, token chunk, 128-row weight block) units from counts[].  The k-blocks of all units form one sequence that
// is dealt out in equal contiguous ranges to the CTAs, so every SM streams the same number of weight bytes however
// few experts are populated and however the unit count divides by the SM count.  A unit whose reduction is cut by a
// range boundary is finished by the CTA that owns its first k-block: the CTAs that own the rest store their raw f32
// accumulators to a per-CTA workspace slot (they do that first thing, the finisher needs them last thing, so nobody
// waits in practice) and the finisher adds them in CTA order -> bit-reproducible.
//
// Same operand convention as linear_tc.cuh (weights = MMA A / TMEM lanes, tokens = MMA B / TMEM columns, hi+lo split
// activations, NA = 2 fuses SwiGLU).  Replaces the per-expert matmul loop of run_moe (transformer/block.rs:1303-1395).
#pragma once
#include "linear_tc.cuh"

namespace lin {

struct SkParams {
  int K;                   // reduction (multiple of 64)
  int x_lo_row_off;        // row offset of the lo activation part inside the X tensor map
  void* out;               // [groups*cap, ldo]
  void* out_lo;            // OUT_T_SPLIT only
  long long ldo;
  int out_mode;            // OUT_T_SPLIT (hi + lo 16-bit) or OUT_F32; NA == 2 stores silu(gate) * up
  const int* counts;       // [groups] rows per group (device)
  int groups;
  int wpg;                 // 128-row weight blocks per group
  int cap;                 // token rows reserved per group
  int w_rows;              // weight rows per group
  const uint8_t* w0_tiled; // pre-tiled weights (see Params::w0_tiled) or nullptr
  const uint8_t* w1_tiled;
  float* ws;               // stream-K partial slots, one per CTA: [NA][BN][128] f32
  int* flags;              // per CTA {arrivals of its partial (8 = complete), consumers done}; 0 between launches
};

template <typename T, int BN, int NA>
__global__ void __launch_bounds__(kThreads, 1)
linear_sk_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_x16, const SkParams p) {
  constexpr int NB = 2;
  using C = Cfg<BN, NA, NB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.K / BK;
  const int G = gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_w0);
    if (NA == 2) ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_x16);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::kTmemCols);

  constexpr int kMaxSeg = 64;
  constexpr int kMaxGroups = 256;
  __shared__ int s_seg[kMaxSeg][8];  // w_row0, x_row0, rows, n0, kb0, kb1, contributors {first, last} (finisher only)
  __shared__ int s_prefix[kMaxGroups + 1];
  __shared__ int s_nseg;
  if (warp == 2) {  // exclusive prefix of units per group
    int carry = 0;
    for (int base = 0; base < p.groups; base += 32) {
      const int g = base + lane;
      int u = 0;
      if (g < p.groups) u = ((p.counts[g] + BN - 1) / BN) * p.wpg;
      int inc = u;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (g < p.groups) s_prefix[g] = carry + inc - u;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_prefix[p.groups] = carry;
  }
  __syncthreads();
  {
    const long long total = (long long)s_prefix[p.groups] * num_kb;  // k-blocks of all units
    const int Ge = (int)min((long long)G, total);                    // CTAs that get a (non-empty) range
    auto start = [&](int b) -> int { return (int)((total * b) / Ge); };
    const bool active = (int)blockIdx.x < Ge;
    const int r0 = active ? start(blockIdx.x) : 0, r1 = active ? start(blockIdx.x + 1) : 0;
    const int u_first = r0 / num_kb;
    const int nseg = r1 > r0 ? (r1 - 1) / num_kb - u_first + 1 : 0;
    if (nseg > kMaxSeg) {
      if (threadIdx.x == 0) printf("linear_sk_kernel: %d work items exceed the per-CTA list\n", nseg);
      __trap();
    }
    if (threadIdx.x == 0) s_nseg = nseg;
    for (int i = threadIdx.x; i < nseg; i += kThreads) {
      const int u = u_first + i;
      int lo = 0, hi = p.groups;  // s_prefix[lo] <= u < s_prefix[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_prefix[mid] <= u) lo = mid; else hi = mid;
      }
      const int local = u - s_prefix[lo];
      const int ch = local / p.wpg, wb = local - ch * p.wpg;
      s_seg[i][0] = lo * p.w_rows + wb * BM;
      s_seg[i][1] = lo * p.cap + ch * BN;
      s_seg[i][2] = min(BN, p.counts[lo] - ch * BN);
      s_seg[i][3] = wb * BM;
      const int kb0 = max(r0 - u * num_kb, 0), kb1 = min(r1 - u * num_kb, num_kb);
      s_seg[i][4] = kb0;
      s_seg[i][5] = kb1;
      // finisher (owns k-block 0 but not the last one): the CTAs whose ranges start inside the rest of this unit
      int c_first = 0, c_last = -1;
      if (kb0 == 0 && kb1 < num_kb) {
        const int unit_end = (u + 1) * num_kb;
        c_first = blockIdx.x + 1;
        c_last = blockIdx.x;
        for (int b = c_first; b < Ge && start(b) < unit_end; ++b) c_last = b;
      }
      s_seg[i][6] = c_first;
      s_seg[i][7] = c_last;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nseg = s_nseg;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nseg; ++i) {
        const int w_row0 = s_seg[i][0], x_row0 = s_seg[i][1], rows = s_seg[i][2];
        const int kb0 = s_seg[i][4], kb1 = s_seg[i][5];
        const int nbox = (rows + 15) >> 4;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::kStageBytes;
          ptx::mbar_expect_tx(&full[stage], NA * C::kABytes + NB * nbox * 2048);
          if (p.w0_tiled) {
            const size_t off = ((size_t)(w_row0 / BM) * num_kb + kb) * C::kABytes;
            ptx::bulk_load(st, p.w0_tiled + off, C::kABytes, &full[stage]);
            if (NA == 2) ptx::bulk_load(st + C::kABytes, p.w1_tiled + off, C::kABytes, &full[stage]);
          } else {
            ptx::tma_load_2d(st, &tm_w0, &full[stage], kb * BK, w_row0);
            if (NA == 2) ptx::tma_load_2d(st + C::kABytes, &tm_w1, &full[stage], kb * BK, w_row0);
          }
          for (int b = 0; b < nbox; ++b) {
            ptx::tma_load_2d(st + NA * C::kABytes + b * 2048, &tm_x16, &full[stage], kb * BK, x_row0 + b * 16);
            ptx::tma_load_2d(st + NA * C::kABytes + C::kBBytes + b * 2048, &tm_x16, &full[stage], kb * BK,
                             p.x_lo_row_off + x_row0 + b * 16);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::idesc_f16(Elem<T>::kFmt, BM, BN);
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nseg; ++i) {
        const int kb0 = s_seg[i][4], kb1 = s_seg[i][5];
        const int buf = i & 1;
        const uint32_t bphase = (i >> 1) & 1;
        ptx::mbar_wait(&tempty[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * C::kAccCols;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + NA * C::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              const uint64_t ad = ptx::smem_desc_sw128(sa + a * C::kABytes + k * 32, 16, 1024);
#pragma unroll
              for (int b = 0; b < NB; ++b) {
                const uint64_t bd = ptx::smem_desc_sw128(sb + b * C::kBBytes + k * 32, 16, 1024);
                ptx::mma_f16_ss(d0 + a * BN, ad, bd, idesc, ((kb - kb0) | k | b) ? 1u : 0u);
              }
            }
          }
          ptx::mma_commit(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;              // which half of the token columns this warp handles
    constexpr int kChunks = BN / 32;
    constexpr int kChunksPerHalf = (kChunks + 1) / 2;
    constexpr int kSlotFloats = NA * BN * BM;
    for (int i = 0; i < nseg; ++i) {
      const int x_row0 = s_seg[i][1], rows = s_seg[i][2], n0 = s_seg[i][3];
      const bool contributor = s_seg[i][4] > 0;
      const int c_first = s_seg[i][6], c_last = s_seg[i][7];
      const int buf = i & 1;
      const uint32_t bphase = (i >> 1) & 1;
      ptx::mbar_wait(&tfull[buf], bphase);
      ptx::tc_fence_after();
      const int n = n0 + quarter * 32 + lane;  // output feature owned by this thread
      const uint32_t trow = tmem_base + buf * C::kAccCols + ((uint32_t)(quarter * 32) << 16);
      if (c_last >= c_first && lane == 0) {  // finisher: the partials were stored long ago, this does not spin
        for (int b = c_first; b <= c_last; ++b) {
          int spins = 0;
          while (ptx::ld_acquire(p.flags + 2 * b) < kEpiWarps) {
            __nanosleep(32);
            if (++spins > (1 << 22)) { printf("linear_sk_kernel: partial of CTA %d never arrived\n", b); __trap(); }
          }
        }
      }
      __syncwarp();
      for (int c = half * kChunksPerHalf; c < min(kChunks, (half + 1) * kChunksPerHalf); ++c) {
        if (c * 32 >= rows) break;  // warp-uniform
        uint32_t v[32];
        ptx::tmem_ld_32x32(trow + c * 32, v);
        uint32_t u[32];
        if (NA == 2) ptx::tmem_ld_32x32(trow + BN + c * 32, u);
        ptx::tmem_ld_wait();
        const int nvalid = min(32, rows - c * 32);
        if (contributor) {
          float* slot = p.ws + (size_t)blockIdx.x * kSlotFloats + (size_t)(c * 32) * BM + quarter * 32 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              __stcg(slot + j * BM, __uint_as_float(v[j]));
              if (NA == 2) __stcg(slot + (BN + j) * BM, __uint_as_float(u[j]));
            }
          }
          continue;
        }
        for (int b = c_first; b <= c_last; ++b) {  // fixed order -> deterministic
          const float* slot = p.ws + (size_t)b * kSlotFloats + (size_t)(c * 32) * BM + quarter * 32 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldcg(slot + j * BM));
              if (NA == 2) u[j] = __float_as_uint(__uint_as_float(u[j]) + __ldcg(slot + (BN + j) * BM));
            }
          }
        }
        const long long base = (long long)(x_row0 + c * 32) * p.ldo + n;
        if (p.out_mode == OUT_T_SPLIT) {
          T* ptr = reinterpret_cast<T*>(p.out) + base;
          T* ptr_lo = reinterpret_cast<T*>(p.out_lo) + base;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              const float r = NA == 2 ? silu(__uint_as_float(v[j])) * __uint_as_float(u[j]) : __uint_as_float(v[j]);
              const T hi = Elem<T>::from(r);
              ptr[j * p.ldo] = hi;
              ptr_lo[j * p.ldo] = Elem<T>::from(r - Elem<T>::to(hi));
            }
          }
        } else {
          float* ptr = reinterpret_cast<float*>(p.out) + base;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid)
              ptr[j * p.ldo] = NA == 2 ? silu(__uint_as_float(v[j])) * __uint_as_float(u[j]) : __uint_as_float(v[j]);
          }
        }
      }
      ptx::tc_fence_before();
      if (contributor) __threadfence();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&tempty[buf]);
        if (contributor) atomicAdd(p.flags + 2 * blockIdx.x, 1);
        // hand the contributors' flags back (zeroed) for the next launch once all epilogue warps have consumed them
        for (int b = c_first; b <= c_last; ++b) {
          if (atomicAdd(p.flags + 2 * b + 1, 1) == kEpiWarps - 1) {
            p.flags[2 * b + 1] = 0;
            p.flags[2 * b] = 0;
          }
        }
      }
    }
  }

  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace lin
