// Dequant-fused tensor-core linear layer for DSQ snapshots (run_quantized_matmul, quantization.rs:164-185;
// QMatMul::forward serves any M, dsq-runtime/src/lib.rs:316-369):
//     out[m, n] = epi( sum_k X[m, k] * dequant(W)[n, k] )        W = Q8_0 / Q4_K / Q6_K blocks (or f32 fallback rows)
// Same persistent, warp-specialised tcgen05 kernel as linear_tc.cuh (weights = MMA-M operand, tokens = N operand, TMEM
// double-buffered accumulators, SwiGLU / split-K / grouped tiles), except that the weight stage is not copied by TMA:
// four producer warps - one thread per weight row of the 128 x 64 stage - read the quantised bytes of their row's
// 64-wide k-block, dequantise them in registers (dsq_dequant64, csrc/dsq_dequant.h, the routine the CPU tests pin to
// gguf-py), split every weight into a hi + lo pair of 16-bit values and write both 128-byte rows into the 128B-swizzled
// stage (chunk ^ (row & 7)); fence.proxy.async + an arrive on the stage's full barrier (count 1 + 128) publishes them to
// the tensor core.  With hi/lo activations the accumulator receives Whi.Xhi + Wlo.Xhi + Whi.Xlo (the lo.lo term is below
// f32 rounding), so the result matches the f32-dequant oracle to ~1e-5 and the quantised bytes are read from HBM once
// per token TILE instead of once per token pair (the per-row GEMV this replaces for prefill and multi-page decode).
#pragma once
#include "dsq_dequant.h"
#include "linear_tc.cuh"

namespace lin {

struct DqWeights {
  dsocr::DsqPlanes w0, w1;  // device planes of QuantWeight (dsq.h); w1 only for the dual (SwiGLU) kernel
  int fmt0, fmt1;           // 8 / 12 / 14 / 0
  long long w_rows;         // rows of the (stacked) weight matrix: rows past it are zero
};

constexpr int kDqWarps = 4;
constexpr int kDqThreads = kThreads + 32 * kDqWarps;  // + the dequant producer warps

template <int BN, int NA>
struct DqCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = 2 * NA * kABytes + 2 * kBBytes;  // (hi, lo) per weight, (hi, lo) activations
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 6 ? 6 : (200 * 1024 / kStageBytes);
  static constexpr int kAccCols = NA * BN;
  static constexpr int kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64
                                   : (2 * kAccCols <= 128) ? 128 : (2 * kAccCols <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "stages");
  static_assert(2 * kAccCols <= 512, "TMEM overflow");
};

// 8 consecutive weights (one 16-byte chunk of the 128-byte stage row) -> hi / lo 16-bit parts at the swizzled position
template <typename T>
__device__ __forceinline__ void dq_store_chunk(const float (&w)[8], uint8_t* hi_row, uint8_t* lo_row, int c, int rsw) {
  __align__(16) T hi[8];
  __align__(16) T lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    hi[i] = Elem<T>::from(w[i]);
    lo[i] = Elem<T>::from(w[i] - Elem<T>::to(hi[i]));
  }
  *reinterpret_cast<uint4*>(hi_row + ((c ^ rsw) << 4)) = *reinterpret_cast<uint4*>(hi);
  *reinterpret_cast<uint4*>(lo_row + ((c ^ rsw) << 4)) = *reinterpret_cast<uint4*>(lo);
}

// One 64-wide k-block of one weight row, straight from the quantised planes into the stage: the same arithmetic as
// dsq_dequant64 (csrc/dsq_dequant.h, the CPU-pinned routine; both are held to the oracle by tests/test_linear_dq_gpu.py)
// with 128-bit loads and no intermediate array in local memory.
template <typename T>
__device__ __forceinline__ void dq_stage_row(int fmt, const dsocr::DsqPlanes& p, long long row, int K, int kb, uint8_t* hi_row,
                                             uint8_t* lo_row, int rsw) {
  const int k0 = kb * 64;
  if (fmt == 8) {  // Q8_0: int8 qs [rows][K], f16 d [rows][K/32]
    const int4* q = reinterpret_cast<const int4*>(p.a + row * K + k0);
    const __half* dp = reinterpret_cast<const __half*>(p.b) + row * (K / 32) + (k0 >> 5);
    const float d0 = __half2float(dp[0]), d1 = __half2float(dp[1]);
    int4 v[4] = {q[0], q[1], q[2], q[3]};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int4 t = v[c >> 1];
      const int w0 = (c & 1) ? t.z : t.x, w1 = (c & 1) ? t.w : t.y;
      const float d = c < 4 ? d0 : d1;
      float w[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        w[i] = d * (float)(int)(signed char)(w0 >> (8 * i));
        w[4 + i] = d * (float)(int)(signed char)(w1 >> (8 * i));
      }
      dq_store_chunk<T>(w, hi_row, lo_row, c, rsw);
    }
  } else if (fmt == 12) {  // Q4_K: 144-byte super-blocks as on disk
    const int sb = k0 >> 8, g = (k0 >> 6) & 3;
    const uint8_t* blk = p.a + (row * (K / 256) + sb) * 144;
    const uint4 head = *reinterpret_cast<const uint4*>(blk);  // d, dmin, scales[12]
    const float d = __half2float(__ushort_as_half((unsigned short)(head.x & 0xFFFF)));
    const float dmin = __half2float(__ushort_as_half((unsigned short)(head.x >> 16)));
    const uint32_t sw[3] = {head.y, head.z, head.w};
    auto sbyte = [&](int i) { return (int)((sw[i >> 2] >> (8 * (i & 3))) & 0xFF); };
    int sc[2], mn[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // ggml get_scale_min_k4
      const int j = 2 * g + h;
      if (j < 4) { sc[h] = sbyte(j) & 63; mn[h] = sbyte(j + 4) & 63; }
      else { sc[h] = (sbyte(j + 4) & 0xF) | ((sbyte(j - 4) >> 6) << 4); mn[h] = (sbyte(j + 4) >> 4) | ((sbyte(j) >> 6) << 4); }
    }
    const float d1 = d * (float)sc[0], m1 = dmin * (float)mn[0], d2 = d * (float)sc[1], m2 = dmin * (float)mn[1];
    const uint4* qp = reinterpret_cast<const uint4*>(blk + 16 + g * 32);
    const uint4 qa = qp[0], qb = qp[1];
    const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int cc = c & 3;  // bytes [8cc, 8cc + 8): low nibbles -> weights 8cc.., high nibbles -> weights 32 + 8cc..
      const uint32_t w0 = qw[2 * cc], w1 = qw[2 * cc + 1];
      const int sh = c < 4 ? 0 : 4;
      const float dd = c < 4 ? d1 : d2, mm = c < 4 ? m1 : m2;
      float w[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        w[i] = dd * (float)((w0 >> (8 * i + sh)) & 0xF) - mm;
        w[4 + i] = dd * (float)((w1 >> (8 * i + sh)) & 0xF) - mm;
      }
      dq_store_chunk<T>(w, hi_row, lo_row, c, rsw);
    }
  } else if (fmt == 14) {  // Q6_K: ql [rows][K/2], qh [rows][K/4], int8 scales [rows][K/16], f16 d [rows][K/256]
    const int sb = k0 >> 8, within = k0 & 255, half = within >> 7, upper = (within >> 6) & 1;
    const uint4* qlp = reinterpret_cast<const uint4*>(p.a + row * (K / 2) + sb * 128 + half * 64);
    const uint4* qhp = reinterpret_cast<const uint4*>(p.b + row * (K / 4) + sb * 64 + half * 32);
    const uint2 scw = *reinterpret_cast<const uint2*>(p.c + row * (K / 16) + sb * 16 + half * 8);
    const float d = __half2float(reinterpret_cast<const __half*>(p.d)[row * (K / 256) + sb]);
    const uint4 l0 = qlp[0], l1 = qlp[1], l2 = qlp[2], l3 = qlp[3], h0 = qhp[0], h1 = qhp[1];
    const uint32_t ql[16] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w, l2.x, l2.y, l2.z, l2.w, l3.x, l3.y, l3.z, l3.w};
    const uint32_t qh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    auto scale = [&](int i) { const uint32_t wv = i < 4 ? scw.x : scw.y; return (float)(int)(signed char)(wv >> (8 * (i & 3))); };
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      // chunk c < 4: weights l = 8c.. (from ql[l], qh[l] bits 0-1 / 4-5); c >= 4: weights 32 + l, l = 8(c-4).. (ql[l+32], bits 2-3 / 6-7)
      const int l0i = (c & 3) * 8;
      const int is = l0i >> 4;
      const int sidx = is + (c < 4 ? 0 : 2) + (upper ? 4 : 0);
      const float ds = d * scale(sidx);
      const int lsh = upper ? 4 : 0;
      const int hsh = (c < 4 ? 0 : 2) + (upper ? 4 : 0);
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int l = l0i + i;
        const int lb = l + (c < 4 ? 0 : 32);
        const int lo4 = (int)((ql[lb >> 2] >> (8 * (lb & 3) + lsh)) & 0xF);
        const int hi2 = (int)((qh[l >> 2] >> (8 * (l & 3) + hsh)) & 3);
        w[i] = ds * (float)((lo4 | (hi2 << 4)) - 32);
      }
      dq_store_chunk<T>(w, hi_row, lo_row, c, rsw);
    }
  } else {  // f32 rows (float fallback of the exporter's chain)
    const float4* wp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.a) + row * K + k0);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 a = wp[2 * c], b = wp[2 * c + 1];
      const float w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      dq_store_chunk<T>(w, hi_row, lo_row, c, rsw);
    }
  }
}

template <typename T, int BN, int NA>
__global__ void __launch_bounds__(kDqThreads, 1)
linear_dq_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x16, const Params p,
                 const DqWeights q) {
  constexpr int NB = 2;
  using C = DqCfg<BN, NA>;
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
  int num_tiles = p.num_tiles_dev ? min(*p.num_tiles_dev, p.num_tiles) : p.num_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_x16);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full[s], 1 + 32 * kDqWarps);  // the TMA thread's expect_tx arrive + every dequant thread
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::kTmemCols);
  // grouped problems: stage this CTA's tile descriptors (with the row counts resolved) in shared memory so that
  // the three roles do not each pay dependent global loads at every tile boundary
  constexpr int kTileCache = 64;
  constexpr int kMaxGroups = 256;
  __shared__ int s_tiles[kTileCache][5];  // w_row0, x_row0, rows, n0 (+pad)
  __shared__ int s_prefix[kMaxGroups + 1];
  if (p.dyn_groups) {
    if (warp == 2) {  // exclusive prefix of units per group
      int carry = 0;
      for (int base = 0; base < p.dyn_groups; base += 32) {
        const int g = base + lane;
        int u = 0;
        if (g < p.dyn_groups) u = ((p.group_counts[g] + BN - 1) / BN) * p.dyn_wpg;
        int inc = u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        if (g < p.dyn_groups) s_prefix[g] = carry + inc - u;
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) s_prefix[p.dyn_groups] = carry;
    }
    __syncthreads();
    num_tiles = s_prefix[p.dyn_groups];
    if (num_tiles > kTileCache * (int)gridDim.x) {
      if (threadIdx.x == 0) printf("linear_kernel: %d grouped units exceed the per-CTA tile cache\n", num_tiles);
      __trap();
    }
    for (int i = threadIdx.x; i < kTileCache; i += kDqThreads) {
      const int u = blockIdx.x + i * gridDim.x;
      if (u < num_tiles) {
        int lo = 0, hi = p.dyn_groups;  // s_prefix[lo] <= u < s_prefix[hi]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_prefix[mid] <= u) lo = mid; else hi = mid;
        }
        const int local = u - s_prefix[lo];
        const int ch = local / p.dyn_wpg, wb = local - ch * p.dyn_wpg;
        s_tiles[i][0] = lo * p.dyn_w_rows + wb * BM;
        s_tiles[i][1] = lo * p.dyn_cap + ch * BN;
        s_tiles[i][2] = min(BN, p.group_counts[lo] - ch * BN);
        s_tiles[i][3] = wb * BM;
      }
    }
  } else if (p.tiles) {
    for (int i = threadIdx.x; i < kTileCache; i += kDqThreads) {
      const int t = blockIdx.x + i * gridDim.x;
      if (t < num_tiles) {
        const Tile tl = p.tiles[t];
        int rows = tl.rows;
        if (p.group_counts) rows = max(0, min(BN, p.group_counts[tl.group] - tl.r0));
        s_tiles[i][0] = tl.w_row0; s_tiles[i][1] = tl.x_row0; s_tiles[i][2] = rows; s_tiles[i][3] = tl.n0;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_batch = p.n_w_blocks * ((p.M + BN - 1) / BN);
  const int base_tiles = tiles_per_batch * p.nbatch;
  auto decode_tile = [&](int t, int& w_row0, int& x_row0, int& rows, int& n0, int& batch, int& split) {
    batch = 0; split = 0;
    if (p.k_splits > 1) { split = t / base_tiles; t -= split * base_tiles; }
    if (p.tiles || p.dyn_groups) {
      const int i = (t - (int)blockIdx.x) / (int)gridDim.x;
      if (i < kTileCache) {  // always true for dyn_groups (checked above)
        w_row0 = s_tiles[i][0]; x_row0 = s_tiles[i][1]; rows = s_tiles[i][2]; n0 = s_tiles[i][3];
      } else {
        const Tile tl = p.tiles[t];
        w_row0 = tl.w_row0; x_row0 = tl.x_row0; rows = tl.rows; n0 = tl.n0;
        if (p.group_counts) rows = max(0, min(BN, p.group_counts[tl.group] - tl.r0));
      }
    } else {
      if (p.nbatch > 1) { batch = t / tiles_per_batch; t -= batch * tiles_per_batch; }
      const int wb = t % p.n_w_blocks;
      const int mb = t / p.n_w_blocks;
      w_row0 = wb * BM; n0 = wb * BM; x_row0 = mb * BN;
      rows = min(BN, p.M - x_row0);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one elected lane)
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int w_row0, x_row0, rows, n0, batch, split;
        decode_tile(t, w_row0, x_row0, rows, n0, batch, split);
        if (rows <= 0) continue;  // empty group chunk: every role skips it identically
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::kStageBytes;
          uint8_t* xs = st + 2 * NA * C::kABytes;
          if (p.x_box16) {
            const int nbox = (rows + 15) >> 4;
            ptx::mbar_expect_tx(&full[stage], NB * nbox * 2048);
            for (int b = 0; b < nbox; ++b) {
              ptx::tma_load_2d(xs + b * 2048, &tm_x16, &full[stage], kb * BK, x_row0 + b * 16);
              ptx::tma_load_2d(xs + C::kBBytes + b * 2048, &tm_x16, &full[stage], kb * BK, p.x_lo_row_off + x_row0 + b * 16);
            }
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          ptx::mbar_expect_tx(&full[stage], NB * C::kBBytes);
          ptx::tma_load_2d(xs, &tm_x, &full[stage], kb * BK, x_row0);
          ptx::tma_load_2d(xs + C::kBBytes, &tm_x, &full[stage], kb * BK, p.x_lo_row_off + x_row0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one elected lane)
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::idesc_f16(Elem<T>::kFmt, BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        if (p.group_counts) {
          int w_row0, x_row0, rows, n0, batch, split;
          decode_tile(t, w_row0, x_row0, rows, n0, batch, split);
          if (rows <= 0) continue;
        }
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        ++it;
        ptx::mbar_wait(&tempty[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * C::kAccCols;
        int kb0 = 0, kb1 = num_kb;
        if (p.k_splits > 1) {
          const int split = t / base_tiles;
          kb0 = split * p.kb_per_split; kb1 = min(num_kb, kb0 + p.kb_per_split);
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);  // [NA][hi, lo] weight tiles
          const uint32_t sb = sa + 2 * NA * C::kABytes;                      // [hi, lo] token tiles
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t b_hi = ptx::smem_desc_sw128(sb + k * 32, 16, 1024);
            const uint64_t b_lo = ptx::smem_desc_sw128(sb + C::kBBytes + k * 32, 16, 1024);
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              const uint64_t a_hi = ptx::smem_desc_sw128(sa + (2 * a) * C::kABytes + k * 32, 16, 1024);
              const uint64_t a_lo = ptx::smem_desc_sw128(sa + (2 * a + 1) * C::kABytes + k * 32, 16, 1024);
              ptx::mma_f16_ss(d0 + a * BN, a_hi, b_hi, idesc, ((kb - kb0) | k) ? 1u : 0u);
              ptx::mma_f16_ss(d0 + a * BN, a_lo, b_hi, idesc, 1u);
              ptx::mma_f16_ss(d0 + a * BN, a_hi, b_lo, idesc, 1u);
            }
          }
          ptx::mma_commit(&empty[stage]);  // frees the smem slot when the MMAs above retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull[buf]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 2 + kEpiWarps) {
    // ------------------------------------------------------------ dequant producers: one thread per weight row of the stage
    const int r = threadIdx.x - (64 + 32 * kEpiWarps);  // 0..127
    const int rsw = r & 7;
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int w_row0, x_row0, rows, n0, batch, split;
      decode_tile(t, w_row0, x_row0, rows, n0, batch, split);
      if (rows <= 0) continue;
      const long long wrow = (long long)w_row0 + r;
      const bool w_ok = wrow < q.w_rows && n0 + r < p.N;
      const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * C::kStageBytes + r * 128;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          uint8_t* hi_row = st + (2 * a) * C::kABytes;
          uint8_t* lo_row = st + (2 * a + 1) * C::kABytes;
          if (w_ok) {
            dq_stage_row<T>(a ? q.fmt1 : q.fmt0, a ? q.w1 : q.w0, wrow, p.K, kb, hi_row, lo_row, rsw);
          } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              *reinterpret_cast<uint4*>(hi_row + (c << 4)) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(lo_row + (c << 4)) = make_uint4(0, 0, 0, 0);
            }
          }
        }
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&full[stage]);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int half = ew >> 2;              // which half of the token columns this warp handles
    constexpr int kChunks = BN / 32;
    constexpr int kChunksPerHalf = (kChunks + 1) / 2;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int w_row0, x_row0, rows, n0, batch, split;
      decode_tile(t, w_row0, x_row0, rows, n0, batch, split);
      if (rows <= 0) continue;
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      ++it;
      ptx::mbar_wait(&tfull[buf], bphase);
      ptx::tc_fence_after();
      const int n = n0 + quarter * 32 + lane;  // output feature owned by this thread
      const bool n_ok = n < p.N;
      const float bias = (p.bias && n_ok && split == 0) ? p.bias[n] : 0.f;
      const uint32_t trow = tmem_base + buf * C::kAccCols + ((uint32_t)(quarter * 32) << 16);
      for (int c = half * kChunksPerHalf; c < min(kChunks, (half + 1) * kChunksPerHalf); ++c) {
        if (c * 32 >= rows) break;  // warp-uniform
        uint32_t v[32];
        ptx::tmem_ld_32x32(trow + c * 32, v);
        uint32_t u[32];
        if (NA == 2) ptx::tmem_ld_32x32(trow + BN + c * 32, u);
        // output row of token j of this chunk: lane j fetches it once, broadcast below
        long long my_orow = x_row0 + c * 32 + lane;
        if (p.row_map) my_orow = (c * 32 + lane < rows) ? p.row_map[my_orow] : -1;
        ptx::tmem_ld_wait();
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = __uint_as_float(v[j]) + bias;
          if (NA == 2 && p.out_mode == OUT_F32_DUAL) {
            // raw partial accumulators; acc1 is stored below
          } else if (NA == 2) {
            t = silu(t) * __uint_as_float(u[j]);
          } else if (p.act == ACT_GELU_ERF) {
            t = gelu_erf(t);
          } else if (p.act == ACT_QUICK_GELU) {
            t = quick_gelu(t);
          }
          r[j] = t;
        }
        const int nvalid = min(32, rows - c * 32);
        const long long col = n + (long long)batch * p.out_batch_stride + (long long)split * p.split_stride;
        if (p.row_map) {
          // remapped rows (SAM window un-partition): only the residual-add mode uses this path
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const long long orow = __shfl_sync(0xffffffffu, my_orow, j);
            old[j] = (j < nvalid && n_ok && orow >= 0) ? reinterpret_cast<const float*>(p.out)[orow * p.ldo + col] : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const long long orow = __shfl_sync(0xffffffffu, my_orow, j);
            if (j < nvalid && n_ok && orow >= 0) reinterpret_cast<float*>(p.out)[orow * p.ldo + col] = old[j] + r[j];
          }
        } else if (n_ok) {
          const long long base = (long long)(x_row0 + c * 32) * p.ldo + col;
          if (p.out_mode == OUT_F32_ADD) {
            // read-modify-write: issue all loads first so their latencies overlap
            float* ptr = reinterpret_cast<float*>(p.out) + base;
            float old[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = j < nvalid ? ptr[j * p.ldo] : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) ptr[j * p.ldo] = old[j] + r[j];
          } else if (p.out_mode == OUT_T) {
            T* ptr = reinterpret_cast<T*>(p.out) + base;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) ptr[j * p.ldo] = Elem<T>::from(r[j]);
          } else if (p.out_mode == OUT_T_SPLIT) {
            T* ptr = reinterpret_cast<T*>(p.out) + base;
            T* ptr_lo = reinterpret_cast<T*>(p.out_lo) + base;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                const T hi = Elem<T>::from(r[j]);
                ptr[j * p.ldo] = hi;
                ptr_lo[j * p.ldo] = Elem<T>::from(r[j] - Elem<T>::to(hi));
              }
            }
          } else {
            float* ptr = reinterpret_cast<float*>(p.out) + base;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                ptr[j * p.ldo] = r[j];
                if (NA == 2 && p.out_mode == OUT_F32_DUAL) ptr[j * p.ldo + p.dual_stride] = __uint_as_float(u[j]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[buf]);
    }
  }

  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::kTmemCols);
}

}  // namespace lin
