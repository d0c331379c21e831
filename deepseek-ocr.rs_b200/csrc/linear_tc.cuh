// Tensor-core linear layer for sm_100a:  out[m, n] = epi( sum_k X[m, k] * W[n, k] )
//
// One persistent, warp-specialised kernel family built on tcgen05.mma with TMEM accumulators and
// TMA-fed, 128B-swizzled shared-memory stages.  The *weight* tile is the MMA "A" operand (128 output
// features = 128 TMEM lanes) and the *token* tile is the "B" operand (BN tokens = BN TMEM columns),
// i.e. the kernel computes a 128 x BN tile of out^T.  Consequences:
//   * a decode step with a handful of tokens streams the weights once at full MMA-M (no padding of
//     the token dimension up to 128),
//   * each epilogue thread owns one output feature: bias / per-feature work is a register scalar and
//     a warp stores 32 consecutive features of one token (64 B for 16-bit outputs).
// NA = 2 accumulates two weight tiles (gate / up) against the same token tile for the fused SwiGLU
// epilogue; NB = 2 accumulates a hi + lo split of the activations (x = hi + lo, both 16-bit) into the
// same accumulator, which keeps decoder numerics at ~f32 activation precision on 16-bit tensor cores.
// Grouped (MoE) problems pass a device-side tile table.
//
// Replaces candle's Tensor::matmul / conv2d(im2col)+cuBLAS call sites listed in SURVEY.md 2a
// (vision/sam.rs:656-701, vision/clip.rs:418-447, transformer/block.rs:966-1134, model/mod.rs:392-444,
// transformer/model.rs:243-270).
#pragma once
#include "ptx.cuh"

namespace lin {

enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_QUICK_GELU = 2 };
enum Out : int {
  OUT_T = 0,        // 16-bit store
  OUT_T_SPLIT = 1,  // 16-bit hi + lo stores (out, out_lo)
  OUT_F32 = 2,      // f32 store
  OUT_F32_ADD = 3,  // f32 read-modify-write (residual add); each element has exactly one writer
  OUT_F32_DUAL = 4, // NA == 2 without the SwiGLU: raw f32 accumulators, acc1 at out + dual_stride (split-K partials)
};

struct Tile {  // grouped problems: one entry per (group, token-chunk, weight-block)
  int w_row0;   // first weight row (A coordinate)
  int x_row0;   // first token row (B coordinate)
  int rows;     // valid tokens in this tile (<= BN); ignored when Params::group_counts is set
  int n0;       // first output feature (column of out) for this tile
  int group;    // expert id (index into group_counts)
  int r0;       // first row of this tile inside its group's fixed-capacity segment
};

struct Params {
  int M, N, K;             // tokens, output features, reduction
  int x_lo_row_off;        // NB == 2: row offset of the lo part inside the X tensor map
  const float* bias;       // [N] or nullptr
  void* out;               // [rows, ldo]
  void* out_lo;            // OUT_T_SPLIT only
  long long ldo;           // output row stride in elements
  const int* row_map;      // optional: token row -> output row (-1 = drop)
  int act;                 // Act
  int out_mode;            // Out
  int swiglu;              // NA == 2: out = silu(acc0) * acc1
  const Tile* tiles;       // optional grouped tile table
  const int* num_tiles_dev;  // optional device-side tile count (grouped)
  const int* group_counts;   // optional: rows of tile t = clamp(group_counts[tile.group] - tile.r0, 0, BN); empty tiles are skipped
  int num_tiles;           // host-side tile count (upper bound when num_tiles_dev != nullptr)
  int n_w_blocks;          // ceil(N / 128) for the dense tile decode
  int nbatch;              // > 1: X is a 3-D map [rows, nbatch, K]; tiles enumerate (batch, m, w)
  long long out_batch_stride;  // elements added to the output offset per batch
  // split-K (deterministic): tile t -> (split, m, w); split s reduces k-blocks [s*kb_per_split, ...) and stores
  // its f32 partial at out + s*split_stride; the consumer kernel sums the partials in a fixed order.
  int k_splits;
  int kb_per_split;
  long long split_stride;
  long long dual_stride;
  // grouped small-M problems: load the token operand in 16-row boxes and only as many as the tile has rows
  // (the unused part of the smem tile keeps stale data; those MMA columns are never stored)
  int x_box16;
  // fixed-capacity groups scheduled on the device (decode-time MoE): group g owns weight rows
  // [g*dyn_w_rows, +dyn_w_rows) and token rows [g*dyn_cap, +group_counts[g]).  Every CTA derives the same compact
  // list of non-empty (group, token-chunk, weight-block) units from group_counts and takes every gridDim-th unit,
  // so the work stays balanced however few groups are populated.
  int dyn_groups;
  int dyn_wpg;     // weight blocks (of 128 rows) per group
  int dyn_cap;
  int dyn_w_rows;
  // weights pre-tiled for streaming (retile_weights): tile (row block nb, k-block kb) = 16 KB contiguous at
  // ((nb * K/64 + kb) * 16384), already in the 128B-swizzled shared-memory layout -> one bulk copy per stage that
  // reads whole DRAM pages.  nullptr = row-major weights through the tensor maps.
  const uint8_t* w0_tiled;
  const uint8_t* w1_tiled;
};

constexpr int BM = 128;  // weight rows per tile (UMMA M)
constexpr int BK = 64;   // K elements per stage (one 128-byte swizzle atom of 16-bit elements)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

template <int BN, int NA, int NB>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = NA * kABytes + NB * kBBytes;
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
  static constexpr int kAccCols = NA * BN;  // TMEM columns per accumulator buffer
  static constexpr int kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64
                                   : (2 * kAccCols <= 128) ? 128 : (2 * kAccCols <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(2 * kAccCols <= 512, "TMEM overflow");
  static_assert(BN % 32 == 0 && BN <= 256, "BN");
};

// gelu(x) = 0.5 x (1 + erf(x / sqrt 2)).  erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// 16-bit output rounding): one MUFU.RCP + one MUFU.EX2 + ~12 FP32 ops, branch-free (the IEEE-rounded __frcp_rn
// costs a Newton step and a slow-path branch per element and bought nothing at this accuracy).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ptx::ex2_approx(-z * z * 1.4426950408889634f);
  const float erf_abs = 1.0f - poly * t * e;
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float quick_gelu(float x) { return x * rcp_approx(1.0f + ptx::ex2_approx(-2.4554669595930157f * x)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <typename T, int BN, int NA, int NB>
__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
              const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x16, const Params p) {
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
  int num_tiles = p.num_tiles_dev ? min(*p.num_tiles_dev, p.num_tiles) : p.num_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_w0);
    if (NA == 2) ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_x);
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
    for (int i = threadIdx.x; i < kTileCache; i += kThreads) {
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
    for (int i = threadIdx.x; i < kTileCache; i += kThreads) {
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
          if (p.x_box16) {
            const int nbox = (rows + 15) >> 4;
            ptx::mbar_expect_tx(&full[stage], NA * C::kABytes + NB * nbox * 2048);
            if (p.w0_tiled) {
            const size_t woff = ((size_t)(w_row0 / BM) * num_kb + kb) * C::kABytes;
            ptx::bulk_load(st, p.w0_tiled + woff, C::kABytes, &full[stage]);
            if (NA == 2) ptx::bulk_load(st + C::kABytes, p.w1_tiled + woff, C::kABytes, &full[stage]);
          } else {
            ptx::tma_load_2d(st, &tm_w0, &full[stage], kb * BK, w_row0);
            if (NA == 2) ptx::tma_load_2d(st + C::kABytes, &tm_w1, &full[stage], kb * BK, w_row0);
          }
            for (int b = 0; b < nbox; ++b) {
              ptx::tma_load_2d(st + NA * C::kABytes + b * 2048, &tm_x16, &full[stage], kb * BK, x_row0 + b * 16);
              if (NB == 2)
                ptx::tma_load_2d(st + NA * C::kABytes + C::kBBytes + b * 2048, &tm_x16, &full[stage], kb * BK,
                                 p.x_lo_row_off + x_row0 + b * 16);
            }
            if (++stage == C::kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          ptx::mbar_expect_tx(&full[stage], C::kStageBytes);
          if (p.w0_tiled) {
            const size_t woff = ((size_t)(w_row0 / BM) * num_kb + kb) * C::kABytes;
            ptx::bulk_load(st, p.w0_tiled + woff, C::kABytes, &full[stage]);
            if (NA == 2) ptx::bulk_load(st + C::kABytes, p.w1_tiled + woff, C::kABytes, &full[stage]);
          } else {
            ptx::tma_load_2d(st, &tm_w0, &full[stage], kb * BK, w_row0);
            if (NA == 2) ptx::tma_load_2d(st + C::kABytes, &tm_w1, &full[stage], kb * BK, w_row0);
          }
          if (p.nbatch > 1)
            ptx::tma_load_3d(st + NA * C::kABytes, &tm_x, &full[stage], kb * BK, batch, x_row0);
          else
            ptx::tma_load_2d(st + NA * C::kABytes, &tm_x, &full[stage], kb * BK, x_row0);
          if (NB == 2)
            ptx::tma_load_2d(st + NA * C::kABytes + C::kBBytes, &tm_x, &full[stage], kb * BK,
                             p.x_lo_row_off + x_row0);
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
          ptx::mma_commit(&empty[stage]);  // frees the smem slot when the MMAs above retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull[buf]);  // accumulator complete -> epilogue
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
