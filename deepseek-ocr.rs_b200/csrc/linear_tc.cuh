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
};

struct Tile {  // grouped problems: one entry per (group, token-chunk, weight-block)
  int w_row0;   // first weight row (A coordinate)
  int x_row0;   // first token row (B coordinate)
  int rows;     // valid tokens in this tile (<= BN)
  int n0;       // first output feature (column of out) for this tile
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
  int num_tiles;           // host-side tile count (upper bound when num_tiles_dev != nullptr)
  int n_w_blocks;          // ceil(N / 128) for the dense tile decode
  int nbatch;              // > 1: X is a 3-D map [rows, nbatch, K]; tiles enumerate (batch, m, w)
  long long out_batch_stride;  // elements added to the output offset per batch
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

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <typename T, int BN, int NA, int NB>
__global__ void __launch_bounds__(kThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
              const __grid_constant__ CUtensorMap tm_x, const Params p) {
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
  const int num_tiles = p.num_tiles_dev ? min(*p.num_tiles_dev, p.num_tiles) : p.num_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_w0);
    if (NA == 2) ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_x);
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
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_batch = p.n_w_blocks * ((p.M + BN - 1) / BN);
  auto decode_tile = [&](int t, int& w_row0, int& x_row0, int& rows, int& n0, int& batch) {
    batch = 0;
    if (p.tiles) {
      const Tile tl = p.tiles[t];
      w_row0 = tl.w_row0; x_row0 = tl.x_row0; rows = tl.rows; n0 = tl.n0;
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
        int w_row0, x_row0, rows, n0, batch;
        decode_tile(t, w_row0, x_row0, rows, n0, batch);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::kStageBytes;
          ptx::mbar_expect_tx(&full[stage], C::kStageBytes);
          ptx::tma_load_2d(st, &tm_w0, &full[stage], kb * BK, w_row0);
          if (NA == 2) ptx::tma_load_2d(st + C::kABytes, &tm_w1, &full[stage], kb * BK, w_row0);
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
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * C::kAccCols;
        for (int kb = 0; kb < num_kb; ++kb) {
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
                ptx::mma_f16_ss(d0 + a * BN, ad, bd, idesc, (kb | k | b) ? 1u : 0u);
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
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int w_row0, x_row0, rows, n0, batch;
      decode_tile(t, w_row0, x_row0, rows, n0, batch);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      ptx::mbar_wait(&tfull[buf], bphase);
      ptx::tc_fence_after();
      const int n = n0 + quarter * 32 + lane;  // output feature owned by this thread
      const bool n_ok = n < p.N;
      const float bias = (p.bias && n_ok) ? p.bias[n] : 0.f;
      const uint32_t trow = tmem_base + buf * C::kAccCols + ((uint32_t)(quarter * 32) << 16);
      for (int c = half * kChunksPerHalf; c < min(kChunks, (half + 1) * kChunksPerHalf); ++c) {
        if (c * 32 >= rows) break;  // warp-uniform
        uint32_t v[32];
        ptx::tmem_ld_32x32(trow + c * 32, v);
        uint32_t u[32];
        if (NA == 2) ptx::tmem_ld_32x32(trow + BN + c * 32, u);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int m = c * 32 + j;
          if (m >= rows || !n_ok) continue;
          float r = __uint_as_float(v[j]) + bias;
          if (NA == 2) {
            r = silu(r) * __uint_as_float(u[j]);
          } else if (p.act == ACT_GELU_ERF) {
            r = gelu_erf(r);
          } else if (p.act == ACT_QUICK_GELU) {
            r = quick_gelu(r);
          }
          long long orow = x_row0 + m;
          if (p.row_map) {
            const int mapped = p.row_map[orow];
            if (mapped < 0) continue;
            orow = mapped;
          }
          const long long o = orow * p.ldo + n + batch * p.out_batch_stride;
          if (p.out_mode == OUT_T) {
            reinterpret_cast<T*>(p.out)[o] = Elem<T>::from(r);
          } else if (p.out_mode == OUT_T_SPLIT) {
            const T hi = Elem<T>::from(r);
            reinterpret_cast<T*>(p.out)[o] = hi;
            reinterpret_cast<T*>(p.out_lo)[o] = Elem<T>::from(r - Elem<T>::to(hi));
          } else if (p.out_mode == OUT_F32) {
            reinterpret_cast<float*>(p.out)[o] = r;
          } else {
            reinterpret_cast<float*>(p.out)[o] += r;
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
