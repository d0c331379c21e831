// DRAFT FOR THE NEXT ROUND - NOT LINKED INTO libdsocr.so AND NOT YET RUN ON A GPU (build.py compiles csrc/*.cu only;
// this file is compile-checked with `nvcc -c`).  DESIGN.md section 8, item 2.
//
// Dequant-fused tensor-core GEMM for DSQ snapshots (prefill, and decode steps of more than 4 pages):
//     out[M, N] (+)= X[M, K] . dequant(W)[N, K]^T          (NA = 2: silu(X.Wg^T) * (X.Wu^T))
// Same tile as linear_tc.cuh - weights are the MMA "A" operand (128 features = 128 TMEM lanes), BN tokens the "B" operand -
// but the weight half of every shared-memory stage is *written by four producer warps* instead of TMA: thread r of the
// 128 dequantises the 64 weights of row (w_row0 + r) for k-block kb with dsq_dequant64 (csrc/dsq_dequant.h, verified on
// the CPU against the oracle), splits each value into hi + lo 16-bit parts and stores the 128-byte row into the
// 128B-swizzled K-major layout TMA would have produced (16-byte chunk q of row r at r*128 + ((q ^ (r & 7)) << 4)).
// After `fence.proxy.async` one lane per warp arrives on the stage's full barrier (count = 1 TMA expect_tx + 4 warps).
// Per 16-wide k step the issuer runs  A_hi.B_hi + A_hi.B_lo + A_lo.B_hi  into one accumulator (the lo.lo term is
// below f32 resolution), so the product carries ~22 significant bits of both operands: logits stay within ~1e-5 of the
// f32-dequant oracle, as on the float path (hi/lo activations), instead of the 2e-3 a single bf16 rounding of the weights
// would cost.  Tokens arrive through the same TMA map as in linear_tc.cuh (hi rows, lo rows at x_lo_row_off).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "../dsq_dequant.h"
#include "../ptx.cuh"

namespace dsocr {
namespace dq {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8, kDqWarps = 4;
constexpr int kThreads = 64 + 32 * kEpiWarps + 32 * kDqWarps;  // TMA | MMA | 8 epilogue | 4 dequant warps

struct Params {
  int M, N, K;
  int fmt;                 // 8 / 12 / 14 / 0
  DsqPlanes w[2];          // planes of the weight (and of the second weight for NA == 2)
  long long w_row_base;    // first row inside the stacked planes (expert * N)
  int x_lo_row_off;        // row offset of the lo activation rows inside the X tensor map
  float* out;              // [M, ldo] f32
  long long ldo;
  int accumulate;          // out += result (residual add; one writer per element)
};

template <int BN, int NA>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;                 // one 16-bit weight tile
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = NA * 2 * kABytes + 2 * kBBytes;   // hi + lo of every weight, hi + lo of the tokens
  static constexpr int kStages = (200 * 1024 / kStageBytes) > 6 ? 6 : (200 * 1024 / kStageBytes);
  static constexpr int kAccCols = NA * BN;
  static constexpr int kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64) ? 64 : (2 * kAccCols <= 128) ? 128
                                   : (2 * kAccCols <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(kStages >= 2, "stage does not fit");
};

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <typename T, int BN, int NA>
__global__ void __launch_bounds__(kThreads, 1)
linear_dq_kernel(const __grid_constant__ CUtensorMap tm_x, const Params p) {
  using C = Cfg<BN, NA>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.K / BK;
  const int n_w_blocks = (p.N + BM - 1) / BM;
  const int num_tiles = n_w_blocks * ((p.M + BN - 1) / BN);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(&full[s], 1 + kDqWarps);  // TMA producer (expect_tx) + one arrival per dequant warp
      ptx::mbar_init(&empty[s], 1);            // tcgen05.commit of the issuer
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

  auto decode_tile = [&](int t, int& w_row0, int& x_row0, int& rows) {
    const int wb = t % n_w_blocks, mb = t / n_w_blocks;
    w_row0 = wb * BM; x_row0 = mb * BN; rows = min(BN, p.M - x_row0);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: token tiles (hi + lo)
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int w_row0, x_row0, rows;
        decode_tile(t, w_row0, x_row0, rows);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sb = smem + stage * C::kStageBytes + NA * 2 * C::kABytes;
          ptx::mbar_expect_tx(&full[stage], 2 * C::kBBytes);
          ptx::tma_load_2d(sb, &tm_x, &full[stage], kb * BK, x_row0);
          ptx::tma_load_2d(sb + C::kBBytes, &tm_x, &full[stage], kb * BK, p.x_lo_row_off + x_row0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::idesc_f16(Elem<T>::kFmt, BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        ++it;
        ptx::mbar_wait(&tempty[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * C::kAccCols;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + NA * 2 * C::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t b_hi = ptx::smem_desc_sw128(sb + k * 32, 16, 1024);
            const uint64_t b_lo = ptx::smem_desc_sw128(sb + C::kBBytes + k * 32, 16, 1024);
#pragma unroll
            for (int a = 0; a < NA; ++a) {
              const uint64_t a_hi = ptx::smem_desc_sw128(sa + (2 * a) * C::kABytes + k * 32, 16, 1024);
              const uint64_t a_lo = ptx::smem_desc_sw128(sa + (2 * a + 1) * C::kABytes + k * 32, 16, 1024);
              ptx::mma_f16_ss(d0 + a * BN, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
              ptx::mma_f16_ss(d0 + a * BN, a_hi, b_lo, idesc, 1u);
              ptx::mma_f16_ss(d0 + a * BN, a_lo, b_hi, idesc, 1u);
            }
          }
          ptx::mma_commit(&empty[stage]);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&tfull[buf]);
      }
    }
  } else if (warp >= 2 + kEpiWarps) {
    // ------------------------------------------------------------ dequant producers: one weight row per thread
    const int r = (warp - 2 - kEpiWarps) * 32 + lane;  // row inside the 128-row tile
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int w_row0, x_row0, rows;
      decode_tile(t, w_row0, x_row0, rows);
      const bool row_ok = w_row0 + r < p.N;
      const long long grow = p.w_row_base + w_row0 + r;
      for (int kb = 0; kb < num_kb; ++kb) {
        float w[NA][64];
#pragma unroll
        for (int a = 0; a < NA; ++a) {  // the quantised bytes are requested before the stage is known to be free
          if (row_ok) dsq_dequant64(p.fmt, p.w[a], grow, p.K, kb, w[a]);
          else {
#pragma unroll
            for (int i = 0; i < 64; ++i) w[a][i] = 0.f;
          }
        }
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * C::kStageBytes;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          uint8_t* hi_row = st + (2 * a) * C::kABytes + r * 128;
          uint8_t* lo_row = st + (2 * a + 1) * C::kABytes + r * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v0 = w[a][8 * q + 2 * j], v1 = w[a][8 * q + 2 * j + 1];
              const float h0 = Elem<T>::to(Elem<T>::from(v0)), h1 = Elem<T>::to(Elem<T>::from(v1));
              hi[j] = pack2<T>(v0, v1);
              lo[j] = pack2<T>(v0 - h0, v1 - h1);
            }
            const int pos = (q ^ (r & 7)) << 4;  // 128B swizzle: chunk q of row r
            *reinterpret_cast<uint4*>(hi_row + pos) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(lo_row + pos) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        ptx::fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&full[stage]);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (f32 store / accumulate / SwiGLU)
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int kChunks = BN / 32;
    constexpr int kChunksPerHalf = (kChunks + 1) / 2;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int w_row0, x_row0, rows;
      decode_tile(t, w_row0, x_row0, rows);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      ++it;
      ptx::mbar_wait(&tfull[buf], bphase);
      ptx::tc_fence_after();
      const int n = w_row0 + quarter * 32 + lane;
      const bool n_ok = n < p.N;
      const uint32_t trow = tmem_base + buf * C::kAccCols + ((uint32_t)(quarter * 32) << 16);
      for (int c = half * kChunksPerHalf; c < min(kChunks, (half + 1) * kChunksPerHalf); ++c) {
        if (c * 32 >= rows) break;
        uint32_t v[32], u[32];
        ptx::tmem_ld_32x32(trow + c * 32, v);
        if (NA == 2) ptx::tmem_ld_32x32(trow + BN + c * 32, u);
        ptx::tmem_ld_wait();
        const int nvalid = min(32, rows - c * 32);
        if (n_ok) {
          float* ptr = p.out + (long long)(x_row0 + c * 32) * p.ldo + n;
          float old[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) old[j] = (p.accumulate && j < nvalid) ? ptr[j * p.ldo] : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float r = __uint_as_float(v[j]);
            if (NA == 2) r = silu(r) * __uint_as_float(u[j]);
            if (j < nvalid) ptr[j * p.ldo] = old[j] + r;
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

// explicit instantiations so that `nvcc -c` type-checks and assembles every variant of the draft
template __global__ void linear_dq_kernel<__nv_bfloat16, 64, 1>(const __grid_constant__ CUtensorMap, const Params);
template __global__ void linear_dq_kernel<__nv_bfloat16, 128, 1>(const __grid_constant__ CUtensorMap, const Params);
template __global__ void linear_dq_kernel<__nv_bfloat16, 64, 2>(const __grid_constant__ CUtensorMap, const Params);
template __global__ void linear_dq_kernel<__half, 128, 1>(const __grid_constant__ CUtensorMap, const Params);

}  // namespace dq
}  // namespace dsocr
