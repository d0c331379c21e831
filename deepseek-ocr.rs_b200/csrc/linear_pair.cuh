// Large-M dense linear layer on CTA pairs (tcgen05 cta_group::2):  out[m, n] = epi( sum_k X[m, k] * W[n, k] + bias[n] )
//
// The vision towers run every projection over 10^5 tokens; at that size the one-CTA kernel of linear_tc.cuh is
// bound by L2 -> shared-memory traffic (48 KB of operands per 128x256x64 MMA block).  Here two CTAs of a cluster
// (the two SMs of a TPC) share one 256 (weights) x 256 (tokens) output tile: each CTA stages its own 128 weight
// rows and only HALF of the token tile (32 KB per k-block for the same MMA work), the tensor cores of both SMs read
// the token halves from both shared memories, and each CTA ends up with its 128 output features x 256 tokens in
// its own TMEM.  One thread of the leader CTA issues the MMAs for the pair; completion barriers are multicast.
//
// Same operand convention and epilogue as linear_tc.cuh (weights = MMA A / TMEM lanes -> one output feature per
// epilogue thread).  Used for the SAM / CLIP / projector GEMMs (vision/sam.rs:656-701, vision/clip.rs:418-447).
#pragma once
#include "linear_tc.cuh"

namespace lin {

struct PairParams {
  int M, N, K;             // tokens, output features (any; tiles of 256, the tail is not stored), reduction (multiple of 64)
  const float* bias;       // [N] or nullptr
  void* out;               // [rows, ldo]
  long long ldo;
  const int* row_map;      // optional: token row -> output row (-1 = drop); OUT_F32_ADD only
  int act;                 // Act
  int out_mode;            // OUT_T, OUT_F32 or OUT_F32_ADD
  int n_w_blocks;          // ceil(N / 256)
  int num_tiles;           // n_w_blocks * ceil(M / 256)
};

constexpr int kPairN = 256;                        // tokens per tile (MMA N); each CTA loads 128 of them
constexpr int kPairStageBytes = 2 * BM * BK * 2;   // 128 weight rows + 128 token rows, 64 k each
constexpr int kPairStages = 5;
// Outputs leave through shared memory: per token-column part two buffers of [16 tokens][128 features], 16-bit ones as
// TMA tile stores, f32 residual updates (x += y, optionally through a row map) as one 512-byte bulk REDUCTION per token
// row - the add happens in the memory system, the SM never loads the old value (the per-thread read-modify-write this
// replaces kept the K = 768 projections at 0.4 of their HBM bound)
constexpr int kPairOutBufBytes = 16 * BM * 2;
constexpr int kPairOutBufBytesF32 = 16 * BM * 4;
constexpr int kPairOutBytes = 4 * 2 * kPairOutBufBytesF32;
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kPairOutBytes + 1024 + 256;
constexpr int kPairEpiWarps = 16;                  // 4 per TMEM lane quarter: the epilogue, not the MMA, is the
constexpr int kPairThreads = 64 + 32 * kPairEpiWarps;  // long pole for K = 768 -> more warps to hide its latencies
constexpr int kPairChunk = 16;                     // token columns per tcgen05.ld / per store burst

// Epilogue variants are compiled out, not branched on per element: ACT (Act) and MODE (Out; kPairMapped = residual
// add through row_map) are template parameters and full 32-token chunks skip the per-element bounds predicates.
constexpr int kPairMapped = 100;

template <typename T, int ACT, int MODE, bool FULL>
__device__ __forceinline__ void pair_store_chunk(const uint32_t (&v)[kPairChunk], float bias, int nvalid,
                                                 const PairParams& p, long long row0, int n, int my_orow) {
  const bool n_ok = n < p.N;  // feature tail of the last 256-block: computed, never stored
  float r[kPairChunk];
#pragma unroll
  for (int j = 0; j < kPairChunk; ++j) {
    float tv = __uint_as_float(v[j]) + bias;
    if (ACT == ACT_GELU_ERF) tv = gelu_erf(tv);
    else if (ACT == ACT_QUICK_GELU) tv = quick_gelu(tv);
    r[j] = tv;
  }
  if (MODE == kPairMapped) {
    float* out = reinterpret_cast<float*>(p.out) + n;
    float old[kPairChunk];
    int off[kPairChunk];  // element offsets fit 31 bits (checked on the host)
#pragma unroll
    for (int j = 0; j < kPairChunk; ++j) {
      const int orow = __shfl_sync(0xffffffffu, my_orow, j, kPairChunk);
      off[j] = orow < 0 ? -1 : orow * (int)p.ldo;
      old[j] = ((FULL || j < nvalid) && orow >= 0 && n_ok) ? out[off[j]] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kPairChunk; ++j)
      if ((FULL || j < nvalid) && off[j] >= 0 && n_ok) out[off[j]] = old[j] + r[j];
  } else if (!n_ok) {
    return;
  } else if (MODE == OUT_F32_ADD) {
    float* ptr = reinterpret_cast<float*>(p.out) + row0 * p.ldo + n;
    float old[kPairChunk];
#pragma unroll
    for (int j = 0; j < kPairChunk; ++j) old[j] = (FULL || j < nvalid) ? ptr[j * p.ldo] : 0.f;
#pragma unroll
    for (int j = 0; j < kPairChunk; ++j) if (FULL || j < nvalid) ptr[j * p.ldo] = old[j] + r[j];
  } else if (MODE == OUT_T) {
    T* ptr = reinterpret_cast<T*>(p.out) + row0 * p.ldo + n;
#pragma unroll
    for (int j = 0; j < kPairChunk; ++j) if (FULL || j < nvalid) ptr[j * p.ldo] = Elem<T>::from(r[j]);
  } else {
    float* ptr = reinterpret_cast<float*>(p.out) + row0 * p.ldo + n;
#pragma unroll
    for (int j = 0; j < kPairChunk; ++j) if (FULL || j < nvalid) ptr[j * p.ldo] = r[j];
  }
}

template <typename T, int ACT, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
linear_pair_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                   const __grid_constant__ CUtensorMap tm_out, const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_stage = smem + kPairStages * kPairStageBytes;  // [4 parts][2][16 tokens][128 features] 16-bit
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kPairOutBytes);
  uint64_t* full = bars;                    // leader only: both CTAs' tiles of a stage have landed
  uint64_t* empty = bars + kPairStages;     // per CTA: the pair's MMAs are done with this stage
  uint64_t* tfull = bars + 2 * kPairStages; // per CTA: accumulator buffer complete
  uint64_t* tempty = tfull + 2;             // leader only: both CTAs' epilogues drained the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb = p.K / BK;
  constexpr int kA = BM * BK * 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_w);
    ptx::prefetch_tmap(&tm_x);
    if (MODE == OUT_T) ptx::prefetch_tmap(&tm_out);
    for (int s = 0; s < kPairStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tfull[b], 1);
      ptx::mbar_init(&tempty[b], 2 * kPairEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) ptx::tmem_alloc_pair(tmem_slot, 512);
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (each CTA loads its halves)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < p.num_tiles; t += num_pairs) {
        const int wb = t % p.n_w_blocks, mb = t / p.n_w_blocks;
        const int w_row0 = wb * 256 + (int)rank * BM;
        const int x_row0 = mb * kPairN + (int)rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * kPairStageBytes;
          if (rank == 0) ptx::mbar_expect_tx(&full[stage], 2 * kPairStageBytes);
          const uint32_t lead_bar = ptx::mapa(ptx::smem_u32(&full[stage]), 0);
          ptx::tma_load_2d_pair(st, &tm_w, lead_bar, kb * BK, w_row0);
          ptx::tma_load_2d_pair(st + kA, &tm_x, lead_bar, kb * BK, x_row0);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA, one lane)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_f16(Elem<T>::kFmt, 256, kPairN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < p.num_tiles; t += num_pairs) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        ++it;
        ptx::mbar_wait(&tempty[buf], bphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + buf * kPairN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * kPairStageBytes);
          const uint32_t sb = sa + kA;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = ptx::smem_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bd = ptx::smem_desc_sw128(sb + k * 32, 16, 1024);
            ptx::mma_f16_ss_pair(d0, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
          ptx::mma_commit_pair(&empty[stage], 3);  // frees the stage in both CTAs
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_pair(&tfull[buf], 3);  // accumulators of both CTAs complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (each CTA: its 128 features)
    const int ew = warp - 2;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int part = ew >> 2;              // which quarter of the token columns this warp handles
    constexpr int kChunks = kPairN / kPairChunk;
    constexpr int kChunksPerPart = kChunks / (kPairEpiWarps / 4);
    int it = 0;
    for (int t = pair; t < p.num_tiles; t += num_pairs) {
      const int wb = t % p.n_w_blocks, mb = t / p.n_w_blocks;
      const int x_row0 = mb * kPairN;
      const int rows = min(kPairN, p.M - x_row0);
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      ++it;
      ptx::mbar_wait(&tfull[buf], bphase);
      ptx::tc_fence_after();
      const int n0 = wb * 256 + (int)rank * BM;                       // first output feature of this CTA's slab
      const bool slab_full = n0 + BM <= p.N && (p.ldo & 3) == 0;      // bulk reductions need whole 16-byte-aligned rows
      const int n = n0 + quarter * 32 + lane;                         // output feature owned by this thread
      const float bias = (p.bias && n < p.N) ? p.bias[n] : 0.f;
      const uint32_t trow = tmem_base + buf * kPairN + ((uint32_t)(quarter * 32) << 16);
      const int c0 = part * kChunksPerPart;
      // software pipeline: the TMEM load of chunk c + 1 is in flight while chunk c is activated and stored
      uint32_t v[2][kPairChunk];
      if (c0 * kPairChunk < rows) ptx::tmem_ld_32x16(trow + c0 * kPairChunk, v[0]);
#pragma unroll
      for (int ci = 0; ci < kChunksPerPart; ++ci) {
        const int c = c0 + ci;
        if (c * kPairChunk >= rows) break;  // warp-uniform
        ptx::tmem_ld_wait();
        if (ci + 1 < kChunksPerPart && (c + 1) * kPairChunk < rows) ptx::tmem_ld_32x16(trow + (c + 1) * kPairChunk, v[(ci + 1) & 1]);
        int my_orow = 0;
        if (MODE == kPairMapped) {
          const int jl = lane & (kPairChunk - 1);
          my_orow = (c * kPairChunk + jl < rows) ? p.row_map[x_row0 + c * kPairChunk + jl] : -1;
        }
        if (MODE == OUT_T) {
          // activation + 16-bit conversion into this part's staging buffer, then one TMA store per 16 x 128 tile
          // (rows past M and features past N are clipped by the tensor map)
          T* stage = reinterpret_cast<T*>(out_stage + (part * 2 + (ci & 1)) * kPairOutBufBytes) + quarter * 32 + lane;
#pragma unroll
          for (int j = 0; j < kPairChunk; ++j) {
            float tv = __uint_as_float(v[ci & 1][j]) + bias;
            if (ACT == ACT_GELU_ERF) tv = gelu_erf(tv);
            else if (ACT == ACT_QUICK_GELU) tv = quick_gelu(tv);
            stage[j * BM] = Elem<T>::from(tv);
          }
          ptx::fence_proxy_async();
          const bool issuer = (ew & 3) == 0 && lane == 0;
          if (issuer) ptx::bulk_wait_read_all();  // the previous store (other buffer) has left shared memory
          ptx::named_bar_sync(1 + part, 128);
          if (issuer) {
            ptx::tma_store_2d(&tm_out, out_stage + (part * 2 + (ci & 1)) * kPairOutBufBytes, wb * 256 + (int)rank * BM,
                              x_row0 + c * kPairChunk);
            ptx::bulk_commit_group();
          }
          continue;
        }
        const int nvalid = min(kPairChunk, rows - c * kPairChunk);
        if ((MODE == kPairMapped || MODE == OUT_F32_ADD || MODE == OUT_F32) && slab_full) {
          float* stage = reinterpret_cast<float*>(out_stage + (part * 2 + (ci & 1)) * kPairOutBufBytesF32);
#pragma unroll
          for (int j = 0; j < kPairChunk; ++j) {
            float tv = __uint_as_float(v[ci & 1][j]) + bias;
            if (ACT == ACT_GELU_ERF) tv = gelu_erf(tv);
            else if (ACT == ACT_QUICK_GELU) tv = quick_gelu(tv);
            stage[j * BM + quarter * 32 + lane] = tv;
          }
          ptx::fence_proxy_async();
          const bool issuer = (ew & 3) == 0;       // lanes 0..15 of this warp own one token row each
          if (issuer) ptx::bulk_wait_read_all();   // their earlier reductions have left shared memory
          ptx::named_bar_sync(1 + part, 128);
          if (issuer && lane < nvalid) {
            const long long orow = MODE == kPairMapped ? (long long)my_orow : (long long)x_row0 + c * kPairChunk + lane;
            float* dst = reinterpret_cast<float*>(p.out) + orow * p.ldo + n0;
            if (MODE == OUT_F32) ptx::bulk_store(dst, stage + lane * BM, BM * 4);  // plain f32 rows (the rel-pos products)
            else if (orow >= 0) ptx::bulk_reduce_add_f32(dst, stage + lane * BM, BM * 4);
            ptx::bulk_commit_group();
          }
          continue;
        }
        if (nvalid == kPairChunk) pair_store_chunk<T, ACT, MODE, true>(v[ci & 1], bias, kPairChunk, p, x_row0 + c * kPairChunk, n, my_orow);
        else pair_store_chunk<T, ACT, MODE, false>(v[ci & 1], bias, nvalid, p, x_row0 + c * kPairChunk, n, my_orow);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) ptx::mbar_arrive(&tempty[buf]);
        else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&tempty[buf]), 0));
      }
    }
  }

  if (MODE == OUT_T && warp >= 2 && ((warp - 2) & 3) == 0 && lane == 0) ptx::bulk_wait_read_all();
  if ((MODE == kPairMapped || MODE == OUT_F32_ADD || MODE == OUT_F32) && warp >= 2 && ((warp - 2) & 3) == 0) ptx::bulk_wait_read_all();
  // nobody leaves while the partner may still read this CTA's shared memory or signal its barriers
  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) ptx::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace lin
