// Flash-style attention for the vision towers on tcgen05 / TMEM (head_dim 64, non-causal):
//   SAM windowed + global attention with the decomposed relative-position bias added in-kernel
//   (replaces vision/sam.rs:804-888 + the host loop compute_relative_bias :1124-1192, which materialises a
//   [B,12,S,S] f32 score tensor AND a same-sized bias tensor via a D2H/H2D round trip per layer), and
//   CLIP-L attention (vision/clip.rs:349-381, 449-453).
//
// One CTA = 128 queries of one (batch, head); it walks the keys in blocks of KV = RB * GW keys (RB whole
// rows of the GW-wide token grid) so that a key column's (kh, kw) is a compile-time function of its column
// index.  Q, K and V tiles are TMA-loaded straight out of the fused qkv activation [rows, 3, H, 64] - no
// q/k/v split or transpose pass.  S = Q K^T goes to TMEM; each softmax thread owns one query row, adds
//   bias[q, k] = Zh[q, qh - kh + GW-1] + Zw[q, qw - kw + GW-1]
// from registers (Z = q . rel_table^T is one small tensor-core GEMM per layer), runs the online softmax in
// f32 and writes P (16-bit) into 128B-swizzled smem; O_blk = P V (V consumed as an MN-major operand, i.e.
// untransposed) accumulates in TMEM across key blocks.  The running maximum is updated lazily: O (TMEM) and
// the row sum are rescaled only when a block's maximum exceeds the reference maximum by more than 2^8, so in
// the steady state the softmax threads never touch O and never wait for the P.V product.  The [S,S] matrix
// never exists in memory.  Two CTAs are co-resident per SM (<= 168 registers, 256 TMEM columns, ~112 KB smem
// each) so that one CTA's softmax (MUFU-bound) overlaps the other's MMAs and TMA waits.
#pragma once
#include "ptx.cuh"

namespace vattn {

constexpr int kThreads = 192;  // warp0 TMA, warp1 MMA, warps 2..5 softmax (one query row per thread)
constexpr int BQ = 128;
constexpr int D = 64;

struct Params {
  int S;             // tokens per (batch) sequence
  int H;             // heads
  int nblk;          // ceil(S / KV)
  float scale_log2;  // softmax scale * log2(e)
  const float* Z;    // [rows, H, zw] rel-pos products (nullptr when !HAS_BIAS)
  int zw;            // row width of Z (= 2 * zhalf)
  int zhalf;         // column offset of the Zw half
  void* out;         // [rows, H*64] 16-bit
};

template <int KV>
struct Cfg {
  // Stages: NS score tiles in TMEM (S_i+1 = Q K_i+1^T is computed while the softmax warps work on S_i), NP probability
  // tiles in shared memory, NST K / V tiles.  Budget per CTA for two CTAs per SM: 256 TMEM columns (O uses [192, 256)),
  // ~113 KB of shared memory.
  static constexpr int NS = 2 * KV <= 192 ? 2 : 1;
  static constexpr int kPAtoms = (KV + 63) / 64;
  static constexpr int kPBytes = kPAtoms * BQ * 128;
  static constexpr int NP = (NS == 2 && kPAtoms == 1) ? 2 : 1;
  static constexpr int NST = NS == 1 ? 2 : (KV <= 64 ? 4 : 3);
  static constexpr int kQBytes = BQ * D * 2;
  static constexpr int kKVBytes = KV * D * 2;
  static constexpr int kOffK = kQBytes;
  static constexpr int kOffV = kOffK + NST * kKVBytes;
  static constexpr int kOffP = (kOffV + NST * kKVBytes + 1023) / 1024 * 1024;
  static constexpr int kOffBar = kOffP + NP * kPBytes;
  static constexpr int kSmemBytes = kOffBar + 256;
  static constexpr int kTmemO = 192;
  static_assert(KV % 16 == 0 && KV <= 128, "KV");
  static_assert(kKVBytes % 1024 == 0, "stage alignment");
  static_assert(NS * KV <= kTmemO, "TMEM budget");
  static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM");
};

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

// 2^x on the FMA / ALU pipes (x <= 2^8 region of interest; very negative x flush to 2^-126): round to the nearest
// integer with the 1.5 * 2^23 trick, a cubic in the remainder f in [-0.5, 0.5] (relative error 7.5e-5, well under the
// 16-bit rounding of P), the integer added into the exponent field.  Takes a share of the exponentials off the MUFU
// pipe (16 per clock per SM), which bounds the softmax of the large-grid attention.
__device__ __forceinline__ float ex2_poly3(float x) {
  x = fmaxf(x, -126.f);
  const float xr = x + 12582912.f;
  const float f = x - (xr - 12582912.f);
  float pl = fmaf(0.0551716685295105f, f, 0.2426111251115799f);
  pl = fmaf(pl, f, 0.6932609677314758f);
  pl = fmaf(pl, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(xr) << 23));
}

// t = S * scale * log2(e) + relw[kw] for one 16-column chunk; columns past kvalid are -inf in the TAIL variants
template <int GW, int NG, bool TAIL>
__device__ __forceinline__ void chunk_t(const uint32_t (&v)[16], int c0, float scale, const float* relw, int kvalid,
                                        float (&t)[16], float (&gmax)[NG]) {
  constexpr int GWD = GW > 0 ? GW : 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = c0 + i;
    float x = GW > 0 ? fmaf(__uint_as_float(v[i]), scale, relw[c % GWD]) : __uint_as_float(v[i]) * scale;
    if (TAIL && c >= kvalid) x = -INFINITY;
    t[i] = x;
    const int g = GW > 0 ? c / GWD : 0;
    gmax[g] = fmaxf(gmax[g], x);
  }
}

// Maximum of t per key group over the S tile (TMEM loads one chunk ahead of the arithmetic).
template <int GW, int KV, int NG, bool TAIL>
__device__ __forceinline__ void max_pass(uint32_t ts, float scale, const float* relw, int kvalid, float (&gmax)[NG]) {
  constexpr int NCH = KV / 16;
#pragma unroll
  for (int g = 0; g < NG; ++g) gmax[g] = -INFINITY;
  uint32_t v[2][16];
  ptx::tmem_ld_32x16(ts, v[0]);
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    ptx::tmem_ld_wait_dep(v[ci & 1]);
    if (ci + 1 < NCH) ptx::tmem_ld_32x16(ts + (ci + 1) * 16, v[(ci + 1) & 1]);
    float t[16];
    chunk_t<GW, NG, TAIL>(v[ci & 1], ci * 16, scale, relw, kvalid, t, gmax);
  }
}

// p = 2^(t - mrow[group]) -> 16-bit, 128B-swizzled smem row; group maxima of t and the f32 row sum on the side.
// POLY of every 16 exponentials run on the FMA pipe (ex2_poly3), the rest on MUFU.
template <typename T, int GW, int KV, int NG, int POLY, bool TAIL>
__device__ __forceinline__ void exp_pass(uint32_t ts, float scale, const float* relw, const float (&mrow)[NG], int kvalid,
                                         uint8_t* prow, int rsw, float (&gmax)[NG], float& rowsum) {
  constexpr int NCH = KV / 16;
  constexpr int GWD = GW > 0 ? GW : 1;
#pragma unroll
  for (int g = 0; g < NG; ++g) gmax[g] = -INFINITY;
  float sum0 = 0.f, sum1 = 0.f;
  uint32_t v[2][16];
  ptx::tmem_ld_32x16(ts, v[0]);
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci) {
    ptx::tmem_ld_wait_dep(v[ci & 1]);
    if (ci + 1 < NCH) ptx::tmem_ld_32x16(ts + (ci + 1) * 16, v[(ci + 1) & 1]);
    const int c0 = ci * 16;
    float t[16], e[16];
    chunk_t<GW, NG, TAIL>(v[ci & 1], c0, scale, relw, kvalid, t, gmax);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float x = t[i] - mrow[GW > 0 ? (c0 + i) / GWD : 0];
      // spread the polynomial ones over the chunk so that MUFU and FMA work interleave
      const bool poly = POLY > 0 && (i % (16 / (POLY > 0 ? POLY : 1))) == 0 && i / (16 / (POLY > 0 ? POLY : 1)) < POLY;
      e[i] = poly ? ex2_poly3(x) : ptx::ex2_approx(x);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) { sum0 += e[i]; sum1 += e[i + 1]; }
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      const int c = c0 + g8 * 8;
      const int atom = c >> 6;
      const int chunk = ((c & 63) >> 3) ^ rsw;
      uint4 pk;
      pk.x = pack2<T>(e[g8 * 8 + 0], e[g8 * 8 + 1]);
      pk.y = pack2<T>(e[g8 * 8 + 2], e[g8 * 8 + 3]);
      pk.z = pack2<T>(e[g8 * 8 + 4], e[g8 * 8 + 5]);
      pk.w = pack2<T>(e[g8 * 8 + 6], e[g8 * 8 + 7]);
      *reinterpret_cast<uint4*>(prow + atom * (BQ * 128) + chunk * 16) = pk;
    }
  }
  rowsum = sum0 + sum1;
}

// GW = token-grid width (0 = no bias), RB = grid rows per key block, KV = keys per block.
// POLY = exponentials per 16 taken by the FMA-pipe polynomial.
template <typename T, int GW, int RB, int KV, int MINB, int POLY>
__global__ void __launch_bounds__(kThreads, MINB)
vattn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const Params p) {
  using C = Cfg<KV>;
  constexpr bool HAS_BIAS = GW > 0;
  constexpr int GWD = GW > 0 ? GW : 1;  // divisor that is never zero
  static_assert(!HAS_BIAS || KV == RB * GW, "key block must be whole grid rows");
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NS = C::NS, NP = C::NP, NST = C::NST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [NST]
  uint64_t* v_full = k_full + NST;         // [NST]
  uint64_t* kv_empty = v_full + NST;       // [NST]
  uint64_t* s_full = kv_empty + NST;       // [NS]  S_i complete in TMEM
  uint64_t* p_full = s_full + NS;          // [NP]  P_i in shared memory and S_i fully read
  uint64_t* p_free = p_full + NP;          // [NP]  P_i . V_i complete: the P stage may be overwritten
  uint64_t* o_full = p_free + NP;          // every P_i . V_i (phase = i & 1); only waited on by the rescale path
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  static_assert((2 + 3 * NST + NS + 2 * NP) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh % p.H;
  const int q0 = blockIdx.x * BQ;
  const int row_base = b * p.S;  // first qkv row of this sequence
  const int col_q = h * D, col_k = (p.H + h) * D, col_v = (2 * p.H + h) * D;

  if (threadIdx.x == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) { printf("vattn: smem base not 1024-aligned\n"); __trap(); }
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_kv);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < NST; ++i) { ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < NS; ++i) ptx::mbar_init(&s_full[i], 1);
    for (int i = 0; i < NP; ++i) { ptx::mbar_init(&p_full[i], 128); ptx::mbar_init(&p_free[i], 1); }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 256);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem;               // S stages: columns [s * KV, (s + 1) * KV)
  const uint32_t tmem_o = tmem + C::kTmemO;   // O: columns [192, 256)

  if (warp == 0) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, C::kQBytes);
      ptx::tma_load_2d(smem, &tm_q, q_full, col_q, row_base + q0);
      for (int j = 0; j < p.nblk; ++j) {
        const int s = j % NST;
        const uint32_t ph = (j / NST) & 1;
        ptx::mbar_wait_long(&kv_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&k_full[s], C::kKVBytes);
        ptx::tma_load_2d(smem + C::kOffK + s * C::kKVBytes, &tm_kv, &k_full[s], col_k, row_base + j * KV);
        ptx::mbar_expect_tx(&v_full[s], C::kKVBytes);
        ptx::tma_load_2d(smem + C::kOffV + s * C::kKVBytes, &tm_kv, &v_full[s], col_v, row_base + j * KV);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_qk = ptx::idesc_f16(Elem<T>::kFmt, BQ, KV);
      constexpr uint32_t idesc_pv = ptx::idesc_f16(Elem<T>::kFmt, BQ, D, 0, 1);  // B (= V) is MN-major
      const uint32_t sq = ptx::smem_u32(smem);
      const uint32_t sp = ptx::smem_u32(smem + C::kOffP);
      auto issue_qk = [&](int j) {  // S_j = Q K_j^T into score stage j % NS
        const int s = j % NST;
        ptx::mbar_wait_long(&k_full[s], (j / NST) & 1);
        ptx::tc_fence_after();
        const uint32_t sk = ptx::smem_u32(smem + C::kOffK + s * C::kKVBytes);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::mma_f16_ss(tmem_s + (j % NS) * KV, ptx::smem_desc_sw128(sq + k * 32, 16, 1024),
                          ptx::smem_desc_sw128(sk + k * 32, 16, 1024), idesc_qk, k ? 1u : 0u);
        ptx::mma_commit(&s_full[j % NS]);
      };
      ptx::mbar_wait_long(q_full, 0);
      for (int j = 0; j < NS && j < p.nblk; ++j) issue_qk(j);
      for (int j = 0; j < p.nblk; ++j) {
        const int s = j % NST;
        ptx::mbar_wait_long(&p_full[j % NP], (j / NP) & 1);  // P_j in smem, S_j fully read
        ptx::mbar_wait_long(&v_full[s], (j / NST) & 1);
        ptx::tc_fence_after();
        const uint32_t sv = ptx::smem_u32(smem + C::kOffV + s * C::kKVBytes);
        const uint32_t spj = sp + (j % NP) * C::kPBytes;
#pragma unroll
        for (int kk = 0; kk < KV / 16; ++kk) {
          const uint64_t ad = ptx::smem_desc_sw128(spj + (kk >> 2) * (BQ * 128) + (kk & 3) * 32, 16, 1024);
          const uint64_t bd = ptx::smem_desc_sw128(sv + kk * 2048, KV * 128, 1024);
          ptx::mma_f16_ss(tmem_o, ad, bd, idesc_pv, (j | kk) ? 1u : 0u);
        }
        ptx::mma_commit(o_full);
        ptx::mma_commit(&p_free[j % NP]);
        ptx::mma_commit(&kv_empty[s]);
        if (j + NS < p.nblk) issue_qk(j + NS);  // its score stage was released by p_full above
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / accumulate warps
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const int q = q0 + r;
    const bool q_ok = q < p.S;
    const int qc = q_ok ? q : p.S - 1;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int NG = HAS_BIAS ? RB : 1;      // key groups of a block that share one relh term (whole grid rows)

    float relw[HAS_BIAS ? GW : 1];
    const float* zrow = nullptr;
    int qh = 0;
    if (HAS_BIAS) {
      zrow = p.Z + ((long long)(row_base + qc) * p.H + h) * p.zw;
      qh = qc / GWD;
      const int qw = qc - qh * GWD;
      const float* zw = zrow + p.zhalf + qw + GW - 1;
#pragma unroll
      for (int kw = 0; kw < GW; ++kw) relw[kw] = zw[-kw] * kLog2e;
    }
    // the Zh terms of block j + 1 are requested while block j is processed (an L2 round trip per block otherwise:
    // 25 % of the softmax warps' stall samples in the capture before this change)
    float relh_next[NG];
    auto load_relh = [&](int j) {
#pragma unroll
      for (int rr = 0; rr < NG; ++rr) {
        const int kh = j * RB + rr;
        relh_next[rr] = (HAS_BIAS && kh < GW) ? zrow[qh - kh + GW - 1] : 0.f;
      }
    };
    load_relh(0);

    float m = -INFINITY, l = 0.f;  // m: reference maximum (log2 domain) all stored probabilities are relative to
    uint8_t* prow = smem + C::kOffP + r * 128;
    const int rsw = r & 7;
    constexpr float kRescaleThreshold = 8.0f;  // p <= 2^8 stays well inside the f16 / bf16 / f32-accumulate range

    for (int j = 0; j < p.nblk; ++j) {
      float relh[NG];
#pragma unroll
      for (int rr = 0; rr < NG; ++rr) relh[rr] = relh_next[rr] * kLog2e;
      if (j + 1 < p.nblk) load_relh(j + 1);
      const int kvalid = p.S - j * KV;  // keys with column index >= kvalid are padding (last block only)
      const bool tail = kvalid < KV;    // uniform over the CTA: the masked variants of the passes are separate code
      const uint32_t ts = tmem_s + (j % NS) * KV + lane_off;
      uint8_t* pj = prow + (j % NP) * C::kPBytes;
      ptx::mbar_wait_long(&s_full[j % NS], (j / NS) & 1);
      // the P stage is free once P_(j-NP) . V has completed (with one score stage s_full already implies it)
      if (NS > 1 && j >= NP) ptx::mbar_wait(&p_free[j % NP], ((j / NP) - 1) & 1);
      ptx::tc_fence_after();

      float gmax[NG];
      if (j == 0) {
        // first block: the reference maximum is the block maximum (a pass without exponentials)
        if (tail) max_pass<GW, KV, NG, true>(ts, p.scale_log2, relw, kvalid, gmax);
        else max_pass<GW, KV, NG, false>(ts, p.scale_log2, relw, kvalid, gmax);
#pragma unroll
        for (int rr = 0; rr < NG; ++rr) m = fmaxf(m, gmax[rr] + relh[rr]);
      }
      // One pass per block in the steady state: p = 2^(t + relh - m) against the CURRENT reference maximum, written
      // (16-bit) to swizzled smem while the block maximum is tracked on the side.  Only when a row's block maximum
      // exceeds its reference by more than 2^8 (lazy rescale) is the reference moved, O (TMEM) and l rescaled, and the
      // pass repeated - warp-uniformly, because tcgen05.ld/st are warp-collective (lanes that keep their maximum
      // rescale by 1).  The second attempt cannot ask again.
      float rowsum;
#pragma unroll 1
      for (;;) {
        float mrow[NG];
#pragma unroll
        for (int rr = 0; rr < NG; ++rr) mrow[rr] = m - relh[rr];
        if (tail) exp_pass<T, GW, KV, NG, POLY, true>(ts, p.scale_log2, relw, mrow, kvalid, pj, rsw, gmax, rowsum);
        else exp_pass<T, GW, KV, NG, POLY, false>(ts, p.scale_log2, relw, mrow, kvalid, pj, rsw, gmax, rowsum);
        float bmax = -INFINITY;
#pragma unroll
        for (int rr = 0; rr < NG; ++rr) bmax = fmaxf(bmax, gmax[rr] + relh[rr]);
        const bool need = bmax > m + kRescaleThreshold;
        if (!__any_sync(0xffffffffu, need)) break;
        float alpha = 1.f;
        if (need) {
          alpha = ptx::ex2_approx(m - bmax);
          l *= alpha;
          m = bmax;
        }
        if (j > 0) {
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
      }
      l += rowsum;
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[j % NP]);
    }
    // epilogue: O / l
    // The last P.V is awaited on ITS P-stage barrier, not on o_full: with two score stages a warp reaches this point
    // while up to two products are outstanding, and a parity wait can only tell "one phase behind" from "complete"
    // (o_full two phases behind reads as complete - the last two key blocks were missing from fast warps' rows).
    // p_free[s] was already awaited for the stage's previous product, so it is at most one phase behind here.
    ptx::mbar_wait(&p_free[(p.nblk - 1) % NP], ((p.nblk - 1) / NP) & 1);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    T* orow = reinterpret_cast<T*>(p.out) + (long long)(row_base + qc) * (p.H * D) + h * D;
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld_32x16(tmem_o + lane_off + c0, v);
      ptx::tmem_ld_wait();
      if (q_ok) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint4 pk;
          pk.x = pack2<T>(__uint_as_float(v[g8 * 8 + 0]) * inv, __uint_as_float(v[g8 * 8 + 1]) * inv);
          pk.y = pack2<T>(__uint_as_float(v[g8 * 8 + 2]) * inv, __uint_as_float(v[g8 * 8 + 3]) * inv);
          pk.z = pack2<T>(__uint_as_float(v[g8 * 8 + 4]) * inv, __uint_as_float(v[g8 * 8 + 5]) * inv);
          pk.w = pack2<T>(__uint_as_float(v[g8 * 8 + 6]) * inv, __uint_as_float(v[g8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + g8 * 8) = pk;
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 256);
}

}  // namespace vattn
