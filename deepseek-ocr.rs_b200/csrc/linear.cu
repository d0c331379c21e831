// Host launcher + instantiations of the tcgen05 linear kernel (linear_tc.cuh).
#include "linear_tc.cuh"
#include "linear_sk.cuh"
#include "linear_pair.cuh"
#include "linear_dq.cuh"
#include "dsq.h"
#include "kernels.h"
#include "tmap.h"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace dsocr {

namespace {

template <typename T, int BN, int NA, int NB>
void launch_inst(const LinearCall& c, const lin::Params& p, const CUtensorMap& w0, const CUtensorMap& w1,
                 const CUtensorMap& x, const CUtensorMap& x16, int grid, cudaStream_t stream) {
  using C = lin::Cfg<BN, NA, NB>;
  auto kern = lin::linear_kernel<T, BN, NA, NB>;
  static PerDeviceOnce once;  // per instantiation
  once.run([&] {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes),
               "linear: set max dynamic smem");
  });
  kern<<<grid, lin::kThreads, C::kSmemBytes, stream>>>(w0, w1, x, x16, p);
  launch_check(c.tag ? c.tag : "linear");
}

template <typename T, int NA, int NB>
void launch_bn(int bn, const LinearCall& c, const lin::Params& p, const CUtensorMap& w0, const CUtensorMap& w1,
               const CUtensorMap& x, const CUtensorMap& x16, int grid, cudaStream_t stream) {
  switch (bn) {
    case 32: launch_inst<T, 32, NA, NB>(c, p, w0, w1, x, x16, grid, stream); break;
    case 64: launch_inst<T, 64, NA, NB>(c, p, w0, w1, x, x16, grid, stream); break;
    case 128: launch_inst<T, 128, NA, NB>(c, p, w0, w1, x, x16, grid, stream); break;
    case 256:
      if constexpr (NA == 1) { launch_inst<T, 256, NA, NB>(c, p, w0, w1, x, x16, grid, stream); break; }
    default: throw std::runtime_error("linear: unsupported token tile " + std::to_string(bn));
  }
}

template <typename T>
void launch_t(int bn, const LinearCall& c, const lin::Params& p, const CUtensorMap& w0, const CUtensorMap& w1,
              const CUtensorMap& x, const CUtensorMap& x16, int grid, cudaStream_t stream) {
  const int na = c.w1 ? 2 : 1;
  const int nb = c.x_parts;
  if (na == 1 && nb == 1) launch_bn<T, 1, 1>(bn, c, p, w0, w1, x, x16, grid, stream);
  else if (na == 1 && nb == 2) launch_bn<T, 1, 2>(bn, c, p, w0, w1, x, x16, grid, stream);
  else if (na == 2 && nb == 1) launch_bn<T, 2, 1>(bn, c, p, w0, w1, x, x16, grid, stream);
  else launch_bn<T, 2, 2>(bn, c, p, w0, w1, x, x16, grid, stream);
}

template <typename T, int BN, int NA>
void launch_sk_inst(const LinearCall& c, const lin::SkParams& p, const CUtensorMap& w0, const CUtensorMap& w1,
                    const CUtensorMap& x16, int grid, cudaStream_t stream) {
  using C = lin::Cfg<BN, NA, 2>;
  auto kern = lin::linear_sk_kernel<T, BN, NA>;
  static PerDeviceOnce once;  // per instantiation
  once.run([&] {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes),
               "linear_sk: set max dynamic smem");
  });
  kern<<<grid, lin::kThreads, C::kSmemBytes, stream>>>(w0, w1, x16, p);
  launch_check(c.tag ? c.tag : "linear_sk");
}

template <typename T>
void launch_sk(int bn, bool dual, const LinearCall& c, const lin::SkParams& p, const CUtensorMap& w0,
               const CUtensorMap& w1, const CUtensorMap& x16, int grid, cudaStream_t stream) {
  if (dual) {
    if (bn == 32) launch_sk_inst<T, 32, 2>(c, p, w0, w1, x16, grid, stream);
    else if (bn == 64) launch_sk_inst<T, 64, 2>(c, p, w0, w1, x16, grid, stream);
    else launch_sk_inst<T, 128, 2>(c, p, w0, w1, x16, grid, stream);
  } else {
    if (bn == 32) launch_sk_inst<T, 32, 1>(c, p, w0, w1, x16, grid, stream);
    else if (bn == 64) launch_sk_inst<T, 64, 1>(c, p, w0, w1, x16, grid, stream);
    else launch_sk_inst<T, 128, 1>(c, p, w0, w1, x16, grid, stream);
  }
}

// Large-M dense GEMM on CTA pairs (linear_pair.cuh).  Returns false when the problem does not qualify.
template <typename T, int ACT, int MODE>
void launch_pair_inst(const LinearCall& c, const lin::PairParams& p, const CUtensorMap& w, const CUtensorMap& x,
                      const CUtensorMap& o, int num_sms, cudaStream_t stream) {
  auto kern = lin::linear_pair_kernel<T, ACT, MODE>;
  static std::atomic<int> max_pairs_s{0};  // per instantiation (the same for every device of one box)
  static PerDeviceOnce once;
  once.run([&] {
    cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lin::kPairSmemBytes),
               "linear_pair: set max dynamic smem");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_sms & ~1); cfg.blockDim = dim3(lin::kPairThreads); cfg.dynamicSmemBytes = lin::kPairSmemBytes;
    int n = 0;
    cuda_check(cudaOccupancyMaxActiveClusters(&n, kern, &cfg), "linear_pair: cluster occupancy");
    if (n < 1) throw std::runtime_error("linear_pair: no CTA pair fits on this device");
    max_pairs_s = std::min(n, num_sms / 2);
  });
  const int max_pairs = max_pairs_s.load();
  const int pairs = std::min(max_pairs, p.num_tiles);
  kern<<<2 * pairs, lin::kPairThreads, lin::kPairSmemBytes, stream>>>(w, x, o, p);
  launch_check(c.tag ? c.tag : "linear_pair");
}

template <typename T>
void launch_pair(const LinearCall& c, const lin::PairParams& p, const CUtensorMap& w, const CUtensorMap& x,
                 const CUtensorMap& o, int num_sms, cudaStream_t stream) {
  if (c.row_map) launch_pair_inst<T, lin::ACT_NONE, lin::kPairMapped>(c, p, w, x, o, num_sms, stream);
  else if (c.out_mode == lin::OUT_F32_ADD) launch_pair_inst<T, lin::ACT_NONE, lin::OUT_F32_ADD>(c, p, w, x, o, num_sms, stream);
  else if (c.out_mode == lin::OUT_F32) launch_pair_inst<T, lin::ACT_NONE, lin::OUT_F32>(c, p, w, x, o, num_sms, stream);
  else if (c.act == lin::ACT_GELU_ERF) launch_pair_inst<T, lin::ACT_GELU_ERF, lin::OUT_T>(c, p, w, x, o, num_sms, stream);
  else if (c.act == lin::ACT_QUICK_GELU) launch_pair_inst<T, lin::ACT_QUICK_GELU, lin::OUT_T>(c, p, w, x, o, num_sms, stream);
  else launch_pair_inst<T, lin::ACT_NONE, lin::OUT_T>(c, p, w, x, o, num_sms, stream);
}

bool linear_pair(const LinearCall& c, DType dt, int num_sms, cudaStream_t stream) {
  static const bool off = getenv("DSOCR_NO_PAIR") != nullptr;
  if (off || c.tiles || c.dyn_groups || c.w1 || c.x_parts != 1 || c.nbatch > 1 || c.k_splits > 1 || c.w_tiled ||
      c.N < 256 || c.M < 2048 || c.out_mode == lin::OUT_T_SPLIT || c.out_mode == lin::OUT_F32_DUAL ||
      (c.row_map && (c.out_mode != lin::OUT_F32_ADD || (long long)c.M * c.ldo >= (1LL << 31))) ||
      (c.act && c.out_mode != lin::OUT_T) ||
      (c.out_mode == lin::OUT_T && ((c.ldo * 2) % 16 || (reinterpret_cast<uintptr_t>(c.out) & 15))))
    return false;
  lin::PairParams p{};
  p.M = c.M; p.N = c.N; p.K = c.K; p.bias = c.bias; p.out = c.out; p.ldo = c.ldo; p.row_map = c.row_map;
  p.act = c.act; p.out_mode = c.out_mode;
  p.n_w_blocks = (c.N + 255) / 256;
  p.num_tiles = p.n_w_blocks * ((c.M + lin::kPairN - 1) / lin::kPairN);
  const long long w_rows = c.w_rows ? c.w_rows : c.N;
  CUtensorMap w = tmap::make_2d_16bit(c.w0, w_rows, c.K, c.ldw ? c.ldw : c.K, lin::BM, lin::BK);
  CUtensorMap x = tmap::make_2d_16bit(c.x, c.x_rows, c.K, c.ldx ? c.ldx : c.K, 128, lin::BK);
  // 16-bit outputs are written by TMA from a [16 tokens][128 features] staging tile; rows >= M / features >= N clip
  CUtensorMap o = c.out_mode == lin::OUT_T
                      ? tmap::make_2d_16bit_store(c.out, c.M, c.N, c.ldo, lin::kPairChunk, lin::BM)
                      : w;
  if (dt == DType::BF16) launch_pair<__nv_bfloat16>(c, p, w, x, o, num_sms, stream);
  else launch_pair<__half>(c, p, w, x, o, num_sms, stream);
  return true;
}

// Dequant-fused GEMM over DSQ snapshot weights (linear_dq.cuh)
template <typename T, int BN, int NA>
void launch_dq_inst(const LinearCall& c, const lin::Params& p, const lin::DqWeights& q, const CUtensorMap& x,
                    const CUtensorMap& x16, int grid, cudaStream_t stream) {
  using C = lin::DqCfg<BN, NA>;
  auto kern = lin::linear_dq_kernel<T, BN, NA>;
  static PerDeviceOnce once;  // per instantiation
  once.run([&] { cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes), "linear_dq: set max dynamic smem"); });
  kern<<<grid, lin::kDqThreads, C::kSmemBytes, stream>>>(x, x16, p, q);
  launch_check(c.tag ? c.tag : "linear_dq");
}

template <typename T>
void launch_dq(int bn, const LinearCall& c, const lin::Params& p, const lin::DqWeights& q, const CUtensorMap& x,
               const CUtensorMap& x16, int grid, cudaStream_t stream) {
  if (c.q1) {
    switch (bn) {
      case 32: launch_dq_inst<T, 32, 2>(c, p, q, x, x16, grid, stream); return;
      case 64: launch_dq_inst<T, 64, 2>(c, p, q, x, x16, grid, stream); return;
      case 128: launch_dq_inst<T, 128, 2>(c, p, q, x, x16, grid, stream); return;
    }
  } else {
    switch (bn) {
      case 32: launch_dq_inst<T, 32, 1>(c, p, q, x, x16, grid, stream); return;
      case 64: launch_dq_inst<T, 64, 1>(c, p, q, x, x16, grid, stream); return;
      case 128: launch_dq_inst<T, 128, 1>(c, p, q, x, x16, grid, stream); return;
      case 256: launch_dq_inst<T, 256, 1>(c, p, q, x, x16, grid, stream); return;
    }
  }
  throw std::runtime_error("linear_dq: unsupported token tile " + std::to_string(bn));
}

int dq_fmt(const QuantWeight& w) {
  switch (w.fmt) {
    case DsqDType::Q8_0: return 8;
    case DsqDType::Q4K: return 12;
    case DsqDType::Q6K: return 14;
    default: return 0;  // float records are stored as f32 rows
  }
}

// Decode-time expert GEMM over fixed-capacity groups (linear_sk.cuh).
void linear_streamk(const LinearCall& c, DType dt, int num_sms, cudaStream_t stream) {
  const bool dual = c.w1 != nullptr;
  const int bn = c.bn;
  if (!c.group_counts || (bn != 32 && bn != 64 && bn != 128) || c.N % lin::BM || c.dyn_groups > 256 || c.k_splits > 1 ||
      c.nbatch > 1 || !c.sk_ws || !c.sk_flags || c.x_parts != 2 || c.bias || c.row_map || c.act ||
      (c.out_mode != lin::OUT_T_SPLIT && c.out_mode != lin::OUT_F32))
    throw std::runtime_error("linear: unsupported device-scheduled grouped problem");
  lin::SkParams p{};
  p.K = c.K; p.x_lo_row_off = c.x_lo_row_off; p.out = c.out; p.out_lo = c.out_lo; p.ldo = c.ldo; p.out_mode = c.out_mode;
  p.counts = c.group_counts; p.groups = c.dyn_groups; p.wpg = c.N / lin::BM; p.cap = c.dyn_cap; p.w_rows = c.N;
  p.ws = c.sk_ws; p.flags = c.sk_flags;
  if (c.w_tiled) { p.w0_tiled = (const uint8_t*)c.w0; p.w1_tiled = (const uint8_t*)c.w1; }
  const long long w_rows = c.w_rows ? c.w_rows : (long long)c.N * c.dyn_groups;
  CUtensorMap w0 = tmap::make_2d_16bit(c.w0, w_rows, c.K, c.ldw ? c.ldw : c.K, lin::BM, lin::BK);
  CUtensorMap w1 = dual ? tmap::make_2d_16bit(c.w1, w_rows, c.K, c.ldw ? c.ldw : c.K, lin::BM, lin::BK) : w0;
  CUtensorMap x16 = tmap::make_2d_16bit(c.x, c.x_rows, c.K, c.ldx ? c.ldx : c.K, 16, lin::BK);
  if (dt == DType::BF16) launch_sk<__nv_bfloat16>(bn, dual, c, p, w0, w1, x16, num_sms, stream);
  else launch_sk<__half>(bn, dual, c, p, w0, w1, x16, num_sms, stream);
}

}  // namespace

int linear_pick_bn(long long m, bool dual) {
  if (m <= 32) return 32;
  if (m <= 64) return 64;
  if (m <= 128 || dual) return 128;
  return 256;
}

int linear_plan_splits(long long M, int N, int K, int num_sms, bool dual) {
  // Decode-sized problems (a few hundred token rows) have fewer output tiles than the GPU has SMs and each tile is a
  // chain of K/64 dependent stages: cut K so that the tiles x splits fill the SMs.  Large M (prefill) never splits.
  if (M > 1024) return 1;
  const int bn = linear_pick_bn(M, dual);
  const int base = ((N + lin::BM - 1) / lin::BM) * (int)((M + bn - 1) / bn);
  const int num_kb = K / lin::BK;
  const int s = std::min(num_kb / 2, num_sms / base);
  if (s < 2) return 1;
  const int kb_per = (num_kb + s - 1) / s;
  return (num_kb + kb_per - 1) / kb_per;  // no empty splits
}

void linear(const LinearCall& c, DType dt, int num_sms, cudaStream_t stream) {
  if (c.K % lin::BK != 0) throw std::runtime_error("linear: K must be a multiple of 64, got " + std::to_string(c.K));
  if (c.M <= 0 && !c.tiles && !c.dyn_groups) return;
  if (c.q0) {
    if (c.x_parts != 2 || c.nbatch > 1 || c.dyn_groups || c.w_tiled || c.w0 || c.w1)
      throw std::runtime_error("linear: the dequant-fused kernel needs hi/lo activations, plain or table-grouped tiles");
    if (c.q0->K != c.K || (c.q1 && (c.q1->K != c.K || c.q1->N != c.q0->N || c.q1->count != c.q0->count)))
      throw std::runtime_error("linear: DSQ weight shape does not match the call");
    const int be = dsq_block_elems(c.q0->fmt), be1 = c.q1 ? dsq_block_elems(c.q1->fmt) : 0;
    if ((be && c.K % be) || (be1 && c.K % be1)) throw std::runtime_error("linear: K is not a multiple of the quantisation block");
  } else {
    if (c.dyn_groups && c.sk_ws) { linear_streamk(c, dt, num_sms, stream); return; }
    if (linear_pair(c, dt, num_sms, stream)) return;
  }
  const bool dual = c.w1 != nullptr || c.q1 != nullptr;
  const int bn = c.bn ? c.bn : linear_pick_bn(c.tiles ? c.tile_rows_hint : c.M, dual);

  lin::Params p{};
  p.M = c.M; p.N = c.N; p.K = c.K;
  p.x_lo_row_off = c.x_lo_row_off;
  p.bias = c.bias; p.out = c.out; p.out_lo = c.out_lo; p.ldo = c.ldo; p.row_map = c.row_map;
  p.act = c.act; p.out_mode = c.out_mode; p.swiglu = dual ? 1 : 0;
  p.tiles = reinterpret_cast<const lin::Tile*>(c.tiles);
  p.num_tiles_dev = c.num_tiles_dev;
  p.group_counts = c.group_counts;
  p.n_w_blocks = (c.N + lin::BM - 1) / lin::BM;
  p.nbatch = c.nbatch > 1 ? c.nbatch : 1;
  p.out_batch_stride = c.out_batch_stride;
  if (c.dyn_groups) {
    if (!c.group_counts || !c.bn || c.N % lin::BM || c.dyn_groups > 256 || c.k_splits > 1 || p.nbatch > 1)
      throw std::runtime_error("linear: device-scheduled groups need counts, a fixed token tile and N % 128 == 0");
    p.dyn_groups = c.dyn_groups; p.dyn_wpg = c.N / lin::BM; p.dyn_cap = c.dyn_cap; p.dyn_w_rows = c.N;
    p.num_tiles = c.dyn_groups * p.dyn_wpg * ((c.dyn_cap + bn - 1) / bn);  // upper bound; the kernel compacts
    if (p.num_tiles > 64 * num_sms) throw std::runtime_error("linear: too many grouped units for the tile cache");
  } else if (c.tiles) p.num_tiles = c.max_tiles;
  else p.num_tiles = p.n_w_blocks * (int)((c.M + bn - 1) / bn) * p.nbatch;
  p.k_splits = 1; p.kb_per_split = c.K / lin::BK; p.split_stride = c.split_stride; p.dual_stride = c.dual_stride;
  if (c.k_splits > 1) {
    if (c.tiles) throw std::runtime_error("linear: split-K is not supported for grouped problems");
    if (c.out_mode != lin::OUT_F32 && c.out_mode != lin::OUT_F32_DUAL) throw std::runtime_error("linear: split-K needs f32 partial outputs");
    const int num_kb = c.K / lin::BK;
    p.kb_per_split = (num_kb + c.k_splits - 1) / c.k_splits;
    p.k_splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
    if (p.k_splits != c.k_splits) throw std::runtime_error("linear: k_splits must divide the k-blocks without empty splits");
    p.num_tiles *= p.k_splits;
  }
  if (p.num_tiles <= 0) return;
  if (c.w_tiled) {
    if (c.ldw && c.ldw != c.K) throw std::runtime_error("linear: tiled weights have no row pitch");
    p.w0_tiled = (const uint8_t*)c.w0; p.w1_tiled = (const uint8_t*)c.w1;
  }

  const long long w_rows = c.w_rows ? c.w_rows : c.N;
  CUtensorMap w0{}, w1{};
  if (!c.q0) {
    w0 = tmap::make_2d_16bit(c.w0, w_rows, c.K, c.ldw ? c.ldw : c.K, lin::BM, lin::BK);
    w1 = dual ? tmap::make_2d_16bit(c.w1, w_rows, c.K, c.ldw ? c.ldw : c.K, lin::BM, lin::BK) : w0;
  }
  CUtensorMap x;
  if (p.nbatch > 1) {
    // X[row, batch, k]: row stride ldx elements, batch stride x_batch_stride elements.
    x = tmap::make_3d_16bit(c.x, c.K, p.nbatch, c.x_rows, (uint64_t)c.x_batch_stride * 2, (uint64_t)c.ldx * 2, lin::BK,
                            1, bn);
  } else {
    x = tmap::make_2d_16bit(c.x, c.x_rows, c.K, c.ldx ? c.ldx : c.K, bn, lin::BK);
  }
  CUtensorMap x16 = x;
  p.x_box16 = 0;
  if ((c.tiles || c.dyn_groups) && p.nbatch == 1 && bn <= 128) {  // grouped: tiles rarely fill the token tile
    x16 = tmap::make_2d_16bit(c.x, c.x_rows, c.K, c.ldx ? c.ldx : c.K, 16, lin::BK);
    p.x_box16 = 1;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (c.q0) {
    lin::DqWeights q{};
    auto planes = [](const QuantWeight& w) {
      return DsqPlanes{(const uint8_t*)w.a.p, (const uint8_t*)w.b.p, (const uint8_t*)w.c.p, (const uint8_t*)w.d.p};
    };
    q.w0 = planes(*c.q0); q.fmt0 = dq_fmt(*c.q0);
    q.w1 = planes(c.q1 ? *c.q1 : *c.q0); q.fmt1 = dq_fmt(c.q1 ? *c.q1 : *c.q0);
    q.w_rows = c.q0->N * c.q0->count;
    if (dt == DType::BF16) launch_dq<__nv_bfloat16>(bn, c, p, q, x, x16, grid, stream);
    else launch_dq<__half>(bn, c, p, q, x, x16, grid, stream);
    return;
  }
  if (dt == DType::BF16) launch_t<__nv_bfloat16>(bn, c, p, w0, w1, x, x16, grid, stream);
  else launch_t<__half>(bn, c, p, w0, w1, x, x16, grid, stream);
}

}  // namespace dsocr
