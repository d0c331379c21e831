// Kernel-level test hooks (include/dsocr_test.h).  Host f32 in/out; the kernels under test are the ones the
// engine launches.
#include "dsocr_test.h"
#include "kernels.h"
#include "dsq.h"
#include "util.h"

#include <algorithm>
#include <vector>

using namespace dsocr;

namespace {
int sm_count() {
  int dev = 0, n = 0;
  cuda_check(cudaGetDevice(&dev), "cudaGetDevice");
  cuda_check(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "sm count");
  return n;
}
DType to_dtype(int dtype) {
  if (dtype == DSOCR_BF16) return DType::BF16;
  if (dtype == DSOCR_F16) return DType::F16;
  throw std::runtime_error("dtype must be DSOCR_F16 or DSOCR_BF16");
}
}  // namespace

extern "C" int dsocr_test_linear(int dtype, int M, int N, int K, const float* x, const float* w0, const float* w1,
                                 const float* bias, int act, int out_mode, int x_parts, int bn, const int* row_map,
                                 int out_rows, float* out) {
  return guarded([&]() {
    const DType dt = to_dtype(dtype);
    const size_t xe = (size_t)M * K, we = (size_t)N * K, oe = (size_t)out_rows * N;
    std::vector<uint16_t> hx = x_parts == 2 ? to16_split(x, xe, dt) : to16(x, xe, dt);
    std::vector<uint16_t> hw0 = to16(w0, we, dt);
    DevBuf dx(hx.size() * 2), dw0(we * 2), dw1, dbias, dmap, dout, dout_lo;
    h2d(dx.p, hx.data(), hx.size() * 2);
    h2d(dw0.p, hw0.data(), we * 2);
    if (w1) { auto h = to16(w1, we, dt); dw1.alloc(we * 2); h2d(dw1.p, h.data(), we * 2); }
    if (bias) { dbias.alloc(N * 4); h2d(dbias.p, bias, N * 4); }
    if (row_map) { dmap.alloc(M * 4); h2d(dmap.p, row_map, M * 4); }
    const bool f32out = out_mode >= 2;
    dout.alloc(oe * (f32out ? 4 : 2));
    cuda_check(cudaMemset(dout.p, 0, dout.bytes), "memset");
    if (out_mode == 1) { dout_lo.alloc(oe * 2); cuda_check(cudaMemset(dout_lo.p, 0, oe * 2), "memset"); }
    if (out_mode == 3) h2d(dout.p, out, oe * 4);

    LinearCall c;
    c.w0 = dw0.p; c.w1 = w1 ? dw1.p : nullptr;
    c.x = dx.p; c.x_rows = (long long)M * x_parts; c.x_parts = x_parts; c.x_lo_row_off = M;
    c.M = M; c.N = N; c.K = K;
    c.bias = bias ? dbias.as<float>() : nullptr;
    c.out = dout.p; c.out_lo = dout_lo.p; c.ldo = N; c.row_map = row_map ? dmap.as<int>() : nullptr;
    c.act = act; c.out_mode = out_mode; c.bn = bn;
    linear(c, dt, sm_count(), 0);
    cuda_check(cudaDeviceSynchronize(), "linear kernel");
    if (f32out) {
      d2h(out, dout.p, oe * 4);
    } else {
      std::vector<uint16_t> ho(oe), hl;
      d2h(ho.data(), dout.p, oe * 2);
      if (out_mode == 1) { hl.resize(oe); d2h(hl.data(), dout_lo.p, oe * 2); }
      for (size_t i = 0; i < oe; ++i) out[i] = f16_to_32(ho[i], dt) + (out_mode == 1 ? f16_to_32(hl[i], dt) : 0.f);
    }
    return 0;
  });
}

extern "C" int dsocr_test_grouped_linear(int dtype, int E, int M, int N, int K, const int* counts, const float* x,
                                         const float* w0, const float* w1, int x_parts, float* out) {
  return guarded([&]() {
    const DType dt = to_dtype(dtype);
    const size_t xe = (size_t)M * K, we = (size_t)E * N * K, oe = (size_t)M * N;
    std::vector<uint16_t> hx = x_parts == 2 ? to16_split(x, xe, dt) : to16(x, xe, dt);
    std::vector<uint16_t> hw0 = to16(w0, we, dt);
    DevBuf dx(hx.size() * 2), dw0(we * 2), dw1, dout(oe * 4), dtiles, dnt(4);
    h2d(dx.p, hx.data(), hx.size() * 2);
    h2d(dw0.p, hw0.data(), we * 2);
    if (w1) { auto h = to16(w1, we, dt); dw1.alloc(we * 2); h2d(dw1.p, h.data(), we * 2); }
    cuda_check(cudaMemset(dout.p, 0, oe * 4), "memset");
    int maxc = 1;
    for (int e = 0; e < E; ++e) maxc = std::max(maxc, counts[e]);
    const int bn = linear_pick_bn(maxc, w1 != nullptr);
    std::vector<LinearTile> tiles;
    int row = 0;
    for (int e = 0; e < E; ++e) {
      for (int r0 = 0; r0 < counts[e]; r0 += bn)
        for (int wb = 0; wb < (N + 127) / 128; ++wb)
          tiles.push_back({e * N + wb * 128, row + r0, std::min(bn, counts[e] - r0), wb * 128});
      row += counts[e];
    }
    const int nt = (int)tiles.size();
    dtiles.alloc(std::max<size_t>(1, tiles.size()) * sizeof(LinearTile));
    if (nt) h2d(dtiles.p, tiles.data(), tiles.size() * sizeof(LinearTile));
    h2d(dnt.p, &nt, 4);
    LinearCall c;
    c.w0 = dw0.p; c.w1 = w1 ? dw1.p : nullptr; c.w_rows = (long long)E * N;
    c.x = dx.p; c.x_rows = (long long)M * x_parts; c.x_parts = x_parts; c.x_lo_row_off = M;
    c.M = M; c.N = N; c.K = K; c.out = dout.p; c.ldo = N; c.out_mode = 2;
    c.tiles = dtiles.as<LinearTile>(); c.num_tiles_dev = dnt.as<int>(); c.max_tiles = nt + 7; c.bn = bn;
    linear(c, dt, sm_count(), 0);
    cuda_check(cudaDeviceSynchronize(), "grouped linear kernel");
    d2h(out, dout.p, oe * 4);
    return 0;
  });
}

extern "C" int dsocr_test_fixedcap_linear(int dtype, int E, int cap, int N, int K, const int* counts, const float* x,
                                          const float* w0, const float* w1, int x_parts, float* out) {
  return guarded([&]() {
    const DType dt = to_dtype(dtype);
    const int M = E * cap;
    const size_t xe = (size_t)M * K, we = (size_t)E * N * K, oe = (size_t)M * N;
    std::vector<uint16_t> hx = x_parts == 2 ? to16_split(x, xe, dt) : to16(x, xe, dt);
    std::vector<uint16_t> hw0 = to16(w0, we, dt);
    DevBuf dx(hx.size() * 2), dw0(we * 2), dw1, dout(oe * 4), dcounts((size_t)E * 4);
    h2d(dx.p, hx.data(), hx.size() * 2);
    h2d(dw0.p, hw0.data(), we * 2);
    h2d(dcounts.p, counts, (size_t)E * 4);
    if (w1) { auto h = to16(w1, we, dt); dw1.alloc(we * 2); h2d(dw1.p, h.data(), we * 2); }
    cuda_check(cudaMemset(dout.p, 0, oe * 4), "memset");
    LinearCall c;
    c.w0 = dw0.p; c.w1 = w1 ? dw1.p : nullptr; c.w_rows = (long long)E * N;
    c.x = dx.p; c.x_rows = (long long)M * x_parts; c.x_parts = x_parts; c.x_lo_row_off = M;
    c.M = M; c.N = N; c.K = K; c.out = dout.p; c.ldo = N; c.out_mode = 2;
    c.dyn_groups = E; c.dyn_cap = cap; c.group_counts = dcounts.as<int>();
    DevBuf skws(linear_streamk_ws_bytes(sm_count())), skfl((size_t)sm_count() * 8);
    cuda_check(cudaMemset(skfl.p, 0, (size_t)sm_count() * 8), "memset");
    c.sk_ws = skws.as<float>(); c.sk_flags = skfl.as<int>();
    c.bn = cap <= 32 ? 32 : (cap <= 64 ? 64 : 128);
    linear(c, dt, sm_count(), 0);
    cuda_check(cudaDeviceSynchronize(), "fixed-capacity grouped linear kernel");
    d2h(out, dout.p, oe * 4);
    return 0;
  });
}

extern "C" int dsocr_test_vision_attention(int dtype, int B, int S, int H, const float* qkv, int grid, const float* rel_h,
                                           const float* rel_w, int rel_rows, float* out) {
  return guarded([&]() {
    const DType dt = to_dtype(dtype);
    const long long rows = (long long)B * S;
    const size_t qe = (size_t)rows * 3 * H * 64, oe = (size_t)rows * H * 64;
    auto hq = to16(qkv, qe, dt);
    DevBuf dq(qe * 2), dout(oe * 2), dz, dtab;
    h2d(dq.p, hq.data(), qe * 2);
    cuda_check(cudaMemset(dout.p, 0, oe * 2), "memset");
    VAttnCall c;
    c.qkv = dq.p; c.rows = rows; c.B = B; c.S = S; c.H = H; c.grid = grid; c.out = dout.p; c.scale = 0.125f;
    if (grid > 0) {
      const int zhalf = (rel_rows + 15) / 16 * 16;
      std::vector<float> tab((size_t)2 * zhalf * 64, 0.f);
      memcpy(tab.data(), rel_h, (size_t)rel_rows * 64 * 4);
      memcpy(tab.data() + (size_t)zhalf * 64, rel_w, (size_t)rel_rows * 64 * 4);
      auto ht = to16(tab.data(), tab.size(), dt);
      dtab.alloc(ht.size() * 2);
      h2d(dtab.p, ht.data(), ht.size() * 2);
      dz.alloc((size_t)rows * H * 2 * zhalf * 4);
      vision_relpos_products(dq.p, rows, H, dtab.p, zhalf, dz.as<float>(), dt, sm_count(), 0);
      c.Z = dz.as<float>(); c.zw = 2 * zhalf; c.zhalf = zhalf;
    }
    vision_attention(c, dt, 0);
    cuda_check(cudaDeviceSynchronize(), "vision attention kernel");
    std::vector<uint16_t> ho(oe);
    d2h(ho.data(), dout.p, oe * 2);
    for (size_t i = 0; i < oe; ++i) out[i] = f16_to_32(ho[i], dt);
    return 0;
  });
}

extern "C" int dsocr_test_linear_dq(int dtype, uint32_t q_dtype, int groups, const int* counts, int M, int N, int K,
                                    const uint8_t* blocks, const uint8_t* blocks1, const float* x, int bn, int k_splits,
                                    float* out) {
  return guarded([&]() -> int {
    const DType dt = to_dtype(dtype);
    const DsqDType qt = static_cast<DsqDType>(q_dtype);
    if (!dsq_block_elems(qt) || K % dsq_block_elems(qt) || K % 64) throw std::runtime_error("K must be a multiple of the block size and of 64");
    if (groups < 1 || (groups > 1 && !counts)) throw std::runtime_error("grouped call needs counts");
    QuantWeight q0, q1;
    const long long rows_total = (long long)N * groups;
    dsq_alloc(q0, qt, N, K, groups);
    dsq_upload_rows(q0, 0, blocks, qt, rows_total);
    if (blocks1) { dsq_alloc(q1, qt, N, K, groups); dsq_upload_rows(q1, 0, blocks1, qt, rows_total); }
    const size_t xe = (size_t)M * K, oe = (size_t)M * N;
    std::vector<uint16_t> hx = to16_split(x, xe, dt);
    DevBuf dx(hx.size() * 2);
    h2d(dx.p, hx.data(), hx.size() * 2);
    const bool dual = blocks1 != nullptr;
    const int ns = k_splits > 1 ? k_splits : 1;
    DevBuf dout(oe * 4 * ns * ((dual && ns > 1) ? 2 : 1));
    cuda_check(cudaMemset(dout.p, 0, dout.bytes), "memset");
    LinearCall c;
    c.q0 = &q0; c.q1 = dual ? &q1 : nullptr;
    c.x = dx.p; c.x_rows = 2LL * M; c.x_parts = 2; c.x_lo_row_off = M;
    c.M = M; c.N = N; c.K = K; c.out = dout.p; c.ldo = N; c.out_mode = 2 /* OUT_F32 */; c.bn = bn;
    std::vector<LinearTile> tiles;
    DevBuf dtiles, dnt;
    if (groups > 1) {
      const int tb = bn ? bn : 128;
      int row = 0;
      for (int g = 0; g < groups; ++g) {
        for (int r0 = 0; r0 < counts[g]; r0 += tb)
          for (int wb = 0; wb < (N + 127) / 128; ++wb) {
            LinearTile t{};
            t.w_row0 = g * N + wb * 128; t.x_row0 = row + r0; t.rows = std::min(tb, counts[g] - r0); t.n0 = wb * 128; t.group = g; t.r0 = r0;
            tiles.push_back(t);
          }
        row += counts[g];
      }
      if (row != M) throw std::runtime_error("counts must sum to M");
      dtiles.alloc(tiles.size() * sizeof(LinearTile));
      h2d(dtiles.p, tiles.data(), tiles.size() * sizeof(LinearTile));
      c.tiles = dtiles.as<LinearTile>(); c.max_tiles = (int)tiles.size(); c.bn = tb; c.w_rows = rows_total;
    }
    if (ns > 1) {
      if (groups > 1) throw std::runtime_error("split-K is not available for grouped calls");
      c.k_splits = ns; c.split_stride = (long long)oe * (dual ? 2 : 1);
      if (dual) { c.out_mode = 4 /* OUT_F32_DUAL */; c.dual_stride = (long long)oe; }
    }
    linear(c, dt, sm_count(), 0);
    cuda_check(cudaDeviceSynchronize(), "linear_dq kernel");
    if (ns == 1) { d2h(out, dout.p, oe * 4); return 0; }
    std::vector<float> part(dout.bytes / 4);
    d2h(part.data(), dout.p, dout.bytes);
    const size_t stride = oe * (dual ? 2 : 1);
    for (size_t i = 0; i < oe; ++i) {
      float g = 0.f, u = 0.f;
      for (int sidx = 0; sidx < ns; ++sidx) { g += part[sidx * stride + i]; if (dual) u += part[sidx * stride + oe + i]; }
      out[i] = dual ? g / (1.f + expf(-g)) * u : g;
    }
    return 0;
  });
}
