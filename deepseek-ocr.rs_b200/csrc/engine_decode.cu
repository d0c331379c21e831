// Engine implementation, part 2: the DeepSeek-V2 MoE decoder (prefill + lock-step batched greedy decode).
// Restates DeepseekOcrModel::generate (model/mod.rs:1870-2048) for a batch of independent pages.
#include <algorithm>
#include <cstring>

#include "engine.h"
#include "linear_tc.cuh"
#include "sampler.h"

namespace dsocr {

// One pass of the 12 decoder layers over `rows` token rows (TransformerDecoder::forward, decoder.rs:62-196;
// TransformerBlock::forward_internal, block.rs:123-190), then final RMSNorm + lm_head on `n_final` selected
// rows (the reference computes logits for every prefill row, transformer/model.rs:243-270; only the last
// row of each page is ever used, model/mod.rs:1941-1947).
void Engine::decoder_forward(float* x, long long rows, const int* row_page, const int* row_pos, int smax,
                             const int* final_rows, int n_final, float* logits, bool decode_mode) {
  const ModelConfig& c = cfg_;
  const int H = c.hidden, heads = c.heads, E = c.n_experts, K = c.topk, mi = c.moe_inter;
  const long long S = (long long)c.moe_inter * c.n_shared;
  const float scale = 1.0f / sqrtf((float)c.head_dim());
  const long long n_assign = rows * K;
  const long long inter_max = std::max<long long>(c.inter, S);

  // hi/lo split 16-bit activation buffers: [2][rows][width], lo part `rows*width` elements after hi
  void* xn16 = ws("dec_xn16", 2 * rows * H * 2).p;
  float* xn32 = ws("dec_xn32", rows * H * 4).as<float>();
  float* qkv = ws("dec_qkv32", rows * 3 * H * 4).as<float>();
  float* q = ws("dec_q32", rows * H * 4).as<float>();
  void* ctx16 = ws("dec_ctx16", 2 * rows * H * 2).p;
  void* h16 = ws("dec_h16", 2 * rows * inter_max * 2).p;
  int* topk_idx = ws("moe_topk_idx", n_assign * 4).as<int>();
  float* topk_w = ws("moe_topk_w", n_assign * 4).as<float>();
  int* counts = ws("moe_counts", 3 * E * 4 + 16).as<int>();
  int* offsets = counts + E;
  int* cursor = counts + 2 * E;
  int* ntiles = counts + 3 * E;  // [2]
  cuda_check(cudaMemsetAsync(counts, 0, E * 4, stream_), "moe counts memset");  // moe_plan re-zeroes after use
  const int bn = linear_pick_bn(std::max<long long>(1, n_assign / E * 2), true);
  const int max_chunks = (int)(n_assign / bn) + E;
  LinearTile* tiles1 = ws("moe_tiles1", (size_t)max_chunks * (mi / 128) * sizeof(LinearTile)).as<LinearTile>();
  LinearTile* tiles2 = ws("moe_tiles2", (size_t)max_chunks * (H / 128) * sizeof(LinearTile)).as<LinearTile>();
  // decode mode: every expert owns a fixed-capacity segment of `cap` rows (a token picks an expert at most once)
  const bool fused = decode_mode && rows <= fused_max_rows_ && !quantized_;  // DSQ weights: table-grouped dequant-fused GEMMs
  // expert parallel: this rank computes its eloc local experts (+ the shared experts of its own tokens) for the tokens of
  // every rank; segments live in buffers the peers can address
  const bool ep = decode_mode && ep_peers_.world > 1;
  if (ep && (!fused || rows * ep_peers_.world > ep_cap_))
    throw std::runtime_error("expert-parallel decode needs the fused float schedule and at most " + std::to_string(ep_cap_ / ep_peers_.world) + " pages per rank");
  const long long cap = ep ? ep_cap_ : rows;
  const int Eg = (ep ? ep_peers_.eloc : E) + c.n_shared;  // decode: routed experts + the shared experts as extra groups of the grouped GEMMs
  const long long perm_rows = fused ? std::max<long long>(n_assign, (long long)Eg * cap) : n_assign;
  int* perm_pos = ws("moe_perm", n_assign * 4).as<int>();
  void* xperm16 = ep ? ep_xperm_.p : ws("moe_xperm16", 2 * perm_rows * H * 2).p;
  void* hperm16 = ep ? ep_hperm_.p : ws("moe_hperm16", 2 * perm_rows * mi * 2).p;
  float* yperm = ep ? ep_y_.as<float>() : ws("moe_yperm32", perm_rows * H * 4).as<float>();
  int* counts_layers = nullptr;
  int fbn = 32;
  if (fused) {
    counts_layers = ep ? ep_counts_.as<int>() : ws("moe_counts_layers", (size_t)c.layers * Eg * 4).as<int>();
    cuda_check(cudaMemsetAsync(counts_layers, 0, (size_t)c.layers * Eg * 4, stream_), "moe counts memset");
    // every rank has consumed the previous step's segments and zeroed its counters before anyone dispatches again
    if (ep) ep_barrier(ep_peers_, ep_gen_.as<int>(), stream_);
    // token tile of the expert GEMMs; the kernel enumerates the non-empty (expert, chunk, block) units itself
    fbn = cap <= 32 ? 32 : (cap <= 64 ? 64 : 128);
    static const int fbn_env = getenv("DSOCR_FBN") ? atoi(getenv("DSOCR_FBN")) : 0;  // A/B switch
    if (fbn_env == 32 || fbn_env == 64 || fbn_env == 128) fbn = fbn_env;
    if (!sk_flags_.p) {  // stream-K hand-off flags: zero once, the kernel returns them zeroed
      sk_flags_.alloc((size_t)num_sms_ * 8);
      cuda_check(cudaMemsetAsync(sk_flags_.p, 0, (size_t)num_sms_ * 8, stream_), "stream-K flags");
      sk_ws_.alloc(linear_streamk_ws_bytes(num_sms_));
    }
  }

  // Small-M (decode) projections are split along K so that they fill the GPU; the f32 partials are reduced
  // in a fixed order by the consumer kernel (RoPE, RMSNorm, SwiGLU-reduce, MoE combine) -> deterministic.
  // weights of a call: the engine's 16-bit matrices, or the snapshot's quantised planes (dequant-fused kernel, linear_dq.cuh)
  auto setw = [&](LinearCall& lc, const DevBuf& w0, const QuantWeight& q0, const DevBuf* w1 = nullptr, const QuantWeight* q1 = nullptr) {
    if (quantized_) { lc.q0 = &q0; lc.q1 = q1; }
    else { lc.w0 = w0.p; lc.w1 = w1 ? w1->p : nullptr; lc.w_tiled = w_tiled_; }
  };
  const int sp_qkv = linear_plan_splits(rows, quantized_ ? H : 3 * H, H, num_sms_);
  const int sp_o = linear_plan_splits(rows, H, H, num_sms_);
  const int sp_dgu = linear_plan_splits(rows, c.inter, H, num_sms_, true);
  const int sp_dd = linear_plan_splits(rows, H, c.inter, num_sms_);
  const int sp_sgu = linear_plan_splits(rows, (int)S, H, num_sms_, true);
  const int sp_sd = linear_plan_splits(rows, H, (int)S, num_sms_);
  const long long part_elems = std::max<long long>({(long long)sp_qkv * rows * 3 * H, (long long)sp_o * rows * H,
                                                    (long long)sp_dgu * 2 * rows * c.inter, (long long)sp_dd * rows * H,
                                                    (long long)sp_sgu * 2 * rows * S, 1LL});
  float* partA = ws("dec_partA32", part_elems * 4).as<float>();                       // qkv / o / gate-up partials
  float* partB = ws("dec_partB32", std::max<long long>(sp_sd, sp_dd) * rows * H * 4).as<float>();  // down partials
  const float* pend = nullptr;  // partials of a down/o projection still to be added to x by the next consumer
  int pend_n = 0;
  long long pend_stride = 0;

  void* xf16 = ws("dec_xf16", 2 * (size_t)n_final * H * 2).p;
  bool have_xn = false;     // this layer's ln1 output was already produced by the previous layer's combine_norm
  bool final_done = false;  // the final norm was fused into the last layer's combine_norm
  for (int l = 0; l < c.layers; ++l) {
    DecLayerW& L = dec_[l];
    if (!have_xn) {
      rmsnorm_split(x, L.ln1.as<float>(), xn16, rows * H, nullptr, nullptr, rows, H, c.rms_eps, pend, pend_n, pend_stride, dt_, stream_);
      pend = nullptr; pend_n = 0;
    }
    have_xn = false;
    if (!quantized_) {
      LinearCall lc;  // fused q/k/v projection
      lc.tag = "dec_qkv"; lc.w0 = L.qkv_w.p; lc.x = xn16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
      lc.M = (int)rows; lc.N = 3 * H; lc.K = H; lc.ldo = 3 * H; lc.out_mode = lin::OUT_F32;
      lc.out = sp_qkv > 1 ? partA : qkv; lc.k_splits = sp_qkv; lc.split_stride = rows * 3 * H;
      lc.w_tiled = w_tiled_; linear(lc, dt_, num_sms_, stream_);
    } else {
      const QuantWeight* wq[3] = {&L.q_q, &L.q_k, &L.q_v};  // the snapshot keeps q / k / v as separate records
      for (int i = 0; i < 3; ++i) {
        LinearCall lc;
        lc.tag = "dec_qkv"; lc.q0 = wq[i]; lc.x = xn16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
        lc.M = (int)rows; lc.N = H; lc.K = H; lc.ldo = 3 * H; lc.out_mode = lin::OUT_F32;
        lc.out = (sp_qkv > 1 ? partA : qkv) + (long long)i * H; lc.k_splits = sp_qkv; lc.split_stride = rows * 3 * H;
        linear(lc, dt_, num_sms_, stream_);
      }
    }
    if (fused) {
      rope_attn_decode(sp_qkv > 1 ? partA : qkv, sp_qkv, rows * 3 * H, rope_cos_.as<float>(), rope_sin_.as<float>(),
                       kc_ptr(l), vc_ptr(l), kv_f16_, row_page, row_pos, ctx16, rows * H, rows, heads, smax, scale, dt_, stream_);
    } else {
      rope_kv(sp_qkv > 1 ? partA : qkv, rope_cos_.as<float>(), rope_sin_.as<float>(), row_page, row_pos, q,
              kc_ptr(l), vc_ptr(l), kv_f16_, rows, heads, smax, sp_qkv, rows * 3 * H, stream_);
      kv_attention(q, kc_ptr(l), vc_ptr(l), kv_f16_, row_page, row_pos, ctx16, rows * H, nullptr, rows, heads, smax, scale,
                   dt_, stream_, decode_mode ? nullptr : ws("prefill_page_spans", (size_t)2 * n_final * 4).as<int>(), decode_mode ? 0 : n_final);
    }
    {
      LinearCall lc;  // o_proj (+ residual add, fused here or in the following RMSNorm when split)
      lc.tag = "dec_o_proj"; setw(lc, L.o_w, L.q_o); lc.x = ctx16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
      lc.M = (int)rows; lc.N = H; lc.K = H; lc.ldo = H;
      if (sp_o > 1) { lc.out = partA; lc.out_mode = lin::OUT_F32; lc.k_splits = sp_o; lc.split_stride = rows * H; }
      else { lc.out = x; lc.out_mode = lin::OUT_F32_ADD; }
      linear(lc, dt_, num_sms_, stream_);
    }
    const bool fused_moe = fused && L.moe;
    if (!fused_moe)
      rmsnorm_split(x, L.ln2.as<float>(), xn16, rows * H, L.moe ? xn32 : nullptr, nullptr, rows, H, c.rms_eps,
                    sp_o > 1 ? partA : nullptr, sp_o, rows * H, dt_, stream_);
    if (!L.moe) {
      {
        LinearCall lc;  // gate/up + SwiGLU (run_dense_mlp, block.rs:1166-1177)
        lc.tag = "dec_dense_gate_up"; setw(lc, L.gate_w, L.q_gate, &L.up_w, &L.q_up); lc.x = xn16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
        lc.M = (int)rows; lc.N = c.inter; lc.K = H; lc.ldo = c.inter;
        if (sp_dgu > 1) {
          lc.out = partA; lc.out_mode = lin::OUT_F32_DUAL; lc.k_splits = sp_dgu; lc.dual_stride = rows * c.inter;
          lc.split_stride = 2 * rows * c.inter;
        } else {
          lc.out = h16; lc.out_lo = (uint16_t*)h16 + rows * c.inter; lc.out_mode = lin::OUT_T_SPLIT;
        }
        linear(lc, dt_, num_sms_, stream_);
        if (sp_dgu > 1) swiglu_reduce(partA, sp_dgu, 2 * rows * c.inter, rows * c.inter, h16, rows * c.inter, rows * c.inter, dt_, stream_);
      }
      {
        LinearCall lc;
        lc.tag = "dec_dense_down"; setw(lc, L.down_w, L.q_down); lc.x = h16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
        lc.M = (int)rows; lc.N = H; lc.K = c.inter; lc.ldo = H;
        if (sp_dd > 1) {
          lc.out = partB; lc.out_mode = lin::OUT_F32; lc.k_splits = sp_dd; lc.split_stride = rows * H;
          pend = partB; pend_n = sp_dd; pend_stride = rows * H;
        } else { lc.out = x; lc.out_mode = lin::OUT_F32_ADD; }
        linear(lc, dt_, num_sms_, stream_);
      }
    } else {
      // run_moe (block.rs:1215-1395).  In decode the shared-experts branch (3 small kernels) runs on a forked
      // stream next to the routed branch (router/plan/dispatch/2 grouped GEMMs): both are latency-bound.
      int* lcounts = fused ? counts_layers + (size_t)l * Eg : counts;
      if (fused) {
        // o_proj partial reduce + residual + RMSNorm(ln2) + router + top-k + dispatch in one kernel
        post_attn(x, sp_o > 1 ? partA : nullptr, sp_o > 1 ? sp_o : 0, rows * H, L.ln2.as<float>(), L.router_wt.as<float>(),
                  xn16, rows * H, topk_idx, topk_w, lcounts, perm_pos, xperm16, (long long)Eg * cap * H, (int)cap, rows, H, E, K,
                  c.n_shared, c.rms_eps, dt_, stream_, ep ? &ep_peers_ : nullptr, l * Eg);
        if (ep) ep_barrier(ep_peers_, ep_gen_.as<int>(), stream_);  // all ranks' rows have arrived in the owners' segments
      } else {
        moe_router(xn32, L.router_wt.as<float>(), topk_idx, topk_w, counts, rows, H, E, K, stream_);
        moe_plan(counts, offsets, cursor, tiles1, ntiles, tiles2, ntiles + 1, E, bn, mi, H, stream_);
        moe_dispatch(topk_idx, offsets, cursor, xn16, rows * H, xperm16, n_assign * H, perm_pos, n_assign, K, H, dt_, stream_);
      }
      const bool fork = !fused && rows <= 256 && !kernel_timing_enabled();
      cudaStream_t sb = fork ? stream2_ : stream_;
      if (fork) {
        cuda_check(cudaEventRecord(ev_fork_, stream_), "fork record");
        cuda_check(cudaStreamWaitEvent(sb, ev_fork_, 0), "fork wait");
      }
      {
        LinearCall lc;  // routed experts: gate/up + SwiGLU, grouped
        lc.tag = "moe_expert_gate_up"; setw(lc, ep ? ep_gate_[l] : L.exp_gate, L.q_exp_gate, ep ? &ep_up_[l] : &L.exp_up, &L.q_exp_up); lc.w_rows = (long long)(fused ? Eg : E) * mi;
        const long long pr = fused ? (long long)Eg * cap : n_assign;
        lc.x = xperm16; lc.x_rows = 2 * pr; lc.x_parts = 2; lc.x_lo_row_off = (int)pr;
        lc.M = (int)pr; lc.N = mi; lc.K = H;
        lc.out = hperm16; lc.out_lo = (uint16_t*)hperm16 + pr * mi; lc.ldo = mi; lc.out_mode = lin::OUT_T_SPLIT;
        if (fused) { lc.dyn_groups = Eg; lc.dyn_cap = (int)cap; lc.group_counts = lcounts; lc.bn = fbn; if (streamk_) { lc.sk_ws = sk_ws_.as<float>(); lc.sk_flags = sk_flags_.as<int>(); } }
        else { lc.tiles = tiles1; lc.num_tiles_dev = ntiles; lc.max_tiles = max_chunks * (mi / 128); lc.bn = bn; }
        linear(lc, dt_, num_sms_, stream_);
      }
      {
        LinearCall lc;  // routed experts: down, grouped
        lc.tag = "moe_expert_down"; setw(lc, ep ? ep_down_[l] : L.exp_down, L.q_exp_down); lc.w_rows = (long long)(fused ? Eg : E) * H;
        const long long pr = fused ? (long long)Eg * cap : n_assign;
        lc.x = hperm16; lc.x_rows = 2 * pr; lc.x_parts = 2; lc.x_lo_row_off = (int)pr;
        lc.M = (int)pr; lc.N = H; lc.K = mi; lc.out = yperm; lc.ldo = H; lc.out_mode = lin::OUT_F32;
        if (fused) { lc.dyn_groups = Eg; lc.dyn_cap = (int)cap; lc.group_counts = lcounts; lc.bn = fbn; if (streamk_) { lc.sk_ws = sk_ws_.as<float>(); lc.sk_flags = sk_flags_.as<int>(); } }
        else { lc.tiles = tiles2; lc.num_tiles_dev = ntiles + 1; lc.max_tiles = max_chunks * (H / 128); lc.bn = bn; }
        linear(lc, dt_, num_sms_, stream_);
      }
      if (!fused) {
        LinearCall lc;  // shared experts (one fused SwiGLU MLP, weights.rs:390-400)
        lc.tag = "moe_shared_gate_up"; setw(lc, L.sh_gate, L.q_sh_gate, &L.sh_up, &L.q_sh_up); lc.x = xn16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
        lc.M = (int)rows; lc.N = (int)S; lc.K = H; lc.ldo = S;
        if (sp_sgu > 1) {
          lc.out = partA; lc.out_mode = lin::OUT_F32_DUAL; lc.k_splits = sp_sgu; lc.dual_stride = rows * S; lc.split_stride = 2 * rows * S;
        } else {
          lc.out = h16; lc.out_lo = (uint16_t*)h16 + rows * S; lc.out_mode = lin::OUT_T_SPLIT;
        }
        linear(lc, dt_, num_sms_, sb);
        if (sp_sgu > 1) swiglu_reduce(partA, sp_sgu, 2 * rows * S, rows * S, h16, rows * S, rows * S, dt_, sb);
      }
      if (!fused) {
        LinearCall lc;
        lc.tag = "moe_shared_down"; setw(lc, L.sh_down, L.q_sh_down); lc.x = h16; lc.x_rows = 2 * rows; lc.x_parts = 2; lc.x_lo_row_off = (int)rows;
        lc.M = (int)rows; lc.N = H; lc.K = (int)S; lc.ldo = H;
        if (sp_sd > 1) { lc.out = partB; lc.out_mode = lin::OUT_F32; lc.k_splits = sp_sd; lc.split_stride = rows * H; }
        else { lc.out = x; lc.out_mode = lin::OUT_F32_ADD; }
        linear(lc, dt_, num_sms_, sb);
      }
      if (fork) {
        cuda_check(cudaEventRecord(ev_join_, sb), "join record");
        cuda_check(cudaStreamWaitEvent(stream_, ev_join_, 0), "join wait");
      }
      if (fused) {
        const bool last = l + 1 == c.layers;
        if (ep) ep_barrier(ep_peers_, ep_gen_.as<int>(), stream_);  // every owner's expert outputs are complete
        combine_norm(x, yperm, perm_pos, topk_w, K, nullptr, 0, rows * H,
                     last ? final_norm_.as<float>() : dec_[l + 1].ln1.as<float>(), last ? xf16 : xn16, rows * H, rows, H,
                     c.rms_eps, c.n_shared, (int)((Eg - c.n_shared) * cap), (int)cap, dt_, stream_, ep ? &ep_peers_ : nullptr);
        if (last) final_done = true; else have_xn = true;
      } else {
        moe_combine(yperm, perm_pos, topk_w, x, rows, K, H, sp_sd > 1 ? partB : nullptr, sp_sd, rows * H, stream_);
      }
    }
    if (record_taps_) record_tap("dec.hidden." + std::to_string(l), x, rows * H);
  }
  if (moe_stats_ && counts_layers) {
    moe_active_stat(counts_layers, c.layers * Eg, moe_stats_dev_.as<unsigned long long>(), stream_);
  }
  // final RMSNorm + lm_head on the selected rows
  if (!final_done)
    rmsnorm_split(x, final_norm_.as<float>(), xf16, (long long)n_final * H, nullptr, final_rows, n_final, H, c.rms_eps,
                  pend, pend_n, pend_stride, dt_, stream_);
  LinearCall lc;
  lc.tag = "lm_head"; setw(lc, lm_head_, q_lm_head_); lc.x = xf16; lc.x_rows = 2 * n_final; lc.x_parts = 2; lc.x_lo_row_off = n_final;
  lc.M = n_final; lc.N = c.vocab; lc.K = H; lc.out = logits; lc.ldo = c.vocab; lc.out_mode = lin::OUT_F32;
  linear(lc, dt_, num_sms_, stream_);
}

Engine::EpBuffers Engine::ep_attach(int rank, int world, int cap) {
  const ModelConfig& c = cfg_;
  if (quantized_) throw std::runtime_error("expert-parallel decode is not available for DSQ engines");
  if (world < 2 || world > 8 || c.n_experts % world) throw std::runtime_error("expert-parallel world size must divide the expert count (2..8)");
  cuda_check(cudaSetDevice(device_), "cudaSetDevice");
  cuda_check(cudaStreamSynchronize(stream_), "ep attach sync");
  const int eloc = c.n_experts / world, Eg = eloc + c.n_shared;
  const long long H = c.hidden, mi = c.moe_inter;
  ep_cap_ = cap;
  ep_counts_.alloc((size_t)c.layers * Eg * 4);
  ep_xperm_.alloc((size_t)2 * Eg * cap * H * 2);
  ep_hperm_.alloc((size_t)2 * Eg * cap * mi * 2);
  ep_y_.alloc((size_t)Eg * cap * H * 4);
  ep_flags_.alloc(64); ep_gen_.alloc(16);
  cuda_check(cudaMemset(ep_counts_.p, 0, ep_counts_.bytes), "ep memset");
  cuda_check(cudaMemset(ep_flags_.p, 0, 64), "ep memset");
  cuda_check(cudaMemset(ep_gen_.p, 0, 16), "ep memset");
  // contiguous weight stacks of this rank: its eloc routed experts, then the shared experts (groups E.. of the full
  // stacks); a group's bytes are contiguous in the row-major and in the pre-tiled layout alike
  ep_gate_.clear(); ep_up_.clear(); ep_down_.clear();
  ep_gate_.resize(c.layers); ep_up_.resize(c.layers); ep_down_.resize(c.layers);
  const size_t gb = (size_t)mi * H * 2;  // bytes of one expert's gate / up / down matrix
  for (int l = 0; l < c.layers; ++l) {
    DecLayerW& L = dec_[l];
    if (!L.moe) continue;
    auto stack = [&](DevBuf& dst, const DevBuf& src) {
      dst.alloc((size_t)Eg * gb);
      cuda_check(cudaMemcpy(dst.p, (const char*)src.p + (size_t)rank * eloc * gb, (size_t)eloc * gb, cudaMemcpyDeviceToDevice), "ep weights");
      cuda_check(cudaMemcpy((char*)dst.p + (size_t)eloc * gb, (const char*)src.p + (size_t)c.n_experts * gb, (size_t)c.n_shared * gb, cudaMemcpyDeviceToDevice), "ep weights");
    };
    stack(ep_gate_[l], L.exp_gate); stack(ep_up_[l], L.exp_up); stack(ep_down_[l], L.exp_down);
  }
  ep_peers_ = EpPeers();
  ep_peers_.rank = rank; ep_peers_.eloc = eloc; ep_peers_.cap = cap;  // world stays 1 until the peer tables are installed
  return EpBuffers{ep_counts_.as<int>(), ep_xperm_.p, ep_y_.as<float>(), ep_flags_.as<int>()};
}

void Engine::ep_set_peers(const EpPeers& peers) { ep_peers_ = peers; }

void Engine::ep_detach() {
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  ep_peers_ = EpPeers();
  ep_cap_ = 0;
  ep_counts_.release(); ep_xperm_.release(); ep_hperm_.release(); ep_y_.release(); ep_flags_.release(); ep_gen_.release();
  ep_gate_.clear(); ep_up_.clear(); ep_down_.clear();
}

void Engine::set_moe_stats(bool on) {
  moe_stats_ = on;
  if (on && !moe_stats_dev_.p) {
    moe_stats_dev_.alloc(16);
    cuda_check(cudaMemset(moe_stats_dev_.p, 0, 16), "stats reset");
  }
}

void Engine::moe_stats(unsigned long long out[2]) {
  out[0] = out[1] = 0;
  if (!moe_stats_dev_.p) return;
  cuda_check(cudaStreamSynchronize(stream_), "stats sync");
  cuda_check(cudaMemcpy(out, moe_stats_dev_.p, 16, cudaMemcpyDeviceToHost), "stats copy");
  cuda_check(cudaMemset(moe_stats_dev_.p, 0, 16), "stats reset");
  // the shared experts ride along as always-populated groups; report routed experts only
  const unsigned long long shared = out[1] * (unsigned long long)(cfg_.n_shared * std::max(0, cfg_.layers - cfg_.first_dense));
  out[0] = out[0] > shared ? out[0] - shared : 0;
}

// DSQ variant of decoder_forward (run_quantized_matmul, quantization.rs:164-185): every decoder linear and the
// lm_head is a dequant-fused GEMV over the snapshot's Q8_0 / Q4_K / Q6_K blocks with f32 activations.  Routed
// experts need no dispatch: assignment row a = (token a/topk, slot a%topk) selects its expert's weights directly.
void Engine::decoder_forward_dsq(float* x, long long rows, const int* row_page, const int* row_pos, int smax,
                                 const int* final_rows, int n_final, float* logits) {
  const ModelConfig& c = cfg_;
  const int H = c.hidden, heads = c.heads, E = c.n_experts, K = c.topk, mi = c.moe_inter;
  const long long S = (long long)c.moe_inter * c.n_shared;
  const float scale = 1.0f / sqrtf((float)c.head_dim());
  const long long n_assign = rows * K;
  const long long inter_max = std::max<long long>({(long long)c.inter, S, (long long)mi * K});
  void* xn16 = ws("dec_xn16", 2 * rows * H * 2).p;  // written by the norm kernel, unused on this path
  float* xn = ws("dec_xn32", rows * H * 4).as<float>();
  float* qkv = ws("dec_qkv32", rows * 3 * H * 4).as<float>();
  float* q = ws("dec_q32", rows * H * 4).as<float>();
  float* ctx = ws("dsq_ctx32", rows * H * 4).as<float>();
  float* g32 = ws("dsq_gate32", rows * inter_max * 4).as<float>();
  float* u32 = ws("dsq_up32", rows * inter_max * 4).as<float>();
  float* h32 = ws("dsq_h32", rows * inter_max * 4).as<float>();
  float* y32 = ws("dsq_y32", n_assign * H * 4).as<float>();
  int* topk_idx = ws("moe_topk_idx", n_assign * 4).as<int>();
  float* topk_w = ws("moe_topk_w", n_assign * 4).as<float>();
  int* counts = ws("moe_counts", 3 * E * 4 + 16).as<int>();
  int* iota = ws("dsq_iota", n_assign * 4).as<int>();
  if (iota_n_ < n_assign) {  // identity slot map, uploaded once (never during graph capture: sizes repeat)
    std::vector<int> h(n_assign);
    for (long long i = 0; i < n_assign; ++i) h[i] = (int)i;
    cuda_check(cudaMemcpyAsync(iota, h.data(), n_assign * 4, cudaMemcpyHostToDevice, stream_), "iota upload");
    cuda_check(cudaStreamSynchronize(stream_), "iota sync");
    iota_n_ = n_assign;
  }
  auto gemv = [&](const QuantWeight& w, const float* xin, long long ldx, float* out, long long ldo, long long nrows,
                  bool acc, const char* tag, const int* row_expert = nullptr, int x_row_div = 1) {
    DsqGemvCall gc;
    gc.w = &w; gc.x = xin; gc.ldx = ldx; gc.out = out; gc.ldo = ldo; gc.rows = nrows; gc.accumulate = acc; gc.tag = tag;
    gc.row_expert = row_expert; gc.x_row_div = x_row_div;
    dsq_gemv(gc, stream_);
  };
  for (int l = 0; l < c.layers; ++l) {
    DecLayerW& L = dec_[l];
    rmsnorm_split(x, L.ln1.as<float>(), xn16, rows * H, xn, nullptr, rows, H, c.rms_eps, nullptr, 0, 0, dt_, stream_);
    gemv(L.q_q, xn, H, qkv, 3 * H, rows, false, "dsq_q_proj");
    gemv(L.q_k, xn, H, qkv + H, 3 * H, rows, false, "dsq_k_proj");
    gemv(L.q_v, xn, H, qkv + 2 * H, 3 * H, rows, false, "dsq_v_proj");
    rope_kv(qkv, rope_cos_.as<float>(), rope_sin_.as<float>(), row_page, row_pos, q, kc_ptr(l), vc_ptr(l), kv_f16_,
            rows, heads, smax, 1, 0, stream_);
    kv_attention(q, kc_ptr(l), vc_ptr(l), kv_f16_, row_page, row_pos, nullptr, 0, ctx, rows, heads, smax, scale, dt_, stream_);
    gemv(L.q_o, ctx, H, x, H, rows, true, "dsq_o_proj");
    rmsnorm_split(x, L.ln2.as<float>(), xn16, rows * H, xn, nullptr, rows, H, c.rms_eps, nullptr, 0, 0, dt_, stream_);
    if (!L.moe) {
      gemv(L.q_gate, xn, H, g32, c.inter, rows, false, "dsq_dense_gate");
      gemv(L.q_up, xn, H, u32, c.inter, rows, false, "dsq_dense_up");
      swiglu_f32(g32, u32, h32, rows * c.inter, stream_);
      gemv(L.q_down, h32, c.inter, x, H, rows, true, "dsq_dense_down");
    } else {
      cuda_check(cudaMemsetAsync(counts, 0, E * 4, stream_), "moe counts memset");
      moe_router(xn, L.router_wt.as<float>(), topk_idx, topk_w, counts, rows, H, E, K, stream_);
      gemv(L.q_exp_gate, xn, H, g32, mi, n_assign, false, "dsq_expert_gate", topk_idx, K);
      gemv(L.q_exp_up, xn, H, u32, mi, n_assign, false, "dsq_expert_up", topk_idx, K);
      swiglu_f32(g32, u32, h32, n_assign * mi, stream_);
      gemv(L.q_exp_down, h32, mi, y32, H, n_assign, false, "dsq_expert_down", topk_idx, 1);
      moe_combine(y32, iota, topk_w, x, rows, K, H, nullptr, 0, 0, stream_);
      gemv(L.q_sh_gate, xn, H, g32, S, rows, false, "dsq_shared_gate");
      gemv(L.q_sh_up, xn, H, u32, S, rows, false, "dsq_shared_up");
      swiglu_f32(g32, u32, h32, rows * S, stream_);
      gemv(L.q_sh_down, h32, S, x, H, rows, true, "dsq_shared_down");
    }
  }
  float* xf = ws("dsq_xf32", (size_t)n_final * H * 4).as<float>();
  void* xf16 = ws("dec_xf16", 2 * (size_t)n_final * H * 2).p;
  rmsnorm_split(x, final_norm_.as<float>(), xf16, (long long)n_final * H, xf, final_rows, n_final, H, c.rms_eps, nullptr, 0, 0, dt_, stream_);
  gemv(q_lm_head_, xf, H, logits, c.vocab, n_final, false, "dsq_lm_head");
}

// Decode step for <= 4 pages with 6 launches per layer (see dsq_decode.cu): f32 activations against either the DSQ
// snapshot's quantised blocks or the float engine's pre-tiled 16-bit weights.  The residual stream ping-pongs between
// x (the embedding rows on entry) and x1: a kernel that folds pending adds into the residual while staging its
// activations writes the sum to the other buffer, which no block of that launch reads.
void Engine::decoder_step_fused_small(float* x, long long rows, const int* row_page, const int* row_pos, int smax, float* logits) {
  const ModelConfig& c = cfg_;
  const int H = c.hidden, heads = c.heads, E = c.n_experts, K = c.topk, mi = c.moe_inter;
  const long long S = (long long)mi * c.n_shared;
  const float scale = 1.0f / sqrtf((float)c.head_dim());
  const long long na = rows * K;
  const int R = (int)rows;
  const bool bf = dt_ == DType::BF16;
  auto W = [&](const QuantWeight& q, const DevBuf& t, long long N, int Kin) {
    return quantized_ ? fused_weight(q) : fused_weight_tiled16(t.p, N, Kin, bf);
  };
  float* x1 = ws("dsqf_x1", rows * H * 4).as<float>();
  float* qkv = ws("dec_qkv32", rows * 3 * H * 4).as<float>();
  float* ctx = ws("dsq_ctx32", rows * H * 4).as<float>();
  float* o32 = ws("dsqf_o32", rows * H * 4).as<float>();
  float* xn = ws("dec_xn32", rows * H * 4).as<float>();
  float* h = ws("dsqf_h32", rows * std::max<long long>(c.inter, (long long)K * mi) * 4).as<float>();
  float* hs = ws("dsqf_hs32", rows * S * 4).as<float>();
  float* y = ws("dsq_y32", na * H * 4).as<float>();
  float* ysh = ws("dsqf_ysh32", rows * H * 4).as<float>();
  float* d32 = ws("dsqf_d32", rows * H * 4).as<float>();
  int* topk_idx = ws("moe_topk_idx", na * 4).as<int>();
  float* topk_w = ws("moe_topk_w", na * 4).as<float>();
  const int nsplit = dsq_attn_splits(smax);
  float* part = ws("dsqf_attn_part", dsq_attn_ws_floats(rows, heads, nsplit) * 4).as<float>();
  int* counters = ws("dsqf_attn_cnt", (size_t)rows * heads * 4).as<int>();
  float* router_ws = ws("dsqf_router_logits", (size_t)rows * E * 4).as<float>();
  int* router_cnt = ws("dsqf_router_cnt", (size_t)rows * 4).as<int>();

  float* cur = x;    // residual as of the last kernel that wrote it
  float* alt = x1;
  DsqFusedStage pend;  // adds still to be folded into the residual by the next staging kernel
  for (int l = 0; l < c.layers; ++l) {
    DecLayerW& L = dec_[l];
    {  // pending adds + RMSNorm(ln1) + q | k | v
      DsqFusedStage st = pend;
      st.write_back = alt; st.norm_w = L.ln1.as<float>(); st.eps = c.rms_eps;
      DsqFusedJob j[3];
      int nj = 1;
      if (quantized_) {
        const QuantWeight* w[3] = {&L.q_q, &L.q_k, &L.q_v};
        for (int i = 0; i < 3; ++i) { j[i].w0 = fused_weight(*w[i]); j[i].out = qkv + (long long)i * H; }
        nj = 3;
      } else {
        j[0].w0 = fused_weight_tiled16(L.qkv_w.p, 3 * H, H, bf); j[0].out = qkv;
      }
      for (int i = 0; i < nj; ++i) { j[i].x = cur; j[i].ldx = H; j[i].rpg = R; j[i].ldo = 3 * H; }
      dsq_fused_gemv(j, nj, st, "fs_qkv", stream_);
      std::swap(cur, alt);
      pend = DsqFusedStage();
    }
    dsq_attn_split(qkv, rope_cos_.as<float>(), rope_sin_.as<float>(), kc_ptr(l), vc_ptr(l), kv_f16_, row_page, row_pos,
                   part, counters, ctx, rows, heads, c.head_dim(), smax, scale, nsplit, stream_);
    {
      DsqFusedJob j;
      j.w0 = W(L.q_o, L.o_w, H, H); j.x = ctx; j.ldx = H; j.rpg = R; j.out = o32; j.ldo = H;
      dsq_fused_gemv(&j, 1, DsqFusedStage(), "fs_o_proj", stream_);
    }
    if (!L.moe) {
      {  // residual + o_proj, RMSNorm(ln2), gate/up + SwiGLU
        DsqFusedStage st;
        st.add1 = o32; st.write_back = alt; st.norm_w = L.ln2.as<float>(); st.eps = c.rms_eps;
        DsqFusedJob j;
        j.w0 = W(L.q_gate, L.gate_w, c.inter, H); j.w1 = W(L.q_up, L.up_w, c.inter, H);
        j.x = cur; j.ldx = H; j.rpg = R; j.out = h; j.ldo = c.inter;
        dsq_fused_gemv(&j, 1, st, "fs_dense_gate_up", stream_);
        std::swap(cur, alt);
      }
      DsqFusedJob j;
      j.w0 = W(L.q_down, L.down_w, H, c.inter); j.x = h; j.ldx = c.inter; j.rpg = R; j.out = d32; j.ldo = H;
      dsq_fused_gemv(&j, 1, DsqFusedStage(), "fs_dense_down", stream_);
      pend.add1 = d32;
    } else {
      dsq_router(cur, o32, alt, L.ln2.as<float>(), L.router_wt.as<float>(), xn, router_ws, router_cnt, topk_idx, topk_w, rows, H, E, K,
                 c.rms_eps, stream_);
      std::swap(cur, alt);
      {  // routed experts (one group per (token, slot)) + shared experts: gate/up + SwiGLU
        DsqFusedJob j[2];
        j[0].w0 = W(L.q_exp_gate, L.exp_gate, mi, H); j[0].w1 = W(L.q_exp_up, L.exp_up, mi, H);
        j[0].x = xn; j[0].ldx = H; j[0].groups = (int)na; j[0].rpg = 1;
        j[0].x_row_div = K; j[0].row_expert = topk_idx; j[0].expert_dep = true; j[0].out = h; j[0].ldo = mi;
        j[1].w0 = W(L.q_sh_gate, L.sh_gate, S, H); j[1].w1 = W(L.q_sh_up, L.sh_up, S, H);
        j[1].x = xn; j[1].ldx = H; j[1].rpg = R; j[1].out = hs; j[1].ldo = S;
        dsq_fused_gemv(j, 2, DsqFusedStage(), "fs_moe_gate_up", stream_);
      }
      {
        DsqFusedJob j[2];
        j[0].w0 = W(L.q_exp_down, L.exp_down, H, mi); j[0].x = h; j[0].ldx = mi; j[0].groups = (int)na; j[0].rpg = 1;
        j[0].row_expert = topk_idx; j[0].out = y; j[0].ldo = H;
        j[1].w0 = W(L.q_sh_down, L.sh_down, H, (int)S); j[1].x = hs; j[1].ldx = S; j[1].rpg = R; j[1].out = ysh; j[1].ldo = H;
        dsq_fused_gemv(j, 2, DsqFusedStage(), "fs_moe_down", stream_);
      }
      pend.ymoe = y; pend.wmoe = topk_w; pend.topk = K; pend.add1 = ysh;
    }
  }
  // last layer's combine + final RMSNorm once (the lm_head blocks would each redo it), then the lm_head GEMV
  dsq_combine_norm(cur, pend.ymoe, pend.wmoe, pend.topk, pend.add1, final_norm_.as<float>(), xn, rows, H, c.rms_eps, stream_);
  DsqFusedJob j;
  j.w0 = W(q_lm_head_, lm_head_, c.vocab, H); j.x = xn; j.ldx = H; j.rpg = R; j.out = logits; j.ldo = c.vocab;
  dsq_fused_gemv(&j, 1, DsqFusedStage(), "fs_lm_head", stream_);
}

void Engine::generate(const GenRequest& rq, int64_t* const* out_tokens, int* n_out) {
  const ModelConfig& c = cfg_;
  const int P = rq.n_pages, H = c.hidden, V = c.vocab;
  if (P <= 0) return;
  const bool forced = rq.forced != nullptr;
  // sampling.rs:62: sampling needs do_sample and temperature > 0, everything else is the (device) argmax path
  const bool host_sample = !forced && rq.params.do_sample && rq.params.temperature > 0.0;
  const float penalty = rq.params.repetition_penalty;
  const int max_new = forced ? rq.n_forced_steps : (int)rq.params.max_new_tokens;
  for (int p = 0; p < P; ++p) n_out[p] = 0;
  if (max_new == 0) return;  // model/mod.rs:1889-1897

  cudaEvent_t ev0, ev1, ev2;
  cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventCreate(&ev2);
  cuda_check(cudaEventRecord(ev0, stream_), "event");

  // ---- host-side prompt bookkeeping
  long long total_rows = 0, total_img = 0;
  int max_T = 0;
  for (int p = 0; p < P; ++p) { total_rows += rq.n_tokens[p]; total_img += rq.n_image_rows[p]; max_T = std::max(max_T, rq.n_tokens[p]); }
  const int smax = (max_T + max_new + 15) / 16 * 16;
  if (smax > rope_len_) throw std::runtime_error("sequence length exceeds max_position_embeddings");
  std::vector<int> src(total_rows), row_page(total_rows), row_pos(total_rows), last_rows(P), hist((size_t)P * smax, 0), hist_len(P);
  // The prompts are prefilled in chunks of whole pages (<= kPrefillRows rows each, so the activation workspaces stay
  // bounded however many pages a call carries); row_page / last_rows are relative to the chunk a page belongs to.
  constexpr long long kPrefillRows = 32768;
  struct Chunk { int page0, n_pages; long long row0, rows; };
  std::vector<Chunk> chunks;
  {
    long long r = 0, img_base = 0;
    for (int p = 0; p < P; ++p) {
      if (chunks.empty() || chunks.back().rows + rq.n_tokens[p] > kPrefillRows) chunks.push_back({p, 0, r, 0});
      Chunk& ch = chunks.back();
      ch.n_pages++; ch.rows += rq.n_tokens[p];
      int img_used = 0;
      for (int t = 0; t < rq.n_tokens[p]; ++t, ++r) {
        const int64_t id = rq.input_ids[p][t];
        if (id < 0 || id >= V) throw std::runtime_error("token id out of range");
        hist[(size_t)p * smax + t] = (int)id;
        row_page[r] = p - ch.page0; row_pos[r] = t;
        if (rq.mask && rq.mask[p] && rq.mask[p][t]) src[r] = -(int)(img_base + img_used++) - 1;
        else src[r] = (int)id;
      }
      if (img_used != rq.n_image_rows[p])
        throw std::runtime_error("image embeddings provide " + std::to_string(rq.n_image_rows[p]) +
                                 " tokens but mask requires " + std::to_string(img_used));
      img_base += rq.n_image_rows[p];
      last_rows[p] = (int)(r - ch.row0) - 1;
      hist_len[p] = rq.n_tokens[p];
    }
  }
  // ---- device state
  const float* img_rows = rq.image_rows_dev;
  if (!img_rows && total_img > 0) {
    float* buf = ws("gen_img_rows32", total_img * H * 4).as<float>();
    long long off = 0;
    for (int p = 0; p < P; ++p) {
      if (rq.n_image_rows[p] == 0) continue;
      cuda_check(cudaMemcpyAsync(buf + off * H, rq.image_rows_host[p], (size_t)rq.n_image_rows[p] * H * 4, cudaMemcpyHostToDevice, stream_), "image rows upload");
      off += rq.n_image_rows[p];
    }
    img_rows = buf;
  }
  const size_t kv_bytes = (size_t)P * c.heads * smax * c.head_dim() * (kv_f16_ ? 2 : 4);
  kcache_.resize(c.layers); vcache_.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) { kcache_[l].ensure(kv_bytes); vcache_[l].ensure(kv_bytes); }
  kv_page_bytes_ = (size_t)c.heads * smax * c.head_dim() * (kv_f16_ ? 2 : 4);
  kv_page_base_ = 0;

  long long max_chunk_rows = P;
  for (const Chunk& ch : chunks) max_chunk_rows = std::max(max_chunk_rows, ch.rows);
  const long long max_rows = std::max<long long>(total_rows, P);
  int* d_src = ws("gen_src", max_rows * 4).as<int>();
  int* d_row_page = ws("gen_row_page", max_rows * 4).as<int>();
  int* d_row_pos = ws("gen_row_pos", max_rows * 4).as<int>();
  int* d_last = ws("gen_last_rows", P * 4).as<int>();
  int* d_hist = ws("gen_hist", (size_t)P * smax * 4).as<int>();
  int* d_state = ws("gen_state", (size_t)P * 3 * 4).as<int>();  // hist_len | gen_count | finished
  int* d_hist_len = d_state, *d_gen_count = d_state + P, *d_finished = d_state + 2 * P;
  float* x = ws("gen_x32", max_chunk_rows * H * 4).as<float>();
  float* logits = ws("gen_logits32", (size_t)P * V * 4).as<float>();
  float* sel_scratch = ws("gen_select_scratch", (size_t)P * kSelectScratchPerPage * 4).as<float>();
  cuda_check(cudaMemsetAsync(sel_scratch, 0, (size_t)P * kSelectScratchPerPage * 4, stream_), "select scratch memset");
  int* d_forced = nullptr, *d_selected = nullptr;
  if (forced) {
    std::vector<int> f((size_t)P * max_new);
    for (int p = 0; p < P; ++p) for (int s = 0; s < max_new; ++s) f[(size_t)p * max_new + s] = (int)rq.forced[p][s];
    d_forced = ws("gen_forced", f.size() * 4).as<int>();
    d_selected = ws("gen_selected", f.size() * 4).as<int>();
    cuda_check(cudaMemcpyAsync(d_forced, f.data(), f.size() * 4, cudaMemcpyHostToDevice, stream_), "forced upload");
    cuda_check(cudaStreamSynchronize(stream_), "forced sync");
  }
  auto up = [&](int* d, const std::vector<int>& h) {
    cuda_check(cudaMemcpyAsync(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice, stream_), "prompt upload");
  };
  up(d_src, src); up(d_row_page, row_page); up(d_row_pos, row_pos); up(d_last, last_rows); up(d_hist, hist); up(d_hist_len, hist_len);
  cuda_check(cudaMemsetAsync(d_gen_count, 0, (size_t)P * 2 * 4, stream_), "state memset");
  cuda_check(cudaStreamSynchronize(stream_), "prompt upload sync");

  const int ngram = (int)rq.params.no_repeat_ngram_size;
  const int eos = forced ? -1 : (int)rq.params.eos_token_id;
  std::vector<float> h_logits;
  auto copy_logits = [&](int step) {
    if (!rq.logits_out) return;
    h_logits.resize((size_t)P * V);
    cuda_check(cudaMemcpyAsync(h_logits.data(), logits, h_logits.size() * 4, cudaMemcpyDeviceToHost, stream_), "logits D2H");
    cuda_check(cudaStreamSynchronize(stream_), "logits sync");
    for (int p = 0; p < P; ++p)
      if (rq.logits_out[p]) memcpy(rq.logits_out[p] + (size_t)step * V, h_logits.data() + (size_t)p * V, (size_t)V * 4);
  };

  // ---- prefill (model/mod.rs:1925-1947)
  kernel_timing_phase("prefill/");
  for (const Chunk& ch : chunks) {
    kv_page_base_ = ch.page0;
    embed_gather(d_src + ch.row0, embed_.p, img_rows, x, ch.rows, H, dt_, stream_);
    if (quantized_ && !dsq_gemm_) decoder_forward_dsq(x, ch.rows, d_row_page + ch.row0, d_row_pos + ch.row0, smax, d_last + ch.page0, ch.n_pages, logits + (size_t)ch.page0 * V);
    else decoder_forward(x, ch.rows, d_row_page + ch.row0, d_row_pos + ch.row0, smax, d_last + ch.page0, ch.n_pages, logits + (size_t)ch.page0 * V, false);
  }
  kv_page_base_ = 0;
  copy_logits(0);
  // host sampling state (do_sample): one RNG per page, seeded like the reference's single-page call (model/mod.rs:1917)
  const SamplingParams sp = sampling_params_of(rq.params);
  std::vector<StdRng> rngs;
  std::vector<std::vector<int64_t>> ctxs;
  std::vector<int> h_chosen, h_count, h_done;
  std::vector<float> h_rows;
  int* d_chosen = nullptr;
  if (host_sample) {
    for (int p = 0; p < P; ++p) {
      rngs.push_back(rq.params.has_seed ? StdRng(rq.params.seed) : StdRng::from_entropy());
      ctxs.emplace_back(rq.input_ids[p], rq.input_ids[p] + rq.n_tokens[p]);
    }
    h_chosen.assign(P, 0); h_count.assign(P, 0); h_done.assign(P, 0);
    h_rows.resize((size_t)P * V);
    d_chosen = ws("gen_chosen", (size_t)P * 4).as<int>();
  }
  auto select = [&]() {
    if (!host_sample) {
      select_token(logits, V, d_hist, smax, d_hist_len, d_gen_count, d_finished, P, ngram, penalty, eos, max_new, d_forced,
                   max_new, d_selected, max_new, sel_scratch, stream_);
      return;
    }
    cuda_check(cudaMemcpyAsync(h_rows.data(), logits, h_rows.size() * 4, cudaMemcpyDeviceToHost, stream_), "logits D2H");
    cuda_check(cudaStreamSynchronize(stream_), "logits sync");
    for (int p = 0; p < P; ++p) {
      if (h_done[p]) continue;
      const int64_t t = select_token_id(h_rows.data() + (size_t)p * V, (size_t)V, sp, ctxs[p].data(), ctxs[p].size(), rngs[p]);
      h_chosen[p] = (int)t;
      if (eos >= 0 && t == eos) { h_done[p] = 1; continue; }
      ctxs[p].push_back(t);
      if (++h_count[p] >= max_new) h_done[p] = 1;
    }
    cuda_check(cudaMemcpyAsync(d_chosen, h_chosen.data(), (size_t)P * 4, cudaMemcpyHostToDevice, stream_), "chosen H2D");
    append_tokens(d_chosen, d_hist, smax, d_hist_len, d_gen_count, d_finished, P, eos, max_new, stream_);
    cuda_check(cudaStreamSynchronize(stream_), "chosen sync");  // h_chosen is reused by the next step
  };
  select();
  cuda_check(cudaEventRecord(ev1, stream_), "event");

  // ---- token loop (model/mod.rs:1977-2034): every page advances one token per step
  std::vector<int> page_ids(P), fin(P), h_hist, h_len(P), delivered(P, 0);
  for (int p = 0; p < P; ++p) page_ids[p] = p;
  up(d_row_page, page_ids);
  // with a callback every step is followed by a sync so that tokens are delivered one at a time, in the order the
  // reference's loop would (model/mod.rs:1978-1982); without one the "all pages finished" test runs every 16 steps
  const int sync_every = (rq.cb || host_sample) ? 1 : 16;
  auto deliver = [&]() {  // streaming callback: (count, all generated ids) after every accepted token
    if (!rq.cb) return;
    h_hist.resize((size_t)P * smax);
    d2h(h_hist.data(), d_hist, h_hist.size() * 4);
    d2h(h_len.data(), d_hist_len, P * 4);
    std::vector<int64_t> toks;
    for (int p = 0; p < P; ++p) {
      const int gen = h_len[p] - rq.n_tokens[p];
      for (int cnt = delivered[p] + 1; cnt <= gen; ++cnt) {
        toks.resize(cnt);
        for (int i = 0; i < cnt; ++i) toks[i] = h_hist[(size_t)p * smax + rq.n_tokens[p] + i];
        rq.cb(rq.user, p + rq.page_offset, (size_t)cnt, toks.data());
      }
      delivered[p] = std::max(delivered[p], gen);
    }
  };
  kernel_timing_phase("decode/");
  // One decode step = ~170 kernels whose arguments do not change between steps (positions, token history and
  // routing live in device memory), so after one eager step (which sizes every workspace) the step is
  // captured once into a CUDA graph and replayed.
  auto run_step = [&](int step) {
    decode_rows(d_hist, smax, d_hist_len, d_src, d_row_pos, P, stream_);
    embed_gather(d_src, embed_.p, nullptr, x, P, H, dt_, stream_);
    if (P <= 4 && small_fused_ && (quantized_ || w_tiled_) && !record_taps_ && ep_peers_.world == 1) decoder_step_fused_small(x, P, d_row_page, d_row_pos, smax, logits);
    else if (quantized_ && !dsq_gemm_) decoder_forward_dsq(x, P, d_row_page, d_row_pos, smax, d_row_page, P, logits);
    else decoder_forward(x, P, d_row_page, d_row_pos, smax, d_row_page /* identity: every row */, P, logits, true);
    copy_logits(step);
    select();
  };
  bool use_graph = !kernel_timing_enabled() && !rq.logits_out && !record_taps_ && !host_sample && !getenv("DSOCR_NO_GRAPH");
  if (use_graph && (stream_ == nullptr || stream_ == cudaStreamLegacy || stream_ == cudaStreamPerThread))
    use_graph = false;  // the default streams cannot be captured
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  long long graph_launches = 0;
  for (int step = 1; step < max_new; ++step) {
    if ((step - 1) % sync_every == 0) {
      cuda_check(cudaStreamSynchronize(stream_), "decode sync");
      d2h(fin.data(), d_finished, P * 4);
      deliver();
      bool all = true;
      for (int p = 0; p < P; ++p) all &= fin[p] != 0;
      // expert parallel: the ranks of a group step in lock-step (their barriers pair up), so nobody leaves early
      if (all && ep_peers_.world == 1) break;
    }
    if (!use_graph || step == 1) {
      run_step(step);
    } else if (!graph_exec) {
      const long long before = launch_counter().load();
      cuda_check(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal), "graph capture begin");
      try {
        run_step(step);
      } catch (...) {
        cudaStreamEndCapture(stream_, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      cuda_check(cudaStreamEndCapture(stream_, &graph), "graph capture end");
      graph_launches = launch_counter().load() - before;
      cuda_check(cudaGraphInstantiate(&graph_exec, graph, 0), "graph instantiate");
      cuda_check(cudaGraphLaunch(graph_exec, stream_), "graph launch");
    } else {
      cuda_check(cudaGraphLaunch(graph_exec, stream_), "graph launch");
      launch_counter() += graph_launches;
    }
  }
  cuda_check(cudaEventRecord(ev2, stream_), "event");
  cuda_check(cudaStreamSynchronize(stream_), "generate sync");
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (graph) cudaGraphDestroy(graph);
  deliver();

  // ---- results
  h_hist.resize((size_t)P * smax);
  d2h(h_hist.data(), d_hist, h_hist.size() * 4);
  d2h(h_len.data(), d_hist_len, P * 4);
  std::vector<int> h_sel;
  if (forced) { h_sel.resize((size_t)P * max_new); d2h(h_sel.data(), d_selected, h_sel.size() * 4); }
  for (int p = 0; p < P; ++p) {
    if (forced) {
      n_out[p] = max_new;
      for (int s = 0; s < max_new; ++s) out_tokens[p][s] = h_sel[(size_t)p * max_new + s];
    } else {
      const int gen = h_len[p] - rq.n_tokens[p];
      n_out[p] = gen;
      for (int i = 0; i < gen; ++i) out_tokens[p][i] = h_hist[(size_t)p * smax + rq.n_tokens[p] + i];
    }
  }
  float ms_prefill = 0, ms_iter = 0;
  cudaEventElapsedTime(&ms_prefill, ev0, ev1);
  cudaEventElapsedTime(&ms_iter, ev1, ev2);
  timings.prefill = ms_prefill; timings.iterative = ms_iter; timings.generate = ms_prefill + ms_iter;
  cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
}

}  // namespace dsocr
