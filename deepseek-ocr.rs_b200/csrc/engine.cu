// Engine implementation, part 1: configuration, checkpoint upload and the vision path
// (SAM ViT-B -> CLIP-L -> projector -> token layout).  See engine.h.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <thread>

#include "hostmath.h"
#include "json.h"
#include "linear_tc.cuh"  // lin::Act / lin::Out enums
#include "safetensors.h"

namespace dsocr {

// ------------------------------------------------------------------------------------------------ config
static ModelConfig parse_config(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("failed to read config file " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  Json root = JsonParser(text.data(), text.size()).parse();
  ModelConfig c;
  // language_config merged over top-level defaults (config/mod.rs:70-92)
  auto lang = [&](const char* key) -> const Json* {
    if (const Json* lc = root.get("language_config"))
      if (const Json* v = lc->get(key)) if (!v->is_null()) return v;
    const Json* v = root.get(key);
    return (v && !v->is_null()) ? v : nullptr;
  };
  auto geti = [&](const char* k, int d) { const Json* v = lang(k); return v ? (int)v->as_int(d) : d; };
  auto getf = [&](const char* k, float d) { const Json* v = lang(k); return v ? (float)v->as_num(d) : d; };
  c.vocab = geti("vocab_size", c.vocab);
  c.hidden = geti("hidden_size", c.hidden);
  c.layers = geti("num_hidden_layers", c.layers);
  c.heads = geti("num_attention_heads", c.heads);
  c.inter = geti("intermediate_size", c.inter);
  c.moe_inter = geti("moe_intermediate_size", c.moe_inter);
  c.n_experts = geti("n_routed_experts", c.n_experts);
  c.n_shared = geti("n_shared_experts", c.n_shared);
  c.topk = geti("num_experts_per_tok", c.topk);
  c.first_dense = geti("first_k_dense_replace", c.first_dense);
  c.rope_theta = getf("rope_theta", c.rope_theta);
  c.rms_eps = getf("rms_norm_eps", c.rms_eps);
  c.eos = geti("eos_token_id", c.eos);
  if (lang("q_lora_rank") || lang("kv_lora_rank"))
    throw std::runtime_error("LoRA attention path not yet implemented");  // block.rs:452-454
  if (geti("num_key_value_heads", c.heads) != c.heads) throw std::runtime_error("GQA decoder is not supported");
  if (const Json* v = lang("use_mla")) if (v->as_bool(false)) throw std::runtime_error("use_mla=true is not supported");
  if (const Json* v = lang("norm_topk_prob")) if (v->as_bool(false)) throw std::runtime_error("norm_topk_prob=true is not supported");
  if (const Json* v = lang("scoring_func")) if (v->kind == Json::Str && v->str != "softmax") throw std::runtime_error("MoE scoring `" + v->str + "` not yet supported");
  // run_moe options this engine's router / combine kernels do not implement: fail at load instead of computing wrong numbers
  if (const Json* v = lang("routed_scaling_factor")) if (fabs(v->as_num(1.0) - 1.0) > 1e-6) throw std::runtime_error("routed_scaling_factor != 1 is not supported");
  if (const Json* v = lang("topk_method")) if (v->kind == Json::Str && v->str != "greedy") throw std::runtime_error("MoE topk_method `" + v->str + "` is not supported");
  if (geti("moe_layer_freq", 1) != 1) throw std::runtime_error("moe_layer_freq != 1 is not supported");
  if (const Json* vc = root.get("vision_config")) {
    if (const Json* w = vc->get("width")) {
      if (const Json* s = w->get("sam_vit_b")) {
        if (const Json* v = s->get("width")) c.sam_dim = (int)v->as_int(c.sam_dim);
        if (const Json* v = s->get("layers")) c.sam_depth = (int)v->as_int(c.sam_depth);
        if (const Json* v = s->get("heads")) c.sam_heads = (int)v->as_int(c.sam_heads);
        if (const Json* v = s->get("patch_size")) c.sam_patch = (int)v->as_int(c.sam_patch);
        if (const Json* v = s->get("image_size")) c.sam_image = (int)v->as_int(c.sam_image);
        if (const Json* v = s->get("global_attn_indexes")) if (v->kind == Json::Arr) {
          c.sam_global.clear();
          for (auto& e : v->arr) c.sam_global.push_back((int)e.num);
        }
        if (const Json* v = s->get("downsample_channels")) if (v->kind == Json::Arr && v->arr.size() == 2) {
          c.sam_out0 = (int)v->arr[0].num; c.sam_out1 = (int)v->arr[1].num;
        }
      }
      if (const Json* s = w->get("clip-l-14-224")) {
        if (const Json* v = s->get("width")) c.clip_dim = (int)v->as_int(c.clip_dim);
        if (const Json* v = s->get("layers")) c.clip_layers = (int)v->as_int(c.clip_layers);
        if (const Json* v = s->get("heads")) c.clip_heads = (int)v->as_int(c.clip_heads);
        if (const Json* v = s->get("patch_size")) c.clip_patch = (int)v->as_int(c.clip_patch);
        if (const Json* v = s->get("image_size")) c.clip_image = (int)v->as_int(c.clip_image);
      }
      if (w->get("qwen2-0-5b")) throw std::runtime_error("DeepSeek-OCR-2 (qwen2 vision) is not supported by this engine");
    }
    if (const Json* v = vc->get("image_size")) if (!v->is_null()) c.sam_image = (int)v->as_int(c.sam_image);
  }
  if (const Json* pc = root.get("projector_config")) {
    if (const Json* v = pc->get("input_dim")) c.proj_in = (int)v->as_int(c.proj_in);
    if (const Json* v = pc->get("n_embed")) c.n_embed = (int)v->as_int(c.n_embed);
  }
  if (c.sam_dim != 768 || c.sam_heads != 12 || c.clip_dim != 1024 || c.clip_heads != 16 || c.sam_neck != 256)
    throw std::runtime_error("unsupported vision tower widths (kernels are specialised for SAM ViT-B / CLIP-L)");
  if (c.head_dim() != 128) throw std::runtime_error("decoder head_dim must be 128");
  if (c.hidden % 128 || c.moe_inter % 128 || c.inter % 64) throw std::runtime_error("unsupported decoder widths");
  if (c.n_embed != c.hidden) throw std::runtime_error("projector n_embed must equal the decoder hidden size");
  return c;
}

// ------------------------------------------------------------------------------------------------ upload
namespace {

inline float bf16_bits_to_f32(uint16_t u) { uint32_t v = (uint32_t)u << 16; float f; memcpy(&f, &v, 4); return f; }

void parallel_for(size_t n, const std::function<void(size_t, size_t)>& fn) {
  const size_t nthreads = n < (1u << 20) ? 1 : std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
  if (nthreads <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  const size_t chunk = (n + nthreads - 1) / nthreads;
  for (size_t t = 0; t < nthreads; ++t) {
    const size_t a = t * chunk, b = std::min(n, a + chunk);
    if (a < b) th.emplace_back([=, &fn] { fn(a, b); });
  }
  for (auto& t : th) t.join();
}

std::vector<float> to_f32(const StTensor& t) {
  const size_t n = (size_t)t.numel();
  std::vector<float> out(n);
  if (t.dtype == "F32") memcpy(out.data(), t.data, n * 4);
  else if (t.dtype == "BF16") {
    const uint16_t* p = (const uint16_t*)t.data;
    parallel_for(n, [&](size_t a, size_t b) { for (size_t i = a; i < b; ++i) out[i] = bf16_bits_to_f32(p[i]); });
  } else if (t.dtype == "F16") {
    const uint16_t* p = (const uint16_t*)t.data;
    parallel_for(n, [&](size_t a, size_t b) { for (size_t i = a; i < b; ++i) out[i] = f16_to_32(p[i], DType::F16); });
  } else throw std::runtime_error("unsupported checkpoint dtype " + t.dtype);
  return out;
}

void upload_f32(DevBuf& dst, const std::vector<float>& v) {
  dst.alloc(v.size() * 4);
  h2d(dst.p, v.data(), v.size() * 4);
}

// tensor -> 16-bit device storage at element offset `off` of dst (dst must be allocated)
void upload16_into(void* dst, size_t off_elems, const StTensor& t, DType dt) {
  const size_t n = (size_t)t.numel();
  uint8_t* d = (uint8_t*)dst + off_elems * 2;
  const bool same = (t.dtype == "BF16" && dt == DType::BF16) || (t.dtype == "F16" && dt == DType::F16);
  if (same) { h2d(d, t.data, n * 2); return; }
  std::vector<uint16_t> tmp(n);
  if (t.dtype == "F32") {
    const float* p = (const float*)t.data;
    parallel_for(n, [&](size_t a, size_t b) { for (size_t i = a; i < b; ++i) tmp[i] = f32_to_16(p[i], dt); });
  } else {
    const uint16_t* p = (const uint16_t*)t.data;
    const bool src_bf16 = t.dtype == "BF16";
    if (!src_bf16 && t.dtype != "F16") throw std::runtime_error("unsupported checkpoint dtype " + t.dtype);
    parallel_for(n, [&](size_t a, size_t b) {
      for (size_t i = a; i < b; ++i)
        tmp[i] = f32_to_16(src_bf16 ? bf16_bits_to_f32(p[i]) : f16_to_32(p[i], DType::F16), dt);
    });
  }
  h2d(d, tmp.data(), n * 2);
}
void upload16(DevBuf& dst, const StTensor& t, DType dt) {
  dst.alloc((size_t)t.numel() * 2);
  upload16_into(dst.p, 0, t, dt);
}
void upload16_vec(DevBuf& dst, const std::vector<float>& v, DType dt) {
  auto h = to16(v.data(), v.size(), dt);
  dst.alloc(h.size() * 2);
  h2d(dst.p, h.data(), h.size() * 2);
}
void expect_shape(const StTensor& t, std::initializer_list<long long> shape, const std::string& name) {
  if (t.shape != std::vector<long long>(shape)) {
    std::string got;
    for (auto d : t.shape) got += std::to_string(d) + ",";
    throw std::runtime_error("tensor `" + name + "` has unexpected shape [" + got + "]");
  }
}
// conv weight [O, C, kh, kw] -> [O, (kh, kw, C)] to match the NHWC im2col column order
std::vector<float> conv_to_khwc(const StTensor& t) {
  const long long O = t.shape[0], C = t.shape[1], KH = t.shape[2], KW = t.shape[3];
  std::vector<float> src = to_f32(t), out(src.size());
  for (long long o = 0; o < O; ++o)
    for (long long c = 0; c < C; ++c)
      for (long long y = 0; y < KH; ++y)
        for (long long x = 0; x < KW; ++x)
          out[((o * KH + y) * KW + x) * C + c] = src[((o * C + c) * KH + y) * KW + x];
  return out;
}

}  // namespace

namespace {
// snapshot record -> QuantWeight rows [row0, row0 + out_dim)
void load_quant(QuantWeight& w, const DsqReader& dsq, const std::string& name, long long N, int K, int count, int idx) {
  const DsqRecord* r = dsq.find(name);
  if (!r) throw std::runtime_error("snapshot is missing tensor `" + name + "`");
  if ((long long)r->out_dim != N || (int)r->in_dim != K) throw std::runtime_error("snapshot tensor `" + name + "` has unexpected dims");
  if (r->has_bias) throw std::runtime_error("snapshot bias tensors are not supported (`" + name + "`)");
  // the reader (like DsqReader::open) only checks in_dim % block for block dtypes; the payload length is checked here,
  // where the reference would fail while building the QTensor (dsq-runtime/src/lib.rs:316-369)
  if (const int be = dsq_block_elems(r->q_dtype))
    if (r->q_len != (uint64_t)r->out_dim * (r->in_dim / be) * dsq_block_bytes(r->q_dtype))
      throw std::runtime_error("snapshot tensor `" + name + "` payload length does not match its dims");
  if (idx == 0) dsq_alloc(w, r->q_dtype, N, K, count);
  dsq_upload_rows(w, (long long)idx * N, dsq.bytes(*r), r->q_dtype, N);
}
}  // namespace

void Engine::retile_inplace(DevBuf& w, long long n, int k) {
  if (k % 64 || (n % 128 && w.bytes < (size_t)n * k * 2)) throw std::runtime_error("retile: unsupported weight shape");
  DevBuf t(retiled_bytes(n, k));
  retile_weights(w.p, t.p, n, k, 0);
  cuda_check(cudaDeviceSynchronize(), "weight retile");
  w = std::move(t);
}

void Engine::load_weights(const std::string& path, const DsqReader* dsq) {
  SafeTensors st(path);
  const ModelConfig& c = cfg_;
  auto T16 = [&](DevBuf& d, const std::string& n, std::initializer_list<long long> shape) {
    const StTensor& t = st.get(n); expect_shape(t, shape, n); upload16(d, t, dt_);
  };
  auto F32 = [&](DevBuf& d, const std::string& n, std::initializer_list<long long> shape) {
    const StTensor& t = st.get(n); expect_shape(t, shape, n); upload_f32(d, to_f32(t));
  };
  // ---- SAM (vision/sam.rs:143-184)
  const std::string s = "model.sam_model.";
  const long long D = c.sam_dim, g0 = c.sam_image / c.sam_patch;
  {
    const StTensor& t = st.get(s + "patch_embed.proj.weight");
    expect_shape(t, {D, 3, c.sam_patch, c.sam_patch}, "patch_embed.proj.weight");
    upload16(patch_w_, t, dt_);
    F32(patch_b_, s + "patch_embed.proj.bias", {D});
    const StTensor& p = st.get(s + "pos_embed");
    expect_shape(p, {1, g0, g0, D}, "pos_embed");
    sam_pos_host_ = to_f32(p);
  }
  sam_.resize(c.sam_depth);
  for (int i = 0; i < c.sam_depth; ++i) {
    const std::string p = s + "blocks." + std::to_string(i) + ".";
    SamBlockW& b = sam_[i];
    F32(b.ln1_w, p + "norm1.weight", {D}); F32(b.ln1_b, p + "norm1.bias", {D});
    F32(b.ln2_w, p + "norm2.weight", {D}); F32(b.ln2_b, p + "norm2.bias", {D});
    T16(b.qkv_w, p + "attn.qkv.weight", {3 * D, D}); F32(b.qkv_b, p + "attn.qkv.bias", {3 * D});
    {
      std::vector<float> w = to_f32(st.get(p + "attn.qkv.weight")), bq = to_f32(st.get(p + "attn.qkv.bias"));
      b.q_w_host.assign(w.begin(), w.begin() + (size_t)D * D);
      for (float& v : b.q_w_host) v = f16_to_32(f32_to_16(v, dt_), dt_);
      b.q_b_host.assign(bq.begin(), bq.begin() + D);
    }
    T16(b.proj_w, p + "attn.proj.weight", {D, D}); F32(b.proj_b, p + "attn.proj.bias", {D});
    const std::string m1 = st.has(p + "mlp.fc1.weight") ? "mlp.fc1" : "mlp.lin1";  // sam.rs:897-915
    const std::string m2 = st.has(p + "mlp.fc2.weight") ? "mlp.fc2" : "mlp.lin2";
    T16(b.fc1_w, p + m1 + ".weight", {4 * D, D}); F32(b.fc1_b, p + m1 + ".bias", {4 * D});
    T16(b.fc2_w, p + m2 + ".weight", {D, 4 * D}); F32(b.fc2_b, p + m2 + ".bias", {D});
    const long long rel = 2 * (c.sam_is_global(i) ? g0 : c.sam_window) - 1;
    const StTensor& rh = st.get(p + "attn.rel_pos_h");
    const StTensor& rw = st.get(p + "attn.rel_pos_w");
    expect_shape(rh, {rel, 64}, "rel_pos_h"); expect_shape(rw, {rel, 64}, "rel_pos_w");
    b.rel_h = to_f32(rh); b.rel_w = to_f32(rw); b.rel_rows = (int)rel;
  }
  T16(neck0_w_, s + "neck.0.weight", {c.sam_neck, D, 1, 1});
  F32(neck1_w_, s + "neck.1.weight", {c.sam_neck}); F32(neck1_b_, s + "neck.1.bias", {c.sam_neck});
  F32(neck3_w_, s + "neck.3.weight", {c.sam_neck}); F32(neck3_b_, s + "neck.3.bias", {c.sam_neck});
  {
    const StTensor& t2 = st.get(s + "neck.2.weight"); expect_shape(t2, {c.sam_neck, c.sam_neck, 3, 3}, "neck.2.weight");
    upload16_vec(neck2_w_, conv_to_khwc(t2), dt_);
    const StTensor& n2 = st.get(s + "net_2.weight"); expect_shape(n2, {c.sam_out0, c.sam_neck, 3, 3}, "net_2.weight");
    upload16_vec(net2_w_, conv_to_khwc(n2), dt_);
    const StTensor& n3 = st.get(s + "net_3.weight"); expect_shape(n3, {c.sam_out1, c.sam_out0, 3, 3}, "net_3.weight");
    upload16_vec(net3_w_, conv_to_khwc(n3), dt_);
  }
  // ---- CLIP (vision/clip.rs:73-88)
  const std::string v = "model.vision_model.";
  const long long C = c.clip_dim, npos = (long long)(c.clip_image / c.clip_patch) * (c.clip_image / c.clip_patch) + 1;
  F32(clip_cls_, v + "embeddings.class_embedding", {C});
  {
    const StTensor& t = st.get(v + "embeddings.position_embedding.weight");
    expect_shape(t, {npos, C}, "position_embedding.weight");
    clip_pos_host_ = to_f32(t);
  }
  F32(clip_preln_w_, v + "pre_layrnorm.weight", {C}); F32(clip_preln_b_, v + "pre_layrnorm.bias", {C});
  clip_.resize(c.clip_layers);
  for (int i = 0; i < c.clip_layers; ++i) {
    const std::string p = v + "transformer.layers." + std::to_string(i) + ".";
    ClipBlockW& b = clip_[i];
    F32(b.ln1_w, p + "layer_norm1.weight", {C}); F32(b.ln1_b, p + "layer_norm1.bias", {C});
    F32(b.ln2_w, p + "layer_norm2.weight", {C}); F32(b.ln2_b, p + "layer_norm2.bias", {C});
    T16(b.qkv_w, p + "self_attn.qkv_proj.weight", {3 * C, C}); F32(b.qkv_b, p + "self_attn.qkv_proj.bias", {3 * C});
    T16(b.out_w, p + "self_attn.out_proj.weight", {C, C}); F32(b.out_b, p + "self_attn.out_proj.bias", {C});
    T16(b.fc1_w, p + "mlp.fc1.weight", {4 * C, C}); F32(b.fc1_b, p + "mlp.fc1.bias", {4 * C});
    T16(b.fc2_w, p + "mlp.fc2.weight", {C, 4 * C}); F32(b.fc2_b, p + "mlp.fc2.bias", {C});
  }
  // ---- projector (model/mod.rs:263-307; note the upstream spelling `view_seperator`)
  T16(proj_w_, "model.projector.layers.weight", {c.n_embed, c.proj_in});
  F32(proj_b_, "model.projector.layers.bias", {c.n_embed});
  F32(newline_, "model.image_newline", {c.n_embed});
  F32(separator_, "model.view_seperator", {c.n_embed});
  // ---- decoder (transformer/weights.rs)
  const long long H = c.hidden;
  T16(embed_, "model.embed_tokens.weight", {c.vocab, H});
  F32(final_norm_, "model.norm.weight", {H});
  if (dsq) load_quant(q_lm_head_, *dsq, "lm_head.weight", c.vocab, (int)H, 1, 0);
  else T16(lm_head_, "lm_head.weight", {c.vocab, H});
  dec_.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    const std::string p = "model.layers." + std::to_string(i) + ".";
    DecLayerW& L = dec_[i];
    F32(L.ln1, p + "input_layernorm.weight", {H});
    F32(L.ln2, p + "post_attention_layernorm.weight", {H});
    L.moe = i >= c.first_dense;  // should_use_moe, weights.rs:609-619 with moe_layer_freq = 1
    if (dsq) {
      load_quant(L.q_q, *dsq, p + "self_attn.q_proj.weight", H, (int)H, 1, 0);
      load_quant(L.q_k, *dsq, p + "self_attn.k_proj.weight", H, (int)H, 1, 0);
      load_quant(L.q_v, *dsq, p + "self_attn.v_proj.weight", H, (int)H, 1, 0);
      load_quant(L.q_o, *dsq, p + "self_attn.o_proj.weight", H, (int)H, 1, 0);
      if (!L.moe) {
        load_quant(L.q_gate, *dsq, p + "mlp.gate_proj.weight", c.inter, (int)H, 1, 0);
        load_quant(L.q_up, *dsq, p + "mlp.up_proj.weight", c.inter, (int)H, 1, 0);
        load_quant(L.q_down, *dsq, p + "mlp.down_proj.weight", H, c.inter, 1, 0);
      } else {
        const long long E = c.n_experts, mi = c.moe_inter, S = (long long)c.moe_inter * c.n_shared;
        const StTensor& t = st.get(p + "mlp.gate.weight"); expect_shape(t, {E, H}, "mlp.gate.weight");
        std::vector<float> w = to_f32(t), wt((size_t)E * H);
        for (long long e = 0; e < E; ++e) for (long long k = 0; k < H; ++k) wt[k * E + e] = w[e * H + k];
        upload_f32(L.router_wt, wt);
        if (st.has(p + "mlp.gate.e_score_correction_bias")) throw std::runtime_error("router score-correction bias is not supported");
        for (long long e = 0; e < E; ++e) {
          const std::string q = p + "mlp.experts." + std::to_string(e) + ".";
          load_quant(L.q_exp_gate, *dsq, q + "gate_proj.weight", mi, (int)H, (int)E, (int)e);
          load_quant(L.q_exp_up, *dsq, q + "up_proj.weight", mi, (int)H, (int)E, (int)e);
          load_quant(L.q_exp_down, *dsq, q + "down_proj.weight", H, (int)mi, (int)E, (int)e);
        }
        load_quant(L.q_sh_gate, *dsq, p + "mlp.shared_experts.gate_proj.weight", S, (int)H, 1, 0);
        load_quant(L.q_sh_up, *dsq, p + "mlp.shared_experts.up_proj.weight", S, (int)H, 1, 0);
        load_quant(L.q_sh_down, *dsq, p + "mlp.shared_experts.down_proj.weight", H, (int)S, 1, 0);
      }
      continue;
    }
    L.qkv_w.alloc((size_t)3 * H * H * 2);
    const char* names[3] = {"q", "k", "v"};
    for (int j = 0; j < 3; ++j) {
      const std::string n = p + "self_attn." + names[j] + "_proj.weight";
      const StTensor& t = st.get(n); expect_shape(t, {H, H}, n);
      upload16_into(L.qkv_w.p, (size_t)j * H * H, t, dt_);
      if (st.has(p + "self_attn." + names[j] + "_proj.bias")) throw std::runtime_error("decoder attention bias is not supported");
    }
    T16(L.o_w, p + "self_attn.o_proj.weight", {H, H});
    if (!L.moe) {
      T16(L.gate_w, p + "mlp.gate_proj.weight", {c.inter, H});
      T16(L.up_w, p + "mlp.up_proj.weight", {c.inter, H});
      T16(L.down_w, p + "mlp.down_proj.weight", {H, c.inter});
    } else {
      const long long E = c.n_experts, mi = c.moe_inter, S = (long long)c.moe_inter * c.n_shared;
      {
        const StTensor& t = st.get(p + "mlp.gate.weight"); expect_shape(t, {E, H}, "mlp.gate.weight");
        std::vector<float> w = to_f32(t), wt((size_t)E * H);
        for (long long e = 0; e < E; ++e) for (long long k = 0; k < H; ++k) wt[k * E + e] = w[e * H + k];
        upload_f32(L.router_wt, wt);
        if (st.has(p + "mlp.gate.e_score_correction_bias")) throw std::runtime_error("router score-correction bias is not supported");
      }
      // routed experts stacked as groups 0..E-1; the shared experts are appended as groups E.. (each mi wide) so the
      // decode path runs them inside the same grouped GEMMs
      const long long Eg = E + c.n_shared;
      L.exp_gate.alloc((size_t)Eg * mi * H * 2); L.exp_up.alloc((size_t)Eg * mi * H * 2); L.exp_down.alloc((size_t)Eg * H * mi * 2);
      for (long long e = 0; e < E; ++e) {
        const std::string q = p + "mlp.experts." + std::to_string(e) + ".";
        const StTensor& tg = st.get(q + "gate_proj.weight"); expect_shape(tg, {mi, H}, q + "gate_proj.weight");
        const StTensor& tu = st.get(q + "up_proj.weight"); expect_shape(tu, {mi, H}, q + "up_proj.weight");
        const StTensor& td = st.get(q + "down_proj.weight"); expect_shape(td, {H, mi}, q + "down_proj.weight");
        upload16_into(L.exp_gate.p, (size_t)e * mi * H, tg, dt_);
        upload16_into(L.exp_up.p, (size_t)e * mi * H, tu, dt_);
        upload16_into(L.exp_down.p, (size_t)e * H * mi, td, dt_);
      }
      T16(L.sh_gate, p + "mlp.shared_experts.gate_proj.weight", {S, H});
      T16(L.sh_up, p + "mlp.shared_experts.up_proj.weight", {S, H});
      T16(L.sh_down, p + "mlp.shared_experts.down_proj.weight", {H, S});
      upload16_into(L.exp_gate.p, (size_t)E * mi * H, st.get(p + "mlp.shared_experts.gate_proj.weight"), dt_);
      upload16_into(L.exp_up.p, (size_t)E * mi * H, st.get(p + "mlp.shared_experts.up_proj.weight"), dt_);
      {
        // down_proj [H, S] -> per shared expert s the column block [H, mi] (y = sum_s W[:, s*mi:(s+1)*mi] h_s)
        const std::vector<float> w = to_f32(st.get(p + "mlp.shared_experts.down_proj.weight"));
        std::vector<float> blk((size_t)H * mi);
        for (long long sidx = 0; sidx < c.n_shared; ++sidx) {
          for (long long n = 0; n < H; ++n) memcpy(&blk[(size_t)n * mi], &w[(size_t)n * S + sidx * mi], (size_t)mi * 4);
          const auto h = to16(blk.data(), blk.size(), dt_);
          h2d((uint8_t*)L.exp_down.p + (size_t)(E + sidx) * H * mi * 2, h.data(), h.size() * 2);
        }
      }
    }
  }
  if (w_tiled_ && !dsq) {
    // decode streams every decoder weight once per step: store them as contiguous, pre-swizzled 16 KB tiles
    const long long E = c.n_experts + c.n_shared, mi = c.moe_inter, S = (long long)c.moe_inter * c.n_shared;
    retile_inplace(lm_head_, c.vocab, (int)H);
    for (DecLayerW& L : dec_) {
      retile_inplace(L.qkv_w, 3 * H, (int)H);
      retile_inplace(L.o_w, H, (int)H);
      if (!L.moe) {
        retile_inplace(L.gate_w, c.inter, (int)H);
        retile_inplace(L.up_w, c.inter, (int)H);
        retile_inplace(L.down_w, H, c.inter);
      } else {
        retile_inplace(L.exp_gate, E * mi, (int)H);
        retile_inplace(L.exp_up, E * mi, (int)H);
        retile_inplace(L.exp_down, E * H, (int)mi);
        retile_inplace(L.sh_gate, S, (int)H);
        retile_inplace(L.sh_up, S, (int)H);
        retile_inplace(L.sh_down, H, (int)S);
      }
    }
    cuda_check(cudaDeviceSynchronize(), "weight retile");
  }
  rope_len_ = 8192;
  std::vector<float> cs, sn;
  rope_tables(c.rope_theta, c.head_dim(), rope_len_, cs, sn);
  upload_f32(rope_cos_, cs);
  upload_f32(rope_sin_, sn);
}

// ------------------------------------------------------------------------------------------------ lifecycle
Engine::Engine(const std::string& config_path, const std::string& weights_path, const std::string& dsq_path, int device,
               DType dtype)
    : dt_(dtype), device_(device) {
  if (dtype != DType::F16 && dtype != DType::BF16)
    throw std::runtime_error("dtype must be f16 or bf16: the B200 engine computes on 16-bit tensor cores with f32 "
                             "accumulation (the reference's f32 CPU mode has no device equivalent here)");
  int ndev = 0;
  cuda_check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount (no CUDA device: there is no CPU fallback)");
  if (device < 0 || device >= ndev) throw std::runtime_error("invalid CUDA device ordinal " + std::to_string(device));
  cuda_check(cudaSetDevice(device), "cudaSetDevice");
  cudaDeviceProp prop;
  cuda_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  if (prop.major != 10)
    throw std::runtime_error(std::string("device `") + prop.name + "` is sm_" + std::to_string(prop.major) +
                             std::to_string(prop.minor) + "; this library contains sm_100a code only");
  device_name = prop.name;
  num_sms_ = prop.multiProcessorCount;
  cuda_check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_check(cudaStreamCreateWithFlags(&stream2_, cudaStreamNonBlocking), "cudaStreamCreate");
  cuda_check(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming), "cudaEventCreate");
  cuda_check(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming), "cudaEventCreate");
  cfg_ = parse_config(config_path);
  if (!dsq_path.empty()) {
    // QuantizedSnapshot::load (dsq-runtime/src/lib.rs:237-241): the decoder linears + lm_head come from the snapshot
    DsqReader dsq(dsq_path);
    quantized_ = true;
    load_weights(weights_path, &dsq);
  } else {
    load_weights(weights_path, nullptr);
  }
  cuda_check(cudaDeviceSynchronize(), "weight upload");
}

Engine::~Engine() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  if (stream_ && owns_stream_) cudaStreamDestroy(stream_);
  if (stream2_) cudaStreamDestroy(stream2_);
  if (ev_fork_) cudaEventDestroy(ev_fork_);
  if (ev_join_) cudaEventDestroy(ev_join_);
}

void Engine::set_stream(cudaStream_t s) {
  cuda_check(cudaStreamSynchronize(stream_), "stream switch sync");
  if (owns_stream_ && stream_) cudaStreamDestroy(stream_);
  stream_ = s;
  owns_stream_ = false;
}

DevBuf& Engine::ws(const std::string& name, size_t bytes) {
  DevBuf& b = ws_[name];
  if (b.bytes < bytes) {
    cuda_check(cudaStreamSynchronize(stream_), "workspace grow sync");
    b.alloc(bytes + bytes / 8);
    if (name == "dsq_iota") iota_n_ = 0;
    if (name == "dsqf_attn_cnt" || name == "dsqf_router_cnt") cuda_check(cudaMemset(b.p, 0, b.bytes), "attention counters memset");
  }
  return b;
}

void Engine::record_tap(const std::string& name, const float* dev, size_t n) {
  if (!record_taps_) return;
  cuda_check(cudaStreamSynchronize(stream_), "tap sync");
  std::vector<float>& v = taps_[name];
  v.resize(n);
  d2h(v.data(), dev, n * 4);
}
void Engine::record_tap16(const std::string& name, const void* dev, size_t n) {
  if (!record_taps_) return;
  cuda_check(cudaStreamSynchronize(stream_), "tap sync");
  std::vector<uint16_t> h(n);
  d2h(h.data(), dev, n * 2);
  std::vector<float>& v = taps_[name];
  v.resize(n);
  for (size_t i = 0; i < n; ++i) v[i] = f16_to_32(h[i], dt_);
}
int Engine::tap(const std::string& name, float* out, size_t capacity, size_t* n_written) {
  auto it = taps_.find(name);
  if (it == taps_.end()) throw std::runtime_error("no tap named `" + name + "` (enable recording first)");
  if (n_written) *n_written = it->second.size();
  if (out) {
    if (capacity < it->second.size()) throw std::runtime_error("tap buffer too small");
    memcpy(out, it->second.data(), it->second.size() * 4);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ tables
const float* Engine::sam_pos_for(int g) {
  auto it = sam_pos_.find(g);
  if (it == sam_pos_.end()) {
    const int g0 = cfg_.sam_image / cfg_.sam_patch;
    std::vector<float> t = resize_table_aa(sam_pos_host_.data(), g0, g0, cfg_.sam_dim, g, g);
    DevBuf b;
    upload_f32(b, t);
    it = sam_pos_.emplace(g, std::move(b)).first;
  }
  return it->second.as<float>();
}
const float* Engine::clip_pos_for(int g3) {
  auto it = clip_pos_.find(g3);
  if (it == clip_pos_.end()) {
    const int s0 = cfg_.clip_image / cfg_.clip_patch, C = cfg_.clip_dim;
    std::vector<float> t((size_t)(g3 * g3 + 1) * C);
    memcpy(t.data(), clip_pos_host_.data(), (size_t)C * 4);  // cls row kept (clip.rs:503, 538-541)
    std::vector<float> grid = resize_table_aa(clip_pos_host_.data() + C, s0, s0, C, g3, g3);
    memcpy(t.data() + C, grid.data(), grid.size() * 4);
    DevBuf b;
    upload_f32(b, t);
    it = clip_pos_.emplace(g3, std::move(b)).first;
  }
  return it->second.as<float>();
}
const void* Engine::rel_table_for(int layer, int size, int* zhalf) {
  SamBlockW& b = sam_[layer];
  const int nr = 2 * size - 1;
  *zhalf = (nr + 15) / 16 * 16;
  auto it = b.rel_table.find(size);
  if (it == b.rel_table.end()) {
    std::vector<float> rh = resize_rel_pos(b.rel_h.data(), b.rel_rows, 64, size);
    std::vector<float> rw = resize_rel_pos(b.rel_w.data(), b.rel_rows, 64, size);
    std::vector<float> tab((size_t)2 * (*zhalf) * 64, 0.f);
    memcpy(tab.data(), rh.data(), rh.size() * 4);
    memcpy(tab.data() + (size_t)(*zhalf) * 64, rw.data(), rw.size() * 4);
    DevBuf d;
    upload16_vec(d, tab, dt_);
    it = b.rel_table.emplace(size, std::move(d)).first;
  }
  return it->second.p;
}

// W'[h*2z + n][k] = sum_d tab[n][d] Wq[h*64 + d][k],  b'[h*2z + n] = sum_d tab[n][d] bq[h*64 + d]  (tab rounded to the
// engine dtype like the table the attention kernel would otherwise multiply with)
const SamBlockW::RelFused& Engine::rel_fused_for(int layer, int size, int* zhalf) {
  SamBlockW& b = sam_[layer];
  const int nr = 2 * size - 1;
  *zhalf = (nr + 15) / 16 * 16;
  auto it = b.rel_fused.find(size);
  if (it != b.rel_fused.end()) return it->second;
  const int D = cfg_.sam_dim, Hh = cfg_.sam_heads, zw = 2 * (*zhalf);
  std::vector<float> rh = resize_rel_pos(b.rel_h.data(), b.rel_rows, 64, size);
  std::vector<float> rw = resize_rel_pos(b.rel_w.data(), b.rel_rows, 64, size);
  std::vector<float> tab((size_t)zw * 64, 0.f);
  memcpy(tab.data(), rh.data(), rh.size() * 4);
  memcpy(tab.data() + (size_t)(*zhalf) * 64, rw.data(), rw.size() * 4);
  for (float& v : tab) v = f16_to_32(f32_to_16(v, dt_), dt_);
  std::vector<float> w((size_t)Hh * zw * D, 0.f), bias((size_t)Hh * zw, 0.f);
  parallel_for((size_t)Hh * zw, [&](size_t r0, size_t r1) {
    for (size_t r = r0; r < r1; ++r) {
      const int h = (int)(r / zw), n = (int)(r % zw);
      float* out = &w[r * D];
      float bacc = 0.f;
      for (int d = 0; d < 64; ++d) {
        const float t = tab[(size_t)n * 64 + d];
        if (t == 0.f) continue;
        const float* q = &b.q_w_host[(size_t)(h * 64 + d) * D];
        for (int k = 0; k < D; ++k) out[k] += t * q[k];
        bacc += t * b.q_b_host[h * 64 + d];
      }
      bias[r] = bacc;
    }
  });
  SamBlockW::RelFused f;
  upload16_vec(f.w, w, dt_);
  upload_f32(f.b, bias);
  return b.rel_fused.emplace(size, std::move(f)).first->second;
}

// ------------------------------------------------------------------------------------------------ SAM
void Engine::sam_forward(int Bv, int G, const void* patches16, float* sam_out) {
  const ModelConfig& c = cfg_;
  const int g = G / c.sam_patch, D = c.sam_dim, Hh = c.sam_heads, win = c.sam_window;
  const long long T = (long long)g * g, rows = Bv * T;
  const int nw = (g + win - 1) / win;
  const long long rows_w = (long long)Bv * nw * nw * win * win;
  const long long rmax = std::max(rows, rows_w);
  if (!vision_attention_supported(g)) throw std::runtime_error("unsupported vision token grid " + std::to_string(g) + " (image side " + std::to_string(G) + ")");
  if (g % 4) throw std::runtime_error("spatial dims cannot be evenly downsampled by stride 2 twice");

  float* x = ws("sam_x", rows * D * 4).as<float>();
  void* xn = ws("sam_xn16", rmax * D * 2).p;
  void* qkv = ws("sam_qkv16", rmax * 3 * D * 2).p;
  void* att = ws("sam_att16", rmax * D * 2).p;
  void* h16 = ws("sam_h16", rows * 4 * D * 2).p;
  int* wmap = ws("sam_winmap", rows_w * 4).as<int>();
  window_row_map(wmap, rows_w, win, g, nw, stream_);

  // patch embed (+bias) accumulated onto the broadcast absolute position embedding (sam.rs:238-267)
  bcast_rows(sam_pos_for(g), x, T, Bv, D, stream_);
  {
    LinearCall lc;
    lc.tag = "sam_patch_embed"; lc.w0 = patch_w_.p; lc.x = patches16; lc.x_rows = rows; lc.M = (int)rows; lc.N = D; lc.K = 3 * c.sam_patch * c.sam_patch;
    lc.bias = patch_b_.as<float>(); lc.out = x; lc.ldo = D; lc.out_mode = lin::OUT_F32_ADD;
    linear(lc, dt_, num_sms_, stream_);
  }
  record_tap("sam.pos_added", x, rows * D);

  for (int i = 0; i < c.sam_depth; ++i) {
    SamBlockW& b = sam_[i];
    const bool glob = c.sam_is_global(i);
    const long long r = glob ? rows : rows_w;
    const int S = glob ? (int)T : win * win;
    const int size = glob ? g : win;
    layernorm(x, b.ln1_w.as<float>(), b.ln1_b.as<float>(), xn, nullptr, r, D, 1e-6f, glob ? 0 : win, g, nw, dt_, stream_);
    {
      LinearCall lc;
      lc.tag = "sam_qkv"; lc.w0 = b.qkv_w.p; lc.x = xn; lc.x_rows = r; lc.M = (int)r; lc.N = 3 * D; lc.K = D;
      lc.bias = b.qkv_b.as<float>(); lc.out = qkv; lc.ldo = 3 * D; lc.out_mode = lin::OUT_T;
      linear(lc, dt_, num_sms_, stream_);
    }
    int zhalf = 0;
    // rel-pos logits Z[row, head, (kh | kw)] straight from the block input (one more GEMM over xn instead of a
    // K = 64 product per head over the rounded q)
    const SamBlockW::RelFused& rf = rel_fused_for(i, size, &zhalf);
    float* Z = ws("sam_z32", r * Hh * 2 * zhalf * 4).as<float>();
    {
      LinearCall lc;
      lc.tag = "sam_relpos_products"; lc.w0 = rf.w.p; lc.x = xn; lc.x_rows = r; lc.M = (int)r; lc.N = Hh * 2 * zhalf; lc.K = D;
      lc.bias = rf.b.as<float>(); lc.out = Z; lc.ldo = (long long)Hh * 2 * zhalf; lc.out_mode = lin::OUT_F32;
      linear(lc, dt_, num_sms_, stream_);
    }
    {
      VAttnCall ac;
      ac.qkv = qkv; ac.rows = r; ac.B = (int)(r / S); ac.S = S; ac.H = Hh; ac.grid = size;
      ac.tag = glob ? "sam_global_attention" : "sam_window_attention"; ac.Z = Z; ac.zw = 2 * zhalf; ac.zhalf = zhalf; ac.out = att; ac.scale = 0.125f;
      vision_attention(ac, dt_, stream_);
    }
    {
      LinearCall lc;
      lc.tag = "sam_proj"; lc.w0 = b.proj_w.p; lc.x = att; lc.x_rows = r; lc.M = (int)r; lc.N = D; lc.K = D;
      lc.bias = b.proj_b.as<float>(); lc.out = x; lc.ldo = D; lc.out_mode = lin::OUT_F32_ADD;
      lc.row_map = glob ? nullptr : wmap;
      linear(lc, dt_, num_sms_, stream_);
    }
    layernorm(x, b.ln2_w.as<float>(), b.ln2_b.as<float>(), xn, nullptr, rows, D, 1e-6f, 0, g, nw, dt_, stream_);
    {
      LinearCall lc;
      lc.tag = "sam_fc1_gelu"; lc.w0 = b.fc1_w.p; lc.x = xn; lc.x_rows = rows; lc.M = (int)rows; lc.N = 4 * D; lc.K = D;
      lc.bias = b.fc1_b.as<float>(); lc.out = h16; lc.ldo = 4 * D; lc.out_mode = lin::OUT_T; lc.act = lin::ACT_GELU_ERF;
      linear(lc, dt_, num_sms_, stream_);
    }
    {
      LinearCall lc;
      lc.tag = "sam_fc2"; lc.w0 = b.fc2_w.p; lc.x = h16; lc.x_rows = rows; lc.M = (int)rows; lc.N = D; lc.K = 4 * D;
      lc.bias = b.fc2_b.as<float>(); lc.out = x; lc.ldo = D; lc.out_mode = lin::OUT_F32_ADD;
      linear(lc, dt_, num_sms_, stream_);
    }
    record_tap("sam.block." + std::to_string(i), x, rows * D);
  }

  // neck + 16x token compressor (sam.rs:475-576), NHWC throughout, LN2d == row LayerNorm over channels
  const int NC = c.sam_neck;
  float* n32 = ws("sam_neck32", rows * NC * 4).as<float>();
  void* n16 = ws("sam_neck16", rows * NC * 2).p;
  void* col = ws("sam_col16", rows * 9 * NC * 2).p;
  cast16(x, xn, rows * D, dt_, stream_);
  {
    LinearCall lc;
    lc.tag = "sam_neck_conv1"; lc.w0 = neck0_w_.p; lc.x = xn; lc.x_rows = rows; lc.M = (int)rows; lc.N = NC; lc.K = D;
    lc.out = n32; lc.ldo = NC; lc.out_mode = lin::OUT_F32;
    linear(lc, dt_, num_sms_, stream_);
  }
  record_tap("sam.neck_conv1", n32, rows * NC);
  layernorm(n32, neck1_w_.as<float>(), neck1_b_.as<float>(), n16, nullptr, rows, NC, 1e-6f, 0, g, nw, dt_, stream_);
  im2col3x3(n16, col, Bv, g, g, NC, 1, dt_, stream_);
  {
    LinearCall lc;
    lc.tag = "sam_neck_conv2"; lc.w0 = neck2_w_.p; lc.x = col; lc.x_rows = rows; lc.M = (int)rows; lc.N = NC; lc.K = 9 * NC;
    lc.out = n32; lc.ldo = NC; lc.out_mode = lin::OUT_F32;
    linear(lc, dt_, num_sms_, stream_);
  }
  record_tap("sam.neck_conv2", n32, rows * NC);
  layernorm(n32, neck3_w_.as<float>(), neck3_b_.as<float>(), n16, nullptr, rows, NC, 1e-6f, 0, g, nw, dt_, stream_);
  const int g2 = g / 2, g3 = g / 4;
  const long long rows2 = (long long)Bv * g2 * g2, rows3 = (long long)Bv * g3 * g3;
  im2col3x3(n16, col, Bv, g, g, NC, 2, dt_, stream_);
  void* d2 = ws("sam_net2_16", rows2 * c.sam_out0 * 2).p;
  {
    LinearCall lc;
    lc.tag = "sam_net2"; lc.w0 = net2_w_.p; lc.x = col; lc.x_rows = rows2; lc.M = (int)rows2; lc.N = c.sam_out0; lc.K = 9 * NC;
    lc.out = d2; lc.ldo = c.sam_out0; lc.out_mode = lin::OUT_T;
    linear(lc, dt_, num_sms_, stream_);
  }
  void* col3 = ws("sam_col3_16", rows3 * 9 * c.sam_out0 * 2).p;
  im2col3x3(d2, col3, Bv, g2, g2, c.sam_out0, 2, dt_, stream_);
  {
    LinearCall lc;
    lc.tag = "sam_net3"; lc.w0 = net3_w_.p; lc.x = col3; lc.x_rows = rows3; lc.M = (int)rows3; lc.N = c.sam_out1; lc.K = 9 * c.sam_out0;
    lc.out = sam_out; lc.ldo = c.sam_out1; lc.out_mode = lin::OUT_F32;
    linear(lc, dt_, num_sms_, stream_);
  }
  record_tap("sam.net3", sam_out, rows3 * c.sam_out1);
}

// ------------------------------------------------------------------------------------------------ CLIP
void Engine::clip_forward(int Bv, int g3, const float* sam_out, float* clip_x) {
  const ModelConfig& c = cfg_;
  const int C = c.clip_dim, n = g3 * g3, S = n + 1, Hh = c.clip_heads;
  const long long rows = (long long)Bv * S;
  float* emb = ws("clip_emb32", rows * C * 4).as<float>();
  void* xn = ws("clip_xn16", rows * C * 2).p;
  void* qkv = ws("clip_qkv16", rows * 3 * C * 2).p;
  void* att = ws("clip_att16", rows * C * 2).p;
  void* h16 = ws("clip_h16", rows * 4 * C * 2).p;
  clip_embed(sam_out, clip_cls_.as<float>(), clip_pos_for(g3), emb, Bv, n, C, stream_);
  record_tap("clip.embeddings", emb, rows * C);
  layernorm(emb, clip_preln_w_.as<float>(), clip_preln_b_.as<float>(), nullptr, clip_x, rows, C, 1e-5f, 0, 0, 0, dt_, stream_);
  record_tap("clip.pre_layernorm", clip_x, rows * C);
  for (int i = 0; i < c.clip_layers; ++i) {
    ClipBlockW& b = clip_[i];
    layernorm(clip_x, b.ln1_w.as<float>(), b.ln1_b.as<float>(), xn, nullptr, rows, C, 1e-5f, 0, 0, 0, dt_, stream_);
    {
      LinearCall lc;
      lc.tag = "clip_qkv"; lc.w0 = b.qkv_w.p; lc.x = xn; lc.x_rows = rows; lc.M = (int)rows; lc.N = 3 * C; lc.K = C;
      lc.bias = b.qkv_b.as<float>(); lc.out = qkv; lc.ldo = 3 * C; lc.out_mode = lin::OUT_T;
      linear(lc, dt_, num_sms_, stream_);
    }
    {
      VAttnCall ac;
      ac.qkv = qkv; ac.rows = rows; ac.B = Bv; ac.S = S; ac.H = Hh; ac.grid = 0; ac.out = att; ac.scale = 0.125f; ac.tag = "clip_attention";
      vision_attention(ac, dt_, stream_);
    }
    {
      LinearCall lc;
      lc.tag = "clip_out_proj"; lc.w0 = b.out_w.p; lc.x = att; lc.x_rows = rows; lc.M = (int)rows; lc.N = C; lc.K = C;
      lc.bias = b.out_b.as<float>(); lc.out = clip_x; lc.ldo = C; lc.out_mode = lin::OUT_F32_ADD;
      linear(lc, dt_, num_sms_, stream_);
    }
    layernorm(clip_x, b.ln2_w.as<float>(), b.ln2_b.as<float>(), xn, nullptr, rows, C, 1e-5f, 0, 0, 0, dt_, stream_);
    {
      LinearCall lc;
      lc.tag = "clip_fc1_quickgelu"; lc.w0 = b.fc1_w.p; lc.x = xn; lc.x_rows = rows; lc.M = (int)rows; lc.N = 4 * C; lc.K = C;
      lc.bias = b.fc1_b.as<float>(); lc.out = h16; lc.ldo = 4 * C; lc.out_mode = lin::OUT_T; lc.act = lin::ACT_QUICK_GELU;
      linear(lc, dt_, num_sms_, stream_);
    }
    {
      LinearCall lc;
      lc.tag = "clip_fc2"; lc.w0 = b.fc2_w.p; lc.x = h16; lc.x_rows = rows; lc.M = (int)rows; lc.N = C; lc.K = 4 * C;
      lc.bias = b.fc2_b.as<float>(); lc.out = clip_x; lc.ldo = C; lc.out_mode = lin::OUT_F32_ADD;
      linear(lc, dt_, num_sms_, stream_);
    }
    record_tap("clip.layer." + std::to_string(i), clip_x, rows * C);
  }
}

// One batch of same-size views: patch gather -> SAM -> CLIP -> concat -> projector.
void Engine::vision_views(int Bv, int G, const void* img_dev, bool is_f32, float* proj_out, const char* tag) {
  const ModelConfig& c = cfg_;
  if (G % (c.sam_patch * 4)) throw std::runtime_error("image side must be a multiple of 64, got " + std::to_string(G));
  const int g = G / c.sam_patch, g3 = g / 4, n = g3 * g3;
  // chunk the batch so that activations stay bounded (~160k SAM tokens per pass)
  const int max_views = std::max(1, (int)(163840 / ((long long)g * g)));
  for (int b0 = 0; b0 < Bv; b0 += max_views) {
    const int nb = std::min(max_views, Bv - b0);
    const long long rows = (long long)nb * g * g;
    if (view_ready) view_ready(tag[0] == 'l', b0 + nb - 1);
    void* patches = ws("patches16", rows * 768 * 2).p;
    if (is_f32) patchify_f32((const float*)img_dev + (size_t)b0 * 3 * G * G, patches, nb, G, dt_, stream_);
    else patchify_u8((const uint8_t*)img_dev + (size_t)b0 * 3 * G * G, patches, nb, G, dt_, stream_);
    float* sam_out = ws("sam_out32", (size_t)nb * n * c.sam_out1 * 4).as<float>();
    sam_forward(nb, G, patches, sam_out);
    float* clip_x = ws("clip_x32", (size_t)nb * (n + 1) * c.clip_dim * 4).as<float>();
    clip_forward(nb, g3, sam_out, clip_x);
    void* pre16 = ws("pre16", (size_t)nb * n * c.proj_in * 2).p;
    float* pre32 = record_taps_ ? ws("pre32", (size_t)nb * n * c.proj_in * 4).as<float>() : nullptr;
    concat_clip_sam(clip_x, sam_out, pre16, pre32, nb, n, c.clip_dim, dt_, stream_);
    if (pre32) record_tap(std::string(tag) + "_pre", pre32, (size_t)nb * n * c.proj_in);
    LinearCall lc;
    lc.tag = "projector"; lc.w0 = proj_w_.p; lc.x = pre16; lc.x_rows = (long long)nb * n; lc.M = nb * n; lc.N = c.n_embed; lc.K = c.proj_in;
    lc.bias = proj_b_.as<float>(); lc.out = proj_out + (size_t)b0 * n * c.n_embed; lc.ldo = c.n_embed;
    lc.out_mode = lin::OUT_F32;
    linear(lc, dt_, num_sms_, stream_);
    if (record_taps_) record_tap(std::string(tag) + "_post", proj_out + (size_t)b0 * n * c.n_embed, (size_t)nb * n * c.n_embed);
  }
}

const float* Engine::vision_encode(int n_pages, const void* globals_dev, bool globals_f32, int G, const void* tiles_dev,
                                   bool tiles_f32, int P, const std::vector<PageViews>& pages, std::vector<int>* n_rows) {
  const ModelConfig& c = cfg_;
  const int qg = G / 64, ql = P > 0 ? P / 64 : 0, Hd = c.n_embed;
  int total_tiles = 0;
  for (auto& pv : pages) total_tiles += pv.n_tiles;
  const size_t grows = (size_t)n_pages * qg * qg, lrows = (size_t)total_tiles * ql * ql;
  kernel_timing_phase("vision/");
  float* proj = ws("proj_all32", (grows + lrows) * Hd * 4).as<float>();
  vision_views(n_pages, G, globals_dev, globals_f32, proj, "global");
  if (total_tiles > 0) vision_views(total_tiles, P, tiles_dev, tiles_f32, proj + grows * Hd, "local");
  // token layout: [local grid + newline column ; global grid + newline column ; view separator]
  std::vector<int> map;
  n_rows->clear();
  int tile_base = 0;
  for (int p = 0; p < n_pages; ++p) {
    const PageViews& pv = pages[p];
    const size_t before = map.size();
    if (pv.n_tiles > 0) {
      if (pv.n_tiles != pv.crop_w * pv.crop_h) throw std::runtime_error("patch count does not match crop grid");
      for (int R = 0; R < pv.crop_h * ql; ++R) {
        for (int Cc = 0; Cc < pv.crop_w * ql; ++Cc) {
          const int crop = (R / ql) * pv.crop_w + (Cc / ql);
          const int tok = (R % ql) * ql + (Cc % ql);
          map.push_back((int)grows + (tile_base + crop) * ql * ql + tok);
        }
        map.push_back(-1);
      }
    }
    for (int r = 0; r < qg; ++r) {
      for (int cc = 0; cc < qg; ++cc) map.push_back(p * qg * qg + r * qg + cc);
      map.push_back(-1);
    }
    map.push_back(-2);
    tile_base += pv.n_tiles;
    n_rows->push_back((int)(map.size() - before));
  }
  int* dmap = ws("tokmap", map.size() * 4).as<int>();
  cuda_check(cudaMemcpyAsync(dmap, map.data(), map.size() * 4, cudaMemcpyHostToDevice, stream_), "token map upload");
  cuda_check(cudaStreamSynchronize(stream_), "token map sync");  // `map` is a stack vector
  float* out = ws("image_rows32", map.size() * Hd * 4).as<float>();
  scatter_tokens(proj, newline_.as<float>(), separator_.as<float>(), dmap, out, (long long)map.size(), Hd, stream_);
  return out;
}

}  // namespace dsocr
