// extern "C" boundary (include/dsocr.h).  Plain pointers and sizes in, status codes out.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <thread>

#include "dsocr.h"
#include "dsocr_test.h"
#include "dsq_dequant.h"
#include "engine.h"
#include "hostmath.h"
#include "sampler.h"

using namespace dsocr;

struct dsocr_engine {
  std::unique_ptr<Engine> impl;
  // pages staged on the device by dsocr_stage_pages (views already resized / tiled, RGB8)
  struct Staged {
    int n_pages = 0;
    dsocr_vision_settings vs{};
    std::vector<int> ntiles, cw, ch;
    DevBuf globals, tiles;
  } staged;
  // device-side integer preprocessing: cached coefficient tables per (in, out) size + scratch
  struct DevCoef { DevBuf start, len, coef; int ksize = 0; };
  std::map<std::pair<int, int>, DevCoef> coefs;
  DevBuf pages_raw, horiz;
  // decode_pages / decode_requests stage their pages on a side stream (H2D copy + resample per page, an event every
  // kPagesPerEvent pages) so that the copies of later pages run under the vision tower of the earlier ones
  static constexpr int kPagesPerEvent = 8;
  static constexpr int kStageAhead = 96;  // pages issued beyond the ones a vision chunk is about to read
  std::function<void(int)> stage_fn;      // valid only while the call that owns the host pages is running
  int stage_next = 0;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> page_events;
  std::vector<int> tile_page;  // staged tile index -> page index
  bool staged_async = false;
  ~dsocr_engine() {
    for (cudaEvent_t ev : page_events) cudaEventDestroy(ev);
    if (copy_stream) cudaStreamDestroy(copy_stream);
  }
  bool host_preprocess = false;
  int decode_batch = 512;  // requests decoded in lock-step per group
};

namespace {
void gpu_prepare_page(dsocr_engine* e, const uint8_t* src, int w, int h, dsocr_vision_settings vs, uint8_t* global_dst,
                      uint8_t* tiles_dst, int gw, int gh, int n_tiles, cudaStream_t st = nullptr);
int status_of(const std::exception& e) {
  const std::string m = e.what();
  if (m.find("prompt/image embedding mismatch") != std::string::npos) return DSOCR_ERR_MISMATCH;
  if (m.find("CUDA") != std::string::npos || m.find("cuda") != std::string::npos) return DSOCR_ERR_CUDA;
  if (m.find("unsupported") != std::string::npos || m.find("not supported") != std::string::npos) return DSOCR_ERR_UNSUPPORTED;
  if (m.find("failed to read") != std::string::npos || m.find("cannot open") != std::string::npos) return DSOCR_ERR_IO;
  return DSOCR_ERR_INTERNAL;
}
template <typename F>
int api(const char* context, F&& f) {
  try {
    f();
    return DSOCR_OK;
  } catch (const std::exception& e) {
    set_last_error(context && *context ? std::string(context) + ": " + e.what() : std::string(e.what()));
    return status_of(e);
  } catch (...) {
    set_last_error("unknown error");
    return DSOCR_ERR_INTERNAL;
  }
}
void bind(dsocr_engine* e) {
  if (!e || !e->impl) throw std::runtime_error("null engine handle");
  cuda_check(cudaSetDevice(e->impl->device()), "cudaSetDevice");
}
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" int dsocr_engine_create(const char* config_json_path, const char* safetensors_path, const char* dsq_path,
                                   int device_ordinal, int dtype, dsocr_engine** out) {
  return api("load_model", [&] {
    if (!config_json_path || !safetensors_path || !out) throw std::runtime_error("null argument");
    if (dtype != DSOCR_F16 && dtype != DSOCR_BF16 && dtype != DSOCR_F32) throw std::runtime_error("invalid dtype code");
    auto h = std::make_unique<dsocr_engine>();
    h->impl = std::make_unique<Engine>(config_json_path, safetensors_path, dsq_path ? dsq_path : "", device_ordinal,
                                       static_cast<DType>(dtype));
    *out = h.release();
  });
}

extern "C" void dsocr_engine_destroy(dsocr_engine* e) { delete e; }

extern "C" int dsocr_engine_info_get(const dsocr_engine* e, dsocr_engine_info* info) {
  return api("", [&] {
    if (!e || !info) throw std::runtime_error("null argument");
    const Engine& en = *e->impl;
    memset(info, 0, sizeof(*info));
    info->device_ordinal = en.device();
    info->dtype = (int)en.dtype();
    info->sm_count = en.sm_count();
    info->hidden_size = en.cfg().hidden; info->num_layers = en.cfg().layers; info->vocab_size = en.cfg().vocab;
    info->n_routed_experts = en.cfg().n_experts;
    info->quantized = en.quantized() ? 1 : 0;
    strncpy(info->device_name, en.device_name.c_str(), sizeof(info->device_name) - 1);
  });
}

extern "C" int dsocr_engine_set_option(dsocr_engine* e, const char* name, int value) {
  return api("", [&] {
    bind(e);
    const std::string n = name ? name : "";
    if (n == "record_taps") e->impl->set_record_taps(value != 0);
    else if (n == "kv_cache_f16") e->impl->set_kv_f16(value != 0);
    else if (n == "moe_stats") e->impl->set_moe_stats(value != 0);
    else if (n == "host_preprocess") e->host_preprocess = value != 0;
    else if (n == "decode_batch") { if (value < 1) throw std::runtime_error("decode_batch must be >= 1"); e->decode_batch = value; }
    else throw std::runtime_error("unknown option `" + n + "`");
  });
}

extern "C" int dsocr_dsq_inspect(const char* path, dsocr_dsq_header* header, dsocr_dsq_record* records, size_t capacity) {
  return api("snapshot open failed", [&] {
    if (!path) throw std::runtime_error("null argument");
    DsqReader rd(path);
    if (header) {
      memset(header, 0, sizeof(*header));
      header->version = 1;
      header->default_qdtype = (uint32_t)rd.default_dtype();
      header->block_size = rd.block_size();
      header->tensor_count = (uint32_t)rd.records().size();
      strncpy(header->candle_version, rd.candle_version.c_str(), sizeof(header->candle_version) - 1);
      strncpy(header->model_id, rd.model_id.c_str(), sizeof(header->model_id) - 1);
      strncpy(header->backend, rd.backend.c_str(), sizeof(header->backend) - 1);
    }
    for (size_t i = 0; records && i < rd.records().size() && i < capacity; ++i) {
      const DsqRecord& r = rd.records()[i];
      dsocr_dsq_record& o = records[i];
      memset(&o, 0, sizeof(o));
      strncpy(o.name, r.name.c_str(), sizeof(o.name) - 1);
      o.out_dim = r.out_dim; o.in_dim = r.in_dim; o.q_dtype = (uint32_t)r.q_dtype;
      o.q_offset = r.q_offset; o.q_len = r.q_len;
      o.bias_offset = r.has_bias ? r.bias_offset : 0; o.bias_len = r.has_bias ? r.bias_len : 0;
      o.bias_dtype = r.has_bias ? r.bias_dtype : 0;
      o.first_q_byte = rd.bytes(r)[0];
    }
  });
}

extern "C" int dsocr_test_dsq_dequant64(uint32_t q_dtype, const uint8_t* blocks, int rows, int K, float* out) {
  return api("", [&] {
    if (!blocks || !out || rows <= 0) throw std::runtime_error("null argument");
    const DsqDType dt = static_cast<DsqDType>(q_dtype);
    const int be = dsq_block_elems(dt);
    if (!be || K % be || K % 64) throw std::runtime_error("K must be a multiple of the block size and of 64");
    // the plane split of dsq_upload_rows, kept on the host
    std::vector<uint8_t> a, b, c, d;
    const size_t nb = (size_t)rows * (K / be);
    if (dt == DsqDType::Q8_0) {
      a.resize((size_t)rows * K); b.resize(nb * 2);
      for (size_t i = 0; i < nb; ++i) { memcpy(&b[2 * i], blocks + i * 34, 2); memcpy(&a[32 * i], blocks + i * 34 + 2, 32); }
    } else if (dt == DsqDType::Q4K) {
      a.assign(blocks, blocks + nb * 144);
    } else {
      a.resize(nb * 128); b.resize(nb * 64); c.resize(nb * 16); d.resize(nb * 2);
      for (size_t i = 0; i < nb; ++i) {
        const uint8_t* blk = blocks + i * 210;
        memcpy(&a[128 * i], blk, 128); memcpy(&b[64 * i], blk + 128, 64); memcpy(&c[16 * i], blk + 192, 16); memcpy(&d[2 * i], blk + 208, 2);
      }
    }
    const DsqPlanes pl{a.data(), b.data(), c.data(), d.data()};
    for (int r = 0; r < rows; ++r)
      for (int kb = 0; kb < K / 64; ++kb) dsq_dequant64((int)q_dtype, pl, r, K, kb, out + (size_t)r * K + (size_t)kb * 64);
  });
}

extern "C" int dsocr_test_chacha_words(const uint8_t* key32, int rounds, int n, uint32_t* out) {
  return api("", [&] {
    if (!key32 || !out || n < 0) throw std::runtime_error("null argument");
    StdRng rng(key32, rounds);
    for (int i = 0; i < n; ++i) out[i] = rng.next_u32();
  });
}

extern "C" int dsocr_test_stdrng_u64(uint64_t seed, int n, uint64_t* out) {
  return api("", [&] {
    if (!out || n < 0) throw std::runtime_error("null argument");
    StdRng rng(seed);
    for (int i = 0; i < n; ++i) out[i] = rng.next_u64();
  });
}

extern "C" int dsocr_test_select_tokens(const float* logits, size_t vocab, int n_steps, const dsocr_decode_params* params,
                                        const int64_t* context, size_t n_context, int64_t* out) {
  return api("", [&] {
    if (!logits || !params || !out || (!context && n_context)) throw std::runtime_error("null argument");
    const SamplingParams sp = sampling_params_of(*params);
    StdRng rng = params->has_seed ? StdRng(params->seed) : StdRng::from_entropy();
    std::vector<int64_t> ctx(context, context + n_context);
    for (int s = 0; s < n_steps; ++s) {
      out[s] = select_token_id(logits + (size_t)s * vocab, vocab, sp, ctx.data(), ctx.size(), rng);
      ctx.push_back(out[s]);
    }
  });
}

struct dsocr_dsq_writer { std::unique_ptr<DsqWriter> impl; };

static DsqDType writer_dtype(uint32_t v) {
  switch (v) {
    case 0: case 1: case 8: case 12: case 14: case 16: return static_cast<DsqDType>(v);
    default: throw std::runtime_error("unsupported tensor dtype code " + std::to_string(v));
  }
}

extern "C" int dsocr_dsq_writer_create(const char* path, const char* candle_version, const char* model_id, const char* backend,
                                       uint32_t default_qdtype, dsocr_dsq_writer** out) {
  return api("snapshot writer", [&] {
    if (!path || !out) throw std::runtime_error("null argument");
    auto h = std::make_unique<dsocr_dsq_writer>();
    h->impl = std::make_unique<DsqWriter>(path, candle_version ? candle_version : "", model_id ? model_id : "",
                                          backend ? backend : "", writer_dtype(default_qdtype));
    *out = h.release();
  });
}

extern "C" int dsocr_dsq_writer_add_tensor(dsocr_dsq_writer* w, const char* name, uint32_t out_dim, uint32_t in_dim, uint32_t q_dtype,
                                           const float* weights, const float* bias) {
  return api("snapshot writer", [&] {
    if (!w || !w->impl || !name || !weights) throw std::runtime_error("null argument");
    w->impl->add_tensor_f32(name, out_dim, in_dim, writer_dtype(q_dtype), weights, bias);
  });
}

extern "C" int dsocr_dsq_writer_add_quantized_bytes(dsocr_dsq_writer* w, const char* name, uint32_t out_dim, uint32_t in_dim,
                                                    uint32_t q_dtype, const uint8_t* qbytes, size_t q_len, const float* bias) {
  return api("snapshot writer", [&] {
    if (!w || !w->impl || !name || !qbytes) throw std::runtime_error("null argument");
    w->impl->add_quantized_bytes(name, out_dim, in_dim, writer_dtype(q_dtype), qbytes, q_len, bias);
  });
}

extern "C" int dsocr_dsq_writer_finalize(dsocr_dsq_writer* w) {
  const int st = api("snapshot writer", [&] {
    if (!w || !w->impl) throw std::runtime_error("null argument");
    w->impl->finalize();
  });
  delete w;
  return st;
}

extern "C" void dsocr_dsq_writer_destroy(dsocr_dsq_writer* w) { delete w; }

extern "C" int dsocr_image_token_count(uint32_t base_size, uint32_t image_size, int crop_mode, int crop_w, int crop_h) {
  return image_token_count((int)base_size, (int)image_size, crop_mode, crop_w, crop_h);
}

extern "C" int dsocr_preprocess(const uint8_t* rgb, int width, int height, dsocr_vision_settings vs, uint8_t* global_out,
                                uint8_t* tiles_out, int* n_tiles, int* crop_w, int* crop_h) {
  return api("vision input failed", [&] {
    if (!rgb || width <= 0 || height <= 0) throw std::runtime_error("empty image");
    const int gsz = vs.crop_mode ? (int)vs.base_size : (int)vs.image_size;
    if (global_out) build_global_view_u8(rgb, width, height, gsz, global_out);
    int gw = 1, gh = 1, n = 0;
    if (vs.crop_mode) n = dynamic_preprocess_u8(rgb, width, height, (int)vs.image_size, tiles_out, &gw, &gh);
    if (n_tiles) *n_tiles = n;
    if (crop_w) *crop_w = gw;
    if (crop_h) *crop_h = gh;
  });
}

extern "C" int dsocr_preprocess_gpu(dsocr_engine* e, const uint8_t* rgb, int width, int height, dsocr_vision_settings vs,
                                    uint8_t* global_out, uint8_t* tiles_out, int* n_tiles, int* crop_w, int* crop_h) {
  return api("vision input failed", [&] {
    bind(e);
    Engine& en = *e->impl;
    if (!rgb || width <= 0 || height <= 0) throw std::runtime_error("empty image");
    const int G = vs.crop_mode ? (int)vs.base_size : (int)vs.image_size, P = (int)vs.image_size;
    int gw = 1, gh = 1, n = 0;
    if (vs.crop_mode) n = dynamic_preprocess_u8(rgb, width, height, P, nullptr, &gw, &gh);
    DevBuf src((size_t)width * height * 3), dg((size_t)G * G * 3), dt(std::max<size_t>(16, (size_t)n * P * P * 3));
    h2d(src.p, rgb, (size_t)width * height * 3);
    gpu_prepare_page(e, src.as<uint8_t>(), width, height, vs, dg.as<uint8_t>(), dt.as<uint8_t>(), gw, gh, n);
    cuda_check(cudaStreamSynchronize(en.stream()), "preprocess sync");
    if (global_out) d2h(global_out, dg.p, (size_t)G * G * 3);
    if (tiles_out && n > 0) d2h(tiles_out, dt.p, (size_t)n * P * P * 3);
    if (n_tiles) *n_tiles = n;
    if (crop_w) *crop_w = gw;
    if (crop_h) *crop_h = gh;
  });
}

extern "C" int dsocr_vision_encode(dsocr_engine* e, const float* global_chw, int global_size, const float* patches_nchw,
                                   int n_patches, int patch_size, int crop_w, int crop_h, float* out_rows, int* n_rows) {
  return api("image embedding failed", [&] {
    bind(e);
    Engine& en = *e->impl;
    if (!global_chw || !out_rows || !n_rows) throw std::runtime_error("null argument");
    const size_t gbytes = (size_t)3 * global_size * global_size * 4;
    const size_t pbytes = (size_t)n_patches * 3 * patch_size * patch_size * 4;
    DevBuf dg(gbytes), dp(std::max<size_t>(pbytes, 16));
    h2d(dg.p, global_chw, gbytes);
    if (n_patches > 0) h2d(dp.p, patches_nchw, pbytes);
    std::vector<Engine::PageViews> pages(1);
    pages[0].n_tiles = n_patches; pages[0].crop_w = crop_w; pages[0].crop_h = crop_h;
    std::vector<int> counts;
    const float* rows = en.vision_encode(1, dg.p, true, global_size, dp.p, true, patch_size, pages, &counts);
    if (counts[0] > *n_rows) throw std::runtime_error("output buffer too small for " + std::to_string(counts[0]) + " rows");
    cuda_check(cudaStreamSynchronize(en.stream()), "vision sync");
    d2h(out_rows, rows, (size_t)counts[0] * en.cfg().n_embed * 4);
    *n_rows = counts[0];
  });
}

namespace {
// uploads u8 views and runs the vision tower; returns device rows + counts
const float* encode_u8(Engine& en, int n_pages, const uint8_t* const* globals_u8, int G, const uint8_t* const* tiles_u8,
                       const int* n_tiles, int P, const int* crop_w, const int* crop_h, std::vector<int>* counts,
                       DevBuf& dg, DevBuf& dt) {
  const size_t gbytes = (size_t)G * G * 3;
  const size_t tbytes = (size_t)P * P * 3;
  int total_tiles = 0;
  std::vector<Engine::PageViews> pages(n_pages);
  for (int p = 0; p < n_pages; ++p) {
    pages[p].n_tiles = n_tiles ? n_tiles[p] : 0;
    pages[p].crop_w = crop_w ? crop_w[p] : 1;
    pages[p].crop_h = crop_h ? crop_h[p] : 1;
    total_tiles += pages[p].n_tiles;
  }
  dg.ensure(gbytes * n_pages);
  dt.ensure(std::max<size_t>(16, tbytes * total_tiles));
  size_t toff = 0;
  for (int p = 0; p < n_pages; ++p) {
    cuda_check(cudaMemcpyAsync((uint8_t*)dg.p + gbytes * p, globals_u8[p], gbytes, cudaMemcpyHostToDevice, en.stream()), "global view H2D");
    if (pages[p].n_tiles > 0) {
      cuda_check(cudaMemcpyAsync((uint8_t*)dt.p + toff, tiles_u8[p], tbytes * pages[p].n_tiles, cudaMemcpyHostToDevice, en.stream()), "tiles H2D");
      toff += tbytes * pages[p].n_tiles;
    }
  }
  return en.vision_encode(n_pages, dg.p, false, G, dt.p, false, P, pages, counts);
}
}  // namespace

extern "C" int dsocr_vision_encode_u8_batch(dsocr_engine* e, int n_pages, const uint8_t* const* globals_u8,
                                            int global_size, const uint8_t* const* tiles_u8, const int* n_tiles,
                                            int patch_size, const int* crop_w, const int* crop_h,
                                            float* const* out_rows, int* n_rows) {
  return api("image embedding failed", [&] {
    bind(e);
    Engine& en = *e->impl;
    DevBuf dg, dt;
    std::vector<int> counts;
    const double t0 = now_ms();
    const float* rows = encode_u8(en, n_pages, globals_u8, global_size, tiles_u8, n_tiles, patch_size, crop_w, crop_h, &counts, dg, dt);
    cuda_check(cudaStreamSynchronize(en.stream()), "vision sync");
    en.timings.vision = now_ms() - t0;
    size_t off = 0;
    for (int p = 0; p < n_pages; ++p) {
      if (out_rows && out_rows[p]) d2h(out_rows[p], rows + off * en.cfg().n_embed, (size_t)counts[p] * en.cfg().n_embed * 4);
      if (n_rows) n_rows[p] = counts[p];
      off += counts[p];
    }
  });
}

extern "C" int dsocr_vision_tap(dsocr_engine* e, const char* name, float* out, size_t capacity, size_t* n_written) {
  return api("", [&] { bind(e); e->impl->tap(name ? name : "", out, capacity, n_written); });
}

extern "C" int dsocr_generate_batch(dsocr_engine* e, int n_pages, const int64_t* const* input_ids,
                                    const uint8_t* const* images_seq_mask, const int* n_tokens,
                                    const float* const* image_rows, const int* n_image_rows,
                                    const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                                    int64_t* const* out_tokens, int* n_out) {
  return api("", [&] {
    bind(e);
    if (!params) throw std::runtime_error("null decode params");
    Engine::GenRequest rq;
    rq.n_pages = n_pages; rq.input_ids = input_ids; rq.mask = images_seq_mask; rq.n_tokens = n_tokens;
    rq.image_rows_host = image_rows; rq.n_image_rows = n_image_rows; rq.params = *params; rq.cb = cb; rq.user = user;
    e->impl->generate(rq, out_tokens, n_out);
  });
}

extern "C" int dsocr_generate_forced(dsocr_engine* e, int n_pages, const int64_t* const* input_ids,
                                     const uint8_t* const* images_seq_mask, const int* n_tokens,
                                     const float* const* image_rows, const int* n_image_rows,
                                     const dsocr_decode_params* params, const int64_t* const* forced_tokens,
                                     int n_steps, int64_t* const* selected_out, float* const* logits_out) {
  return api("", [&] {
    bind(e);
    if (!params || !forced_tokens) throw std::runtime_error("null argument");
    Engine::GenRequest rq;
    rq.n_pages = n_pages; rq.input_ids = input_ids; rq.mask = images_seq_mask; rq.n_tokens = n_tokens;
    rq.image_rows_host = image_rows; rq.n_image_rows = n_image_rows; rq.params = *params;
    rq.forced = forced_tokens; rq.n_forced_steps = n_steps; rq.logits_out = logits_out;
    std::vector<int> n_out(n_pages);
    e->impl->generate(rq, selected_out, n_out.data());
  });
}

namespace {
const dsocr_engine::DevCoef& coef_for(dsocr_engine* e, int in_size, int out_size) {
  auto key = std::make_pair(in_size, out_size);
  auto it = e->coefs.find(key);
  if (it == e->coefs.end()) {
    ResampleCoeffs rc = resample_coeffs_public(in_size, out_size);
    dsocr_engine::DevCoef d;
    d.ksize = rc.ksize;
    d.start.alloc(rc.start.size() * 4); h2d(d.start.p, rc.start.data(), rc.start.size() * 4);
    d.len.alloc(rc.len.size() * 4); h2d(d.len.p, rc.len.data(), rc.len.size() * 4);
    d.coef.alloc(rc.coef.size() * 4); h2d(d.coef.p, rc.coef.data(), rc.coef.size() * 4);
    it = e->coefs.emplace(key, std::move(d)).first;
  }
  return it->second;
}

// resize_bicubic on the device: src [h,w,3] u8 -> either the global canvas or the tile stack (see resample_v)
void gpu_resize(dsocr_engine* e, const uint8_t* src, int w, int h, int dw, int dh, uint8_t* dst, int canvas, int x_off,
                int y_off, int tile, int tiles_w, cudaStream_t st) {
  Engine& en = *e->impl;
  const auto& cx = coef_for(e, w, dw);
  const auto& cy = coef_for(e, h, dh);
  if (e->horiz.bytes < (size_t)h * dw * 3) {
    cuda_check(cudaStreamSynchronize(en.stream()), "horiz grow sync");
    if (st != en.stream()) cuda_check(cudaStreamSynchronize(st), "horiz grow sync");
    e->horiz.alloc((size_t)h * dw * 3);
  }
  resample_h(src, w, h, e->horiz.as<uint8_t>(), dw, cx.start.as<int>(), cx.len.as<int>(), cx.coef.as<int>(), cx.ksize, st);
  resample_v(e->horiz.as<uint8_t>(), dw, dh, dst, cy.start.as<int>(), cy.len.as<int>(), cy.coef.as<int>(), cy.ksize, canvas,
             x_off, y_off, tile, tiles_w, st);
}

// prepare_vision_input_from_image for one page already resident on the device (raw RGB8)
void gpu_prepare_page(dsocr_engine* e, const uint8_t* src, int w, int h, dsocr_vision_settings vs, uint8_t* global_dst,
                      uint8_t* tiles_dst, int gw, int gh, int n_tiles, cudaStream_t st) {
  Engine& en = *e->impl;
  if (!st) st = en.stream();
  const int G = vs.crop_mode ? (int)vs.base_size : (int)vs.image_size, P = (int)vs.image_size;
  if (w == G && h == G) {
    cuda_check(cudaMemcpyAsync(global_dst, src, (size_t)G * G * 3, cudaMemcpyDeviceToDevice, st), "global copy");
  } else {
    int nw, nh, xo, yo;
    global_view_geometry(w, h, G, &nw, &nh, &xo, &yo);
    cuda_check(cudaMemsetAsync(global_dst, 127, (size_t)G * G * 3, st), "canvas fill");
    gpu_resize(e, src, w, h, nw, nh, global_dst, G, xo, yo, 0, 0, st);
  }
  if (n_tiles > 0) gpu_resize(e, src, w, h, P * gw, P * gh, tiles_dst, 0, 0, 0, P, gw, st);
}

// async_ok: the caller keeps the host pages alive until the decode that follows has finished (decode_pages / decode_requests),
// so the copies may still be in flight on return; the vision tower waits for the event covering the views of each chunk.
void stage_pages_gpu(dsocr_engine* e, int n_pages, const uint8_t* const* rgb, const int* widths, const int* heights,
                     dsocr_vision_settings vs, bool async_ok) {
  Engine& en = *e->impl;
  auto& sg = e->staged;
  const double t0 = now_ms();
  const bool async = async_ok && !kernel_timing_enabled() && !getenv("DSOCR_SYNC_STAGING");
  if (async && !e->copy_stream) cuda_check(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking), "copy stream");
  cudaStream_t st = async ? e->copy_stream : en.stream();
  e->staged_async = false;
  const int G = vs.crop_mode ? (int)vs.base_size : (int)vs.image_size, P = (int)vs.image_size;
  sg.ntiles.assign(n_pages, 0); sg.cw.assign(n_pages, 1); sg.ch.assign(n_pages, 1);
  sg.n_pages = n_pages; sg.vs = vs;
  size_t raw_bytes = 0, total_tiles = 0;
  std::vector<size_t> raw_off(n_pages);
  for (int p = 0; p < n_pages; ++p) {
    if (!rgb[p] || widths[p] <= 0 || heights[p] <= 0) throw std::runtime_error("empty image");
    raw_off[p] = raw_bytes;
    raw_bytes += ((size_t)widths[p] * heights[p] * 3 + 255) & ~(size_t)255;
    if (vs.crop_mode) sg.ntiles[p] = dynamic_preprocess_u8(rgb[p], widths[p], heights[p], P, nullptr, &sg.cw[p], &sg.ch[p]);
    total_tiles += sg.ntiles[p];
  }
  const size_t gbytes = (size_t)G * G * 3, tbytes = (size_t)P * P * 3;
  if (e->pages_raw.bytes < raw_bytes || sg.globals.bytes < gbytes * n_pages || sg.tiles.bytes < std::max<size_t>(16, tbytes * total_tiles)) {
    cuda_check(cudaStreamSynchronize(en.stream()), "stage grow sync");
    if (e->copy_stream) cuda_check(cudaStreamSynchronize(e->copy_stream), "stage grow sync");
  }
  e->pages_raw.ensure(raw_bytes);
  sg.globals.ensure(gbytes * n_pages);
  sg.tiles.ensure(std::max<size_t>(16, tbytes * total_tiles));
  constexpr int ppe = dsocr_engine::kPagesPerEvent;
  if (async) {
    const size_t n_ev = (size_t)(n_pages + ppe - 1) / ppe;
    while (e->page_events.size() < n_ev) {
      cudaEvent_t ev;
      cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "page event");
      e->page_events.push_back(ev);
    }
    e->tile_page.clear();
    for (int p = 0; p < n_pages; ++p) e->tile_page.insert(e->tile_page.end(), sg.ntiles[p], p);
  }
  std::vector<size_t> tile_off(n_pages);
  {
    size_t toff = 0;
    for (int p = 0; p < n_pages; ++p) { tile_off[p] = toff; toff += tbytes * sg.ntiles[p]; }
  }
  // pages [stage_next, upto) -> copy + resample on `st`.  In the asynchronous form this is called again and again by the
  // vision tower (a chunk ahead of what it reads): a stream accepts only ~1000 pending operations, so issuing all pages at
  // once would block the host here for the whole transfer and nothing would overlap.
  e->stage_next = 0;
  e->stage_fn = [=](int upto) {
    auto& sgr = e->staged;
    upto = std::min(upto, n_pages);
    for (int p = e->stage_next; p < upto; ++p) {
      uint8_t* src = e->pages_raw.as<uint8_t>() + raw_off[p];
      cuda_check(cudaMemcpyAsync(src, rgb[p], (size_t)widths[p] * heights[p] * 3, cudaMemcpyHostToDevice, st), "page H2D");
      gpu_prepare_page(e, src, widths[p], heights[p], vs, sgr.globals.as<uint8_t>() + gbytes * p, sgr.tiles.as<uint8_t>() + tile_off[p],
                       sgr.cw[p], sgr.ch[p], sgr.ntiles[p], st);
      if (async && ((p + 1) % ppe == 0 || p + 1 == n_pages)) cuda_check(cudaEventRecord(e->page_events[p / ppe], st), "page event record");
    }
    e->stage_next = std::max(e->stage_next, upto);
  };
  if (async) {
    e->stage_fn(dsocr_engine::kStageAhead / 2);  // enough for the first chunk of global views; the hook keeps it topped up
    e->staged_async = true;
  } else {
    e->stage_fn(n_pages);
    e->stage_fn = nullptr;
    cuda_check(cudaStreamSynchronize(st), "stage sync");
  }
  en.timings.prepare = now_ms() - t0;
}

void stage_pages(dsocr_engine* e, int n_pages, const uint8_t* const* rgb, const int* widths, const int* heights,
                 dsocr_vision_settings vs, bool async_ok = false) {
  e->staged_async = false;
  if (!e->host_preprocess) { stage_pages_gpu(e, n_pages, rgb, widths, heights, vs, async_ok); return; }
  Engine& en = *e->impl;
  auto& sg = e->staged;
  // prepare_vision_inputs (model/mod.rs:2457-2492): integer resample / tiling on the host cores
  const double t0 = now_ms();
  const int G = vs.crop_mode ? (int)vs.base_size : (int)vs.image_size, P = (int)vs.image_size;
  std::vector<std::vector<uint8_t>> globals(n_pages), tiles(n_pages);
  sg.ntiles.assign(n_pages, 0); sg.cw.assign(n_pages, 1); sg.ch.assign(n_pages, 1);
  sg.n_pages = n_pages; sg.vs = vs;
  std::vector<const uint8_t*> gsrc(n_pages, nullptr);
  {
    std::vector<std::thread> th;
    std::vector<std::string> errs(n_pages);
    const int nthreads = std::min<int>(n_pages, std::max(1u, std::thread::hardware_concurrency()));
    for (int t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        for (int p = t; p < n_pages; p += nthreads) {
          try {
            if (!rgb[p] || widths[p] <= 0 || heights[p] <= 0) throw std::runtime_error("empty image");
            if (widths[p] == G && heights[p] == G) {
              gsrc[p] = rgb[p];  // resize to the same size and paste at (0,0) is the identity
            } else {
              globals[p].resize((size_t)G * G * 3);
              build_global_view_u8(rgb[p], widths[p], heights[p], G, globals[p].data());
              gsrc[p] = globals[p].data();
            }
            if (vs.crop_mode) {
              int n = dynamic_preprocess_u8(rgb[p], widths[p], heights[p], P, nullptr, &sg.cw[p], &sg.ch[p]);
              if (n > 0) {
                tiles[p].resize((size_t)n * P * P * 3);
                dynamic_preprocess_u8(rgb[p], widths[p], heights[p], P, tiles[p].data(), &sg.cw[p], &sg.ch[p]);
              }
              sg.ntiles[p] = n;
            }
          } catch (const std::exception& ex) { errs[p] = ex.what(); }
        }
      });
    for (auto& t : th) t.join();
    for (auto& s : errs) if (!s.empty()) throw std::runtime_error(s);
  }
  const size_t gbytes = (size_t)G * G * 3, tbytes = (size_t)P * P * 3;
  size_t total_tiles = 0;
  for (int p = 0; p < n_pages; ++p) total_tiles += sg.ntiles[p];
  sg.globals.ensure(gbytes * n_pages);
  sg.tiles.ensure(std::max<size_t>(16, tbytes * total_tiles));
  size_t toff = 0;
  for (int p = 0; p < n_pages; ++p) {
    cuda_check(cudaMemcpyAsync((uint8_t*)sg.globals.p + gbytes * p, gsrc[p], gbytes, cudaMemcpyHostToDevice, en.stream()), "global view H2D");
    if (sg.ntiles[p] > 0) {
      cuda_check(cudaMemcpyAsync((uint8_t*)sg.tiles.p + toff, tiles[p].data(), tbytes * sg.ntiles[p], cudaMemcpyHostToDevice, en.stream()), "tiles H2D");
      toff += tbytes * sg.ntiles[p];
    }
  }
  cuda_check(cudaStreamSynchronize(en.stream()), "stage sync");
  en.timings.prepare = now_ms() - t0;
}

// One request = the reference's `decode(prompt, images)` (model/mod.rs:2370-2454): n_images staged views in a row and
// n_images + 1 already-tokenised text segments around the <image> slots.
struct RequestSpec {
  int n_images = 0;
  std::vector<const int64_t*> seg;
  std::vector<int> seg_len;
};

// Requests [r0, r1) of a staged set: vision tower over their images, prompt build, generate.
void decode_request_group(dsocr_engine* e, const std::vector<RequestSpec>& reqs, int r0, int r1, int img0, size_t tile0,
                          int64_t image_token_id, const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                          int64_t* const* out_tokens, int* n_out, int* prompt_tokens, std::string& stage, Timings& total) {
  Engine& en = *e->impl;
  auto& sg = e->staged;
  const int n_req = r1 - r0;
  int n_images = 0;
  for (int r = r0; r < r1; ++r) n_images += reqs[r].n_images;
  const dsocr_vision_settings vs = sg.vs;
  const int G = vs.crop_mode ? (int)vs.base_size : (int)vs.image_size, P = (int)vs.image_size;
  const size_t gbytes = (size_t)G * G * 3, tbytes = (size_t)P * P * 3;
  // ---- compute_image_embeddings (model/mod.rs:2494-2534): nothing to do for text-only prompts
  stage = "image embedding failed";
  const double t1 = now_ms();
  std::vector<int> counts;
  const float* rows = nullptr;
  if (n_images > 0) {
    std::vector<Engine::PageViews> pages(n_images);
    for (int p = 0; p < n_images; ++p) { pages[p].n_tiles = sg.ntiles[img0 + p]; pages[p].crop_w = sg.cw[img0 + p]; pages[p].crop_h = sg.ch[img0 + p]; }
    if (e->staged_async) {
      // pages are still arriving on the copy stream: before a chunk of views is read, wait for the event that covers
      // the last page the chunk touches (copies and resamples complete in page order)
      en.view_ready = [e, &en, img0, tile0](bool local, int last_view) {
        const int page = local ? e->tile_page[tile0 + last_view] : img0 + last_view;
        if (e->stage_fn) e->stage_fn((page / dsocr_engine::kPagesPerEvent + 1) * dsocr_engine::kPagesPerEvent + dsocr_engine::kStageAhead);
        cuda_check(cudaStreamWaitEvent(en.stream(), e->page_events[page / dsocr_engine::kPagesPerEvent], 0), "page event wait");
      };
    }
    try {
      rows = en.vision_encode(n_images, sg.globals.as<uint8_t>() + gbytes * img0, false, G, sg.tiles.as<uint8_t>() + tbytes * tile0,
                              false, P, pages, &counts);
    } catch (...) {
      en.view_ready = nullptr;
      throw;
    }
    en.view_ready = nullptr;
    cuda_check(cudaStreamSynchronize(en.stream()), "vision sync");
  }
  total.vision += now_ms() - t1;
  // ---- build_prompt_tokens (model/mod.rs:2536-2603): BOS + seg[0] + <image> x n_0 + seg[1] + ...
  stage = "prompt formatting failed";
  std::vector<std::vector<int64_t>> ids(n_req);
  std::vector<std::vector<uint8_t>> masks(n_req);
  std::vector<const int64_t*> idp(n_req);
  std::vector<const uint8_t*> mp(n_req);
  std::vector<int> nt(n_req), n_img_rows(n_req, 0);
  int img = 0;
  for (int r = 0; r < n_req; ++r) {
    const RequestSpec& rq = reqs[r0 + r];
    const int slots = std::max(0, (int)rq.seg.size() - 1);
    if (slots != rq.n_images)
      throw std::runtime_error("prompt/image embedding mismatch: " + std::to_string(slots) + " slots vs " +
                               std::to_string(rq.n_images) + " embeddings");
    ids[r].push_back(0); masks[r].push_back(0);  // bos_id = 0 (model/mod.rs:2547)
    for (size_t sidx = 0; sidx < rq.seg.size(); ++sidx) {
      for (int i = 0; i < rq.seg_len[sidx]; ++i) { ids[r].push_back(rq.seg[sidx][i]); masks[r].push_back(0); }
      if ((int)sidx < rq.n_images) {
        const int expect = image_token_count((int)vs.base_size, (int)vs.image_size, vs.crop_mode, sg.cw[img0 + img], sg.ch[img0 + img]);
        if (expect != counts[img])
          throw std::runtime_error("placeholder count " + std::to_string(expect) + " does not match expected " + std::to_string(counts[img]));
        for (int i = 0; i < counts[img]; ++i) { ids[r].push_back(image_token_id); masks[r].push_back(1); }
        n_img_rows[r] += counts[img];
        ++img;
      }
    }
    idp[r] = ids[r].data(); mp[r] = masks[r].data(); nt[r] = (int)ids[r].size();
    if (prompt_tokens) prompt_tokens[r0 + r] = nt[r];
  }
  stage = "";
  Engine::GenRequest rq;
  rq.n_pages = n_req; rq.input_ids = idp.data(); rq.mask = mp.data(); rq.n_tokens = nt.data();
  rq.image_rows_dev = rows; rq.n_image_rows = n_img_rows.data(); rq.params = *params; rq.cb = cb; rq.user = user;
  rq.page_offset = r0;
  en.generate(rq, out_tokens + r0, n_out + r0);
  total.prefill += en.timings.prefill; total.iterative += en.timings.iterative; total.generate += en.timings.generate;
}

// All requests of a staged set, `decode_batch` requests per lock-step group (option "decode_batch", default 512): the
// decode step's weight traffic is shared by the pages of a group, its KV / activation workspaces grow with the group.
void decode_staged_requests(dsocr_engine* e, const std::vector<RequestSpec>& reqs, int64_t image_token_id,
                            const dsocr_decode_params* params, dsocr_token_cb cb, void* user, int64_t* const* out_tokens,
                            int* n_out, int* prompt_tokens, std::string& stage) {
  Engine& en = *e->impl;
  auto& sg = e->staged;
  const int n_req = (int)reqs.size();
  int n_images = 0;
  for (const auto& r : reqs) n_images += r.n_images;
  if (n_images != sg.n_pages) throw std::runtime_error("staged views do not match the requests' image lists");
  Timings total;
  total.prepare = en.timings.prepare;
  const int group = std::max(1, e->decode_batch);
  int img0 = 0;
  size_t tile0 = 0;
  for (int r0 = 0; r0 < n_req; r0 += group) {
    const int r1 = std::min(n_req, r0 + group);
    decode_request_group(e, reqs, r0, r1, img0, tile0, image_token_id, params, cb, user, out_tokens, n_out, prompt_tokens, stage, total);
    for (int r = r0; r < r1; ++r)
      for (int i = 0; i < reqs[r].n_images; ++i) tile0 += sg.ntiles[img0++];
  }
  en.timings = total;
}

void decode_staged(dsocr_engine* e, const int64_t* seg0, int n_seg0, const int64_t* seg1, int n_seg1,
                   int64_t image_token_id, const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                   int64_t* const* out_tokens, int* n_out, int* prompt_tokens, std::string& stage) {
  if (e->staged.n_pages <= 0) throw std::runtime_error("no pages staged");
  RequestSpec one;
  one.n_images = 1; one.seg = {seg0, seg1}; one.seg_len = {n_seg0, n_seg1};
  std::vector<RequestSpec> reqs(e->staged.n_pages, one);
  decode_staged_requests(e, reqs, image_token_id, params, cb, user, out_tokens, n_out, prompt_tokens, stage);
}
}  // namespace

extern "C" int dsocr_stage_pages(dsocr_engine* e, int n_pages, const uint8_t* const* rgb, const int* widths,
                                 const int* heights, dsocr_vision_settings vs) {
  return api("vision input failed", [&] { bind(e); stage_pages(e, n_pages, rgb, widths, heights, vs); });
}

extern "C" int dsocr_decode_staged(dsocr_engine* e, const int64_t* seg0, int n_seg0, const int64_t* seg1, int n_seg1,
                                   int64_t image_token_id, const dsocr_decode_params* params, dsocr_token_cb cb,
                                   void* user, int64_t* const* out_tokens, int* n_out, int* prompt_tokens) {
  std::string stage;
  return api("", [&] {
    try {
      bind(e);
      if (!params) throw std::runtime_error("null decode params");
      decode_staged(e, seg0, n_seg0, seg1, n_seg1, image_token_id, params, cb, user, out_tokens, n_out, prompt_tokens, stage);
    } catch (const std::exception& ex) {
      throw std::runtime_error(stage.empty() ? std::string(ex.what()) : stage + ": " + ex.what());
    }
  });
}

extern "C" int dsocr_decode_pages(dsocr_engine* e, int n_pages, const uint8_t* const* rgb, const int* widths,
                                  const int* heights, dsocr_vision_settings vs, const int64_t* seg0, int n_seg0,
                                  const int64_t* seg1, int n_seg1, int64_t image_token_id,
                                  const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                                  int64_t* const* out_tokens, int* n_out, int* prompt_tokens) {
  std::string stage = "vision input failed";
  return api("", [&] {
    try {
      bind(e);
      if (!params) throw std::runtime_error("null decode params");
      stage_pages(e, n_pages, rgb, widths, heights, vs, /*async_ok=*/true);
      decode_staged(e, seg0, n_seg0, seg1, n_seg1, image_token_id, params, cb, user, out_tokens, n_out, prompt_tokens, stage);
      e->staged_async = false; e->stage_fn = nullptr;
    } catch (const std::exception& ex) {
      if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);  // the caller may free its pages once we return
      e->staged_async = false; e->stage_fn = nullptr;
      throw std::runtime_error(stage.empty() ? std::string(ex.what()) : stage + ": " + ex.what());
    }
  });
}

extern "C" int dsocr_decode_requests(dsocr_engine* e, int n_requests, const dsocr_request* requests, dsocr_vision_settings vs,
                                     int64_t image_token_id, const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                                     int64_t* const* out_tokens, int* n_out, int* prompt_tokens) {
  std::string stage = "vision input failed";
  return api("", [&] {
    try {
      bind(e);
      if (!params || (!requests && n_requests > 0)) throw std::runtime_error("null argument");
      std::vector<RequestSpec> reqs(n_requests);
      std::vector<const uint8_t*> rgb;
      std::vector<int> ws, hs;
      for (int r = 0; r < n_requests; ++r) {
        const dsocr_request& q = requests[r];
        if (q.n_images < 0 || q.n_segments < 0) throw std::runtime_error("negative count in request");
        reqs[r].n_images = q.n_images;
        for (int i = 0; i < q.n_images; ++i) { rgb.push_back(q.rgb[i]); ws.push_back(q.widths[i]); hs.push_back(q.heights[i]); }
        for (int i = 0; i < q.n_segments; ++i) { reqs[r].seg.push_back(q.segments[i]); reqs[r].seg_len.push_back(q.segment_lens[i]); }
      }
      if (!rgb.empty()) stage_pages(e, (int)rgb.size(), rgb.data(), ws.data(), hs.data(), vs, /*async_ok=*/true);
      else { e->staged.n_pages = 0; e->staged.vs = vs; e->impl->timings.prepare = 0; e->staged_async = false; }
      decode_staged_requests(e, reqs, image_token_id, params, cb, user, out_tokens, n_out, prompt_tokens, stage);
      e->staged_async = false; e->stage_fn = nullptr;
    } catch (const std::exception& ex) {
      if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);  // the caller may free its pages once we return
      e->staged_async = false; e->stage_fn = nullptr;
      throw std::runtime_error(stage.empty() ? std::string(ex.what()) : stage + ": " + ex.what());
    }
  });
}

// ---- expert-parallel groups (BASELINE configs[4])
struct dsocr_ep_group { std::vector<dsocr_engine*> engines; };

extern "C" int dsocr_ep_group_create(dsocr_engine* const* engines, int n, int max_pages_per_engine, dsocr_ep_group** out) {
  return api("expert-parallel group", [&] {
    if (!engines || !out || n < 2 || n > 8 || max_pages_per_engine < 1) throw std::runtime_error("invalid argument");
    for (int i = 0; i < n; ++i) if (!engines[i] || !engines[i]->impl) throw std::runtime_error("null engine handle");
    // peer mappings: every engine's device must be able to address the others' buffers (NVLink / NVSwitch P2P; the
    // same device is allowed, which is how the single-GPU tests run a 2-rank group)
    for (int i = 0; i < n; ++i) {
      const int di = engines[i]->impl->device();
      cuda_check(cudaSetDevice(di), "cudaSetDevice");
      for (int j = 0; j < n; ++j) {
        const int dj = engines[j]->impl->device();
        if (di == dj) continue;
        int ok = 0;
        cuda_check(cudaDeviceCanAccessPeer(&ok, di, dj), "cudaDeviceCanAccessPeer");
        if (!ok) throw std::runtime_error("device " + std::to_string(di) + " cannot address device " + std::to_string(dj) + " (no P2P)");
        const cudaError_t st = cudaDeviceEnablePeerAccess(dj, 0);
        if (st != cudaSuccess && st != cudaErrorPeerAccessAlreadyEnabled) cuda_check(st, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
      }
    }
    const int cap = max_pages_per_engine * n;
    std::vector<Engine::EpBuffers> bufs(n);
    for (int i = 0; i < n; ++i) bufs[i] = engines[i]->impl->ep_attach(i, n, cap);
    for (int i = 0; i < n; ++i) {
      EpPeers p;
      p.world = n; p.rank = i; p.eloc = engines[i]->impl->cfg().n_experts / n; p.cap = cap;
      for (int j = 0; j < n; ++j) { p.counts[j] = bufs[j].counts; p.xperm[j] = bufs[j].xperm; p.y[j] = bufs[j].y; p.flags[j] = bufs[j].flags; }
      engines[i]->impl->ep_set_peers(p);
    }
    auto g = std::make_unique<dsocr_ep_group>();
    g->engines.assign(engines, engines + n);
    *out = g.release();
  });
}

extern "C" void dsocr_ep_group_destroy(dsocr_ep_group* g) {
  if (!g) return;
  for (dsocr_engine* e : g->engines) if (e && e->impl) e->impl->ep_detach();
  delete g;
}

extern "C" int dsocr_engine_set_stream(dsocr_engine* e, void* cuda_stream) {
  return api("", [&] { bind(e); e->impl->set_stream(reinterpret_cast<cudaStream_t>(cuda_stream)); });
}

extern "C" int dsocr_kernel_timing_begin(dsocr_engine* e) {
  return api("", [&] { bind(e); kernel_timing_begin(e->impl->stream()); });
}

extern "C" int dsocr_kernel_timing_end(dsocr_engine* e, char* json_out, size_t capacity) {
  return api("", [&] {
    bind(e);
    const std::string js = kernel_timing_end_json();
    if (!json_out || capacity < js.size() + 1) throw std::runtime_error("timing report buffer too small");
    memcpy(json_out, js.c_str(), js.size() + 1);
  });
}

extern "C" int dsocr_last_timings(const dsocr_engine* e, double* ms_out, int n) {
  return api("", [&] {
    if (!e || !ms_out) throw std::runtime_error("null argument");
    const Timings& t = e->impl->timings;
    const double v[5] = {t.prepare, t.vision, t.prefill, t.iterative, t.generate};
    for (int i = 0; i < n && i < 5; ++i) ms_out[i] = v[i];
  });
}

extern "C" int dsocr_moe_stats(dsocr_engine* e, double* out2) {
  return api("", [&] {
    if (!e || !out2) throw std::runtime_error("null argument");
    bind(e);
    unsigned long long v[2];
    e->impl->moe_stats(v);
    out2[0] = (double)v[0]; out2[1] = (double)v[1];
  });
}

extern "C" long long dsocr_launch_count(const dsocr_engine*) { return launch_counter().load(); }
