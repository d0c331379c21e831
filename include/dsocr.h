/* dsocr.h - C ABI of the B200-native DeepSeek-OCR per-page forward path.
 *
 * This is the drop-in boundary for TimmyOVO/deepseek-ocr.rs's `OcrEngine` / `load_model`
 * (crates/core/src/inference.rs:179-209, crates/infer-deepseek/src/model/mod.rs:90-115).  A Rust shim
 * implements `OcrEngine::decode` (model/mod.rs:2370-2454) on top of these calls; see INTEGRATION.md.
 *
 * Conventions
 *  - plain C, opaque handles, no exceptions cross the boundary;
 *  - every call returns DSOCR_OK (0) or a negative status; the message of the last failure on the calling
 *    thread is returned by dsocr_last_error() (same role as the reference's anyhow context chain; the
 *    literal prefixes "vision input failed", "image embedding failed", "prompt formatting failed" and
 *    "prompt/image embedding mismatch" are preserved because crates/server/src/generation.rs:111-115
 *    pattern-matches them);
 *  - host buffers are caller-owned, device memory is engine-owned;
 *  - one engine is bound to one CUDA device ordinal and is NOT re-entrant (the reference wraps its engine in
 *    Arc<Mutex<..>>, crates/server/src/generation.rs:84-86): serialise calls per engine;
 *  - there is no CPU fallback: engine creation fails if no sm_100 device is present.
 */
#ifndef DSOCR_H_
#define DSOCR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DSOCR_API __attribute__((visibility("default")))
#else
#define DSOCR_API
#endif

typedef struct dsocr_engine dsocr_engine;

enum dsocr_status {
  DSOCR_OK = 0,
  DSOCR_ERR_INVALID_ARGUMENT = -1,
  DSOCR_ERR_IO = -2,
  DSOCR_ERR_CUDA = -3,
  DSOCR_ERR_UNSUPPORTED = -4,
  DSOCR_ERR_INTERNAL = -5,
  DSOCR_ERR_MISMATCH = -6 /* prompt/image embedding mismatch (maps to HTTP 400 in the reference server) */
};

/* candle DType as used by ModelLoadArgs.dtype (crates/core/src/inference.rs:179-186). */
enum dsocr_dtype { DSOCR_F32 = 0, DSOCR_F16 = 1, DSOCR_BF16 = 2 };

/* VisionSettings (crates/core/src/inference.rs:10-16). */
typedef struct dsocr_vision_settings {
  uint32_t base_size;  /* 1024 */
  uint32_t image_size; /* 640  */
  int32_t crop_mode;   /* 1    */
} dsocr_vision_settings;

/* DecodeParameters (crates/core/src/inference.rs:18-34, defaults :66-78).  Greedy selection (do_sample == 0 or
 * temperature <= 0: sampling.rs:62 falls through to the argmax) runs on the device, including the repetition
 * penalty and the no-repeat-n-gram ban.  With do_sample != 0 and temperature > 0 every step's logits rows are copied
 * to the host and crates/core/src/sampling.rs:34-96 is restated there (temperature, top-k, top-p, StdRng seeded with
 * `seed`: ChaCha12 + WeightedIndex, the draw sequence of rand 0.8); each page of a batched call owns an RNG seeded
 * like the reference's single-page call. */
typedef struct dsocr_decode_params {
  uint32_t max_new_tokens;       /* 512 */
  int32_t do_sample;             /* 0 */
  float repetition_penalty;      /* 1.0; <= 0 or within f32 epsilon of 1 disables (sampling.rs:121-123) */
  uint32_t no_repeat_ngram_size; /* 20; 0 or 1 disables */
  int64_t eos_token_id;          /* < 0 disables EOS stopping */
  int32_t use_cache;             /* 1 (0 is accepted: the cached path computes the same tokens) */
  int32_t has_seed;              /* 0: seed from the OS entropy source, like StdRng::from_entropy */
  double temperature;            /* 0.0 */
  double top_p;                  /* Option<f64>: < 0 = None; only values in [0, 1) filter (sampling.rs:74) */
  uint32_t top_k;                /* Option<usize>: 0 = None */
  uint32_t reserved_;
  uint64_t seed;
} dsocr_decode_params;

/* Progress callback == the reference's `stream: Option<&dyn Fn(usize, &[i64])>` (inference.rs:205-207):
 * invoked synchronously on the calling thread after every accepted token with (count, all generated ids).
 * `page` is the index inside a batched call (0 for single-page calls).  When a callback is set the decode loop
 * synchronises after every step, so each page's callbacks arrive one token at a time as in model/mod.rs:1978-1982. */
typedef void (*dsocr_token_cb)(void* user, int32_t page, size_t count, const int64_t* tokens);

typedef struct dsocr_engine_info {
  int32_t device_ordinal;
  int32_t dtype; /* dsocr_dtype of the 16-bit operand type */
  int32_t sm_count;
  int32_t hidden_size, num_layers, vocab_size, n_routed_experts;
  int32_t quantized; /* 1 when a DSQ snapshot is attached */
  char device_name[64];
} dsocr_engine_info;

DSOCR_API const char* dsocr_last_error(void);
DSOCR_API const char* dsocr_version(void);

/* load_model(ModelLoadArgs{config_path, weights_path, snapshot_path, device, dtype}) -> Box<dyn OcrEngine>
 * (model/mod.rs:90-115, :946-1105).  dsq_path may be NULL.  dtype: DSOCR_F16 or DSOCR_BF16 select the
 * tensor-core operand type; as in the reference, accumulation / norms / softmax / residuals stay f32. */
DSOCR_API int dsocr_engine_create(const char* config_json_path, const char* safetensors_path, const char* dsq_path,
                                  int device_ordinal, int dtype, dsocr_engine** out);
DSOCR_API void dsocr_engine_destroy(dsocr_engine* e);
DSOCR_API int dsocr_engine_info_get(const dsocr_engine* e, dsocr_engine_info* info);
/* Engine options: "record_taps" (0/1) keeps host copies of the debug-trace taps of the next vision call;
 * "kv_cache_f16" (0/1) stores the KV cache in f16 instead of the reference's f32 (model/mod.rs:82-88): half the
 * decode-attention bytes, K/V rounded to 11 bits (off by default; parity numbers are quoted for both);
 * "host_preprocess" (0/1) runs the integer resample / tiling on the host cores instead of the device;
 * "decode_batch" (n >= 1, default 512): dsocr_decode_pages / _staged / _requests split their pages into lock-step groups
 * of at most n (vision + prefill + token loop per group); results do not depend on it. */
DSOCR_API int dsocr_engine_set_option(dsocr_engine* e, const char* name, int value);

/* DsqReader::open + header() / records() (crates/dsq/src/lib.rs:208-262): maps a `.dsq` snapshot, validates header and
 * records exactly as the reference does (same order, same messages through dsocr_last_error) and reports the records.
 * Host only - needs no GPU.  `records` may be NULL (count only); at most `capacity` records are written. */
typedef struct dsocr_dsq_record {
  char name[192];
  uint32_t out_dim, in_dim;
  uint32_t q_dtype; /* 8 Q8_0, 12 Q4_K, 14 Q6_K, 1 F16, 16 BF16, 0 F32 */
  uint64_t q_offset, q_len;
  uint64_t bias_offset, bias_len; /* bias_len == 0: no bias */
  uint32_t bias_dtype;
  uint8_t first_q_byte; /* tensor_bytes(record)[0] */
} dsocr_dsq_record;
typedef struct dsocr_dsq_header {
  uint32_t version, default_qdtype, block_size, tensor_count;
  char candle_version[64], model_id[128], backend[32];
} dsocr_dsq_header;
DSOCR_API int dsocr_dsq_inspect(const char* path, dsocr_dsq_header* header, dsocr_dsq_record* records, size_t capacity);

/* DsqWriter (crates/dsq-writer/src/lib.rs:93-527), host only: `create` (the path gets the `.dsq` extension, as
 * Path::with_extension does), `add_tensor` quantises a row-major f32 [out_dim, in_dim] matrix to q_dtype (8 Q8_0 with the
 * reference's quantize_q8_0; 12 / 14 Q4_K / Q6_K with ggml-style reference quantisers - valid blocks, not byte-pinned to
 * candle's from_float; 0 / 1 / 16 float payloads), `add_quantized_bytes` == the reference method of that name,
 * `finalize` writes the file and frees the handle.  bias: out_dim f32 values or NULL. */
typedef struct dsocr_dsq_writer dsocr_dsq_writer;
DSOCR_API int dsocr_dsq_writer_create(const char* path, const char* candle_version, const char* model_id, const char* backend,
                                      uint32_t default_qdtype, dsocr_dsq_writer** out);
DSOCR_API int dsocr_dsq_writer_add_tensor(dsocr_dsq_writer* w, const char* name, uint32_t out_dim, uint32_t in_dim, uint32_t q_dtype,
                                          const float* weights, const float* bias);
DSOCR_API int dsocr_dsq_writer_add_quantized_bytes(dsocr_dsq_writer* w, const char* name, uint32_t out_dim, uint32_t in_dim,
                                                   uint32_t q_dtype, const uint8_t* qbytes, size_t q_len, const float* bias);
DSOCR_API int dsocr_dsq_writer_finalize(dsocr_dsq_writer* w);
DSOCR_API void dsocr_dsq_writer_destroy(dsocr_dsq_writer* w);

/* image_token_count: rows `compute_image_embeddings` will produce == placeholders
 * `build_image_placeholders` emits (model/mod.rs:2605-2689). */
DSOCR_API int dsocr_image_token_count(uint32_t base_size, uint32_t image_size, int crop_mode, int crop_w, int crop_h);

/* prepare_vision_input_from_image (model/mod.rs:1707-1758) + build_global_view (:2308-2330) +
 * dynamic_preprocess_with_params (vision/preprocess.rs:67-138) + resize_bicubic (vision/resample.rs:101-160).
 * Input: RGB8 HWC.  Outputs (caller-allocated): global view RGB8 [G,G,3] (G = base_size if crop_mode else
 * image_size); tiles RGB8 [n,image_size,image_size,3] with n <= 9 (may be NULL to only query the grid);
 * crop grid (w,h).  Bit-exact with the reference's integer resampler. */
DSOCR_API int dsocr_preprocess(const uint8_t* rgb, int width, int height, dsocr_vision_settings vs,
                               uint8_t* global_out, uint8_t* tiles_out, int* n_tiles, int* crop_w, int* crop_h);

/* Same as dsocr_preprocess but with the resample / tiling kernels of the engine's device (the path
 * dsocr_stage_pages / dsocr_decode_pages use; option "host_preprocess" = 1 selects the host cores instead).
 * Bit-exact with dsocr_preprocess. */
DSOCR_API int dsocr_preprocess_gpu(dsocr_engine* e, const uint8_t* rgb, int width, int height, dsocr_vision_settings vs,
                                   uint8_t* global_out, uint8_t* tiles_out, int* n_tiles, int* crop_w, int* crop_h);

/* compute_image_embeddings for one page (model/mod.rs:1276-1377; VisionContext :711-924).
 * global_chw: f32 [3,G,G] normalised as image_to_tensor does (:2332-2347); patches_nchw: f32 [n,3,P,P] or NULL.
 * out_rows: f32 [n_rows,hidden] = [local ; global ; view_separator].  *n_rows in: capacity, out: rows written. */
DSOCR_API int dsocr_vision_encode(dsocr_engine* e, const float* global_chw, int global_size, const float* patches_nchw,
                                  int n_patches, int patch_size, int crop_w, int crop_h, float* out_rows, int* n_rows);

/* Same, from RGB8 HWC views (normalisation fused into the patch gather on the device) and for a batch of
 * pages that share one (global_size, patch_size) setting.  tiles of page i: tiles_u8[i] = [n_i,P,P,3] or NULL. */
DSOCR_API int dsocr_vision_encode_u8_batch(dsocr_engine* e, int n_pages, const uint8_t* const* globals_u8,
                                           int global_size, const uint8_t* const* tiles_u8, const int* n_tiles,
                                           int patch_size, const int* crop_w, const int* crop_h,
                                           float* const* out_rows, int* n_rows);

/* Debug taps == SamDebugTrace / ClipDebugTrace / VisionProjectionOutputs (vision/sam.rs:130-140,
 * vision/clip.rs:65-70, model/mod.rs:144-153) of the most recent vision call.  Copies tap `name`
 * ("sam.patch_embed", "sam.block.3", "sam.net3", "clip.layer.7", "global_pre", ...) into out (f32). */
DSOCR_API int dsocr_vision_tap(dsocr_engine* e, const char* name, float* out, size_t capacity, size_t* n_written);

/* DeepseekOcrModel::generate (model/mod.rs:1870-2048) for a batch of independent pages decoded in lock-step:
 * prefill (embed_tokens + inject_image_tokens :1760-1857 + 12 decoder layers + last-row lm_head), then the
 * greedy loop with the no-repeat-ngram ban and first-index argmax of crates/core/src/sampling.rs:34-158,
 * stopping per page at EOS (not appended) or at max_new_tokens.
 *  input_ids[i]: i64 [T_i]; images_seq_mask[i]: u8 [T_i]; image_rows[i]: f32 [n_img_i,hidden] (host).
 *  out_tokens[i]: i64 capacity max_new_tokens; n_out[i]: generated count. */
DSOCR_API int dsocr_generate_batch(dsocr_engine* e, int n_pages, const int64_t* const* input_ids,
                                   const uint8_t* const* images_seq_mask, const int* n_tokens,
                                   const float* const* image_rows, const int* n_image_rows,
                                   const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                                   int64_t* const* out_tokens, int* n_out);

/* Teacher-forced variant used by parity tests (== tests/baseline.rs "teacher-forced logits"): feeds
 * forced_tokens[i][t] instead of the selected token, returns the selected tokens and, when logits_out[i] is
 * non-NULL, the f32 logits of every step [n_steps, vocab]. */
DSOCR_API int dsocr_generate_forced(dsocr_engine* e, int n_pages, const int64_t* const* input_ids,
                                    const uint8_t* const* images_seq_mask, const int* n_tokens,
                                    const float* const* image_rows, const int* n_image_rows,
                                    const dsocr_decode_params* params, const int64_t* const* forced_tokens,
                                    int n_steps, int64_t* const* selected_out, float* const* logits_out);

/* OcrEngine::decode minus tokenizer (model/mod.rs:2370-2454): RGB8 pages -> preprocess -> vision ->
 * build_prompt_tokens (:2536-2603; text segments are passed already tokenised: the tokenizer stays in the
 * host language) -> generate.  One image per page with the text segments seg0 / seg1 around its single <image> slot
 * (dsocr_decode_requests takes any number of images per prompt).  This is the end-to-end call the throughput
 * benchmark times. */
DSOCR_API int dsocr_decode_pages(dsocr_engine* e, int n_pages, const uint8_t* const* rgb, const int* widths,
                                 const int* heights, dsocr_vision_settings vs, const int64_t* seg0, int n_seg0,
                                 const int64_t* seg1, int n_seg1, int64_t image_token_id,
                                 const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                                 int64_t* const* out_tokens, int* n_out, int* prompt_tokens);

/* OcrEngine::decode (model/mod.rs:2370-2454) in full generality, for a batch of independent requests: request r has
 * images[n_images] (prepare_vision_inputs :2457-2492 maps over them; 0 images = text-only prompt, the vision tower is
 * skipped) and the n_segments = n_images + 1 text segments of `prompt.split("<image>")` (:2551), already tokenised.
 * Slot i receives the placeholders / embedding rows of image i (:2573-2592).  A segment count that does not match
 * fails with "prompt formatting failed: prompt/image embedding mismatch: S slots vs N embeddings" (DSOCR_ERR_MISMATCH).
 * Outputs are indexed by request; the callback's `page` is the request index. */
typedef struct dsocr_request {
  int32_t n_images;
  const uint8_t* const* rgb;    /* [n_images] RGB8 HWC */
  const int32_t* widths;        /* [n_images] */
  const int32_t* heights;       /* [n_images] */
  int32_t n_segments;
  const int64_t* const* segments; /* [n_segments] token ids of each text segment (may be empty) */
  const int32_t* segment_lens;    /* [n_segments] */
} dsocr_request;
DSOCR_API int dsocr_decode_requests(dsocr_engine* e, int n_requests, const dsocr_request* requests, dsocr_vision_settings vs,
                                    int64_t image_token_id, const dsocr_decode_params* params, dsocr_token_cb cb, void* user,
                                    int64_t* const* out_tokens, int* n_out, int* prompt_tokens);

/* The two halves of dsocr_decode_pages, for callers that keep pages resident on the device:
 * dsocr_stage_pages = host integer preprocessing + host->device copy of the RGB8 views (engine-owned);
 * dsocr_decode_staged = vision + prompt build + generate on the staged views (no image H2D traffic). */
DSOCR_API int dsocr_stage_pages(dsocr_engine* e, int n_pages, const uint8_t* const* rgb, const int* widths,
                                const int* heights, dsocr_vision_settings vs);
DSOCR_API int dsocr_decode_staged(dsocr_engine* e, const int64_t* seg0, int n_seg0, const int64_t* seg1, int n_seg1,
                                  int64_t image_token_id, const dsocr_decode_params* params, dsocr_token_cb cb,
                                  void* user, int64_t* const* out_tokens, int* n_out, int* prompt_tokens);

/* Expert-parallel decode (BASELINE.json configs[4]; the reference rejects ep_size > 1, block.rs:1248-1252, so this has no
 * counterpart to mirror): n engines of ONE process, one per GPU, form a group in which rank r computes the routed experts
 * [r*E/n, (r+1)*E/n) for the tokens of every rank during the batched decode steps.  Pages stay data-parallel (each engine
 * decodes its own pages, attention / dense layers / shared experts / lm_head local); per MoE layer the token rows go to the
 * owners' expert segments by peer stores over NVLink (slot reservation = system-scope atomic on the owner's counter) and the
 * expert outputs come back by peer loads, with flag barriers between the phases - no NCCL on the data path.  While a group
 * exists its engines must be driven in lock-step: dsocr_generate_batch / dsocr_decode_* are called on all n engines
 * concurrently (one host thread each) with the same max_new_tokens, at most max_pages_per_engine pages each and more than
 * 4 pages per call; steps are not cut short when an engine's pages all hit EOS.  Prefill stays local. */
typedef struct dsocr_ep_group dsocr_ep_group;
DSOCR_API int dsocr_ep_group_create(dsocr_engine* const* engines, int n, int max_pages_per_engine, dsocr_ep_group** out);
DSOCR_API void dsocr_ep_group_destroy(dsocr_ep_group* g);

/* Run all engine work on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. torch's current
 * stream, so that the caller's CUDA events bracket the engine's kernels. */
DSOCR_API int dsocr_engine_set_stream(dsocr_engine* e, void* cuda_stream);

/* Per-kernel device timing (CUDA events after every launch while enabled).  dsocr_kernel_timing_end writes a
 * JSON array [{"name": "decode/lm_head", "launches": n, "ms": t}, ...] into json_out. */
DSOCR_API int dsocr_kernel_timing_begin(dsocr_engine* e);
DSOCR_API int dsocr_kernel_timing_end(dsocr_engine* e, char* json_out, size_t capacity);

/* Stage timings of the most recent dsocr_decode_pages / vision / generate call, in milliseconds, under the
 * reference's Timer names (crates/core/src/benchmark.rs):
 * 0 vision.prepare_inputs, 1 vision.compute_embeddings, 2 decode.prefill, 3 decode.iterative, 4 decode.generate */
DSOCR_API int dsocr_last_timings(const dsocr_engine* e, double* ms_out, int n);

/* Diagnostics for the roofline model (no reference counterpart): with option "moe_stats" = 1 every batched decode
 * step adds the number of routed-expert weight segments it actually touched (non-empty (layer, expert) pairs) to
 * out2[0] and 1 to out2[1].  Reading resets both. */
DSOCR_API int dsocr_moe_stats(dsocr_engine* e, double* out2);

/* Number of kernels this library launched on the engine's streams since creation. */
DSOCR_API long long dsocr_launch_count(const dsocr_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* DSOCR_H_ */
