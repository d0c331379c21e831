/* dsocr_test.h - kernel-level test hooks of libdsocr.so (used by tests/ only).
 * Host f32 in, host f32 out; the hook rounds operands to the 16-bit tensor-core type, runs the CUDA kernel
 * the engine uses for that op on device 0 and copies the result back.  Not part of the drop-in boundary. */
#ifndef DSOCR_TEST_H_
#define DSOCR_TEST_H_
#include "dsocr.h"

#ifdef __cplusplus
extern "C" {
#endif

/* out[M,N] = epi(x[M,K] * w0[N,K]^T (+bias)); w1 != NULL -> silu(x*w0^T) * (x*w1^T).
 * act: 0 none, 1 gelu(erf), 2 quick-gelu.  out_mode: 0 16-bit, 1 16-bit hi+lo (returned summed), 2 f32,
 * 3 f32 accumulate into the given `out` contents.  x_parts: 1, or 2 = hi/lo split activations.
 * bn: token tile (0 = auto).  row_map: optional int[M]. */
DSOCR_API int dsocr_test_linear(int dtype, int M, int N, int K, const float* x, const float* w0, const float* w1,
                                const float* bias, int act, int out_mode, int x_parts, int bn, const int* row_map,
                                int out_rows, float* out);

/* Grouped linear (MoE expert GEMM): x[M,K] rows are grouped by expert; counts[E] rows per expert (sum = M);
 * w0/w1: [E,N,K].  out[M,N]. */
DSOCR_API int dsocr_test_grouped_linear(int dtype, int E, int M, int N, int K, const int* counts, const float* x,
                                        const float* w0, const float* w1, int x_parts, float* out);

/* Decode-time MoE expert GEMM with fixed-capacity segments: expert e owns rows [e*cap, e*cap + counts[e]) of
 * x[E*cap, K] / out[E*cap, N]; the kernel enumerates the non-empty (expert, chunk, block) units on the device.
 * Rows past counts[e] are left untouched (zero). */
DSOCR_API int dsocr_test_fixedcap_linear(int dtype, int E, int cap, int N, int K, const int* counts, const float* x,
                                         const float* w0, const float* w1, int x_parts, float* out);

/* Vision attention over a qkv buffer [B*S, 3, H, 64] (16-bit after rounding) with the decomposed
 * relative-position bias: rel_h/rel_w tables [2*g-1... already resolved to [g, g, 64]] or NULL (CLIP).
 * grid_w * grid_h == S when tables are given.  out[B*S, H*64] f32. */
DSOCR_API int dsocr_test_vision_attention(int dtype, int B, int S, int H, const float* qkv, int grid, const float* rel_h,
                                          const float* rel_w, int rel_rows, float* out);

/* Host-only: repacks on-disk ggml blocks (rows x K/blk blocks of q_dtype 8 / 12 / 14) into the device plane layouts and
 * dequantises every 64-wide k-block with dsq_dequant64 (csrc/dsq_dequant.h), the routine the dequant-fused GEMM's
 * producer threads run.  out[rows, K] f32. */
DSOCR_API int dsocr_test_dsq_dequant64(uint32_t q_dtype, const uint8_t* blocks, int rows, int K, float* out);

/* Dequant-fused tensor-core GEMM (csrc/linear_dq.cuh): out[M,N] = x[M,K] * dequant(W)[N,K]^T, or with blocks1 != NULL
 * silu(x.W0^T) * (x.W1^T).  blocks / blocks1: on-disk ggml blocks (q_dtype 8 / 12 / 14) of `groups` stacked [N,K] matrices;
 * groups > 1: x rows are grouped, counts[g] rows use matrix g (the table-grouped MoE form).  x is split into hi + lo
 * 16-bit parts like the decoder's activations.  k_splits > 1: deterministic split-K, the f32 partials are summed here. */
DSOCR_API int dsocr_test_linear_dq(int dtype, uint32_t q_dtype, int groups, const int* counts, int M, int N, int K,
                                   const uint8_t* blocks, const uint8_t* blocks1, const float* x, int bn, int k_splits,
                                   float* out);

/* Host-only hooks of the sampling path (csrc/sampler.cpp; sampling.rs:34-96 + rand 0.8 StdRng).
 * dsocr_test_chacha_words: first n keystream words of the block RNG for a 32-byte key and a round count (12 = StdRng);
 * dsocr_test_stdrng_u64: first n next_u64() draws of StdRng::seed_from_u64(seed);
 * dsocr_test_select_tokens: n_steps calls of select_token_id on logits[step] with one RNG made from params
 * (seeded or from entropy) and a context that grows by the selected token after every call, as generate does. */
DSOCR_API int dsocr_test_chacha_words(const uint8_t* key32, int rounds, int n, uint32_t* out);
DSOCR_API int dsocr_test_stdrng_u64(uint64_t seed, int n, uint64_t* out);
DSOCR_API int dsocr_test_select_tokens(const float* logits, size_t vocab, int n_steps, const dsocr_decode_params* params,
                                       const int64_t* context, size_t n_context, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif
