// Host-side token selection for the sampling path (do_sample with temperature > 0): restates
// crates/core/src/sampling.rs:34-96 and the pieces of rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6 it draws from
// (Cargo.lock:3178-3205), so that a seeded call produces the reference's draw sequence.  Greedy selection never comes
// here: it runs on the device (select_token_kernel).
#pragma once
#include <cstdint>
#include <vector>

#include "dsocr.h"

namespace dsocr {

// rand::rngs::StdRng of rand 0.8 = ChaCha12Rng: 256-bit key = seed, 64-bit block counter, stream 0; the block RNG
// buffers four 64-byte blocks and next_u64 reads two consecutive words, low word first (rand_core::block::BlockRng).
class StdRng {
 public:
  explicit StdRng(uint64_t seed);            // SeedableRng::seed_from_u64: PCG32 expansion of the seed into 32 bytes
  explicit StdRng(const uint8_t key[32], int rounds = 12);  // from_seed (rounds exposed for the known-answer tests)
  static StdRng from_entropy();              // 32 bytes from the OS entropy source
  uint32_t next_u32();
  uint64_t next_u64();

 private:
  void refill();
  uint32_t key_[8];
  uint64_t counter_ = 0;
  int rounds_ = 12;
  uint32_t buf_[64];
  int index_ = 64;
};

struct SamplingParams {  // TokenSelectionParams (sampling.rs:16-23)
  bool do_sample = false;
  double temperature = 0.0;
  bool has_top_p = false; double top_p = 1.0;
  bool has_top_k = false; size_t top_k = 0;
  float repetition_penalty = 1.0f;
  size_t no_repeat_ngram = 0;  // 0 = None
};

// DecodeParameters as they cross the C ABI -> TokenSelectionParams
inline SamplingParams sampling_params_of(const dsocr_decode_params& d) {
  SamplingParams p;
  p.do_sample = d.do_sample != 0;
  p.temperature = d.temperature;
  p.has_top_p = d.top_p >= 0.0; p.top_p = d.top_p;
  p.has_top_k = d.top_k > 0; p.top_k = d.top_k;
  p.repetition_penalty = d.repetition_penalty;
  p.no_repeat_ngram = d.no_repeat_ngram_size;
  return p;
}

// select_token_id (sampling.rs:34-96) on one logits row.
int64_t select_token_id(const float* logits, size_t V, const SamplingParams& p, const int64_t* context, size_t n_context,
                        StdRng& rng);

}  // namespace dsocr
