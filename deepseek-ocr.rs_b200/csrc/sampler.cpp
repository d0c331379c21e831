// Host-side token selection for the sampling path; see sampler.h.
#include "sampler.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <unordered_set>

namespace dsocr {

namespace {
inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
inline void quarter(uint32_t* s, int a, int b, int c, int d) {
  s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
  s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
  s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
  s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
}
// one 64-byte ChaCha block: constants | key | 64-bit block counter | 64-bit stream id (0)
void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
  uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  for (int i = 0; i < 8; ++i) in[4 + i] = key[i];
  in[12] = (uint32_t)counter; in[13] = (uint32_t)(counter >> 32); in[14] = 0; in[15] = 0;
  uint32_t s[16];
  memcpy(s, in, sizeof(s));
  for (int r = 0; r < rounds; r += 2) {
    quarter(s, 0, 4, 8, 12); quarter(s, 1, 5, 9, 13); quarter(s, 2, 6, 10, 14); quarter(s, 3, 7, 11, 15);
    quarter(s, 0, 5, 10, 15); quarter(s, 1, 6, 11, 12); quarter(s, 2, 7, 8, 13); quarter(s, 3, 4, 9, 14);
  }
  for (int i = 0; i < 16; ++i) out[i] = s[i] + in[i];
}
}  // namespace

StdRng::StdRng(uint64_t state) {
  // rand_core::SeedableRng::seed_from_u64: PCG32 (XSH-RR) output words, little endian, fill the 32-byte seed
  const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
  for (int i = 0; i < 8; ++i) {
    state = state * MUL + INC;
    const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    const uint32_t rot = (uint32_t)(state >> 59);
    key_[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
}

StdRng::StdRng(const uint8_t key[32], int rounds) : rounds_(rounds) {
  for (int i = 0; i < 8; ++i)
    key_[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
}

StdRng StdRng::from_entropy() {
  uint8_t key[32];
  FILE* f = fopen("/dev/urandom", "rb");
  if (!f || fread(key, 1, 32, f) != 32) {
    if (f) fclose(f);
    throw std::runtime_error("cannot read the OS entropy source for an unseeded sampling call");
  }
  fclose(f);
  return StdRng(key, 12);
}

void StdRng::refill() {
  for (int b = 0; b < 4; ++b) chacha_block(key_, counter_ + b, rounds_, buf_ + 16 * b);
  counter_ += 4;
}

uint32_t StdRng::next_u32() {
  if (index_ >= 64) { refill(); index_ = 0; }
  return buf_[index_++];
}

uint64_t StdRng::next_u64() {  // rand_core::block::BlockRng::next_u64
  if (index_ < 63) {
    const uint64_t v = ((uint64_t)buf_[index_ + 1] << 32) | buf_[index_];
    index_ += 2;
    return v;
  }
  if (index_ >= 64) {
    refill(); index_ = 2;
    return ((uint64_t)buf_[1] << 32) | buf_[0];
  }
  const uint64_t x = buf_[63];
  refill(); index_ = 1;
  return ((uint64_t)buf_[0] << 32) | x;
}

namespace {
inline bool finite32(float v) { return std::isfinite(v); }

// argmax_index (sampling.rs:104-118): first index among equal maxima, non-finite skipped
long long argmax_index(const std::vector<float>& v) {
  long long best = -1;
  float cur = 0.f;
  for (size_t i = 0; i < v.size(); ++i) {
    if (!finite32(v[i])) continue;
    if (best < 0 || v[i] > cur) { best = (long long)i; cur = v[i]; }
  }
  return best;
}

// apply_top_k (sampling.rs:160-174)
void apply_top_k(std::vector<double>& l, size_t k) {
  if (k == 0 || l.empty()) return;
  std::vector<size_t> idx;
  for (size_t i = 0; i < l.size(); ++i) if (std::isfinite(l[i])) idx.push_back(i);
  if (idx.size() <= k) return;
  std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return l[a] > l[b]; });
  for (size_t i = k; i < idx.size(); ++i) l[idx[i]] = -std::numeric_limits<double>::infinity();
}

// apply_top_p (sampling.rs:176-224)
void apply_top_p(std::vector<double>& l, double top_p) {
  if (!(top_p >= 0.0 && top_p < 1.0) || l.empty()) return;
  std::vector<std::pair<size_t, double>> pairs;
  for (size_t i = 0; i < l.size(); ++i) if (std::isfinite(l[i])) pairs.emplace_back(i, l[i]);
  if (pairs.empty()) return;
  std::stable_sort(pairs.begin(), pairs.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
  const double mx = pairs[0].second;
  std::vector<double> w(pairs.size());
  double total = 0.0;
  for (size_t i = 0; i < pairs.size(); ++i) { w[i] = std::exp(pairs[i].second - mx); total += w[i]; }
  if (total <= 0.0) return;
  double cum = 0.0;
  size_t keep = pairs.size();
  for (size_t i = 0; i < w.size(); ++i) {
    cum += w[i] / total;
    if (cum > top_p) { keep = i + 1; break; }
  }
  if (keep == 0) keep = 1;
  std::vector<char> mask(l.size(), 0);
  for (size_t i = 0; i < keep; ++i) mask[pairs[i].first] = 1;
  for (size_t i = 0; i < l.size(); ++i) if (!mask[i]) l[i] = -std::numeric_limits<double>::infinity();
}

// sample_from_logits (sampling.rs:226-259) with rand 0.8.5's WeightedIndex<f64> / UniformFloat<f64>
long long sample_from_logits(const std::vector<double>& l, StdRng& rng) {
  std::vector<size_t> idx;
  for (size_t i = 0; i < l.size(); ++i) if (std::isfinite(l[i])) idx.push_back(i);
  if (idx.empty()) return -1;
  double mx = -std::numeric_limits<double>::infinity();
  for (size_t i : idx) mx = std::max(mx, l[i]);
  if (!std::isfinite(mx)) return -1;
  std::vector<double> w(idx.size());
  bool any = false;
  for (size_t j = 0; j < idx.size(); ++j) {
    const double e = std::exp(l[idx[j]] - mx);
    w[j] = (std::isfinite(e) && e > 0.0) ? e : 0.0;
    any |= w[j] > 0.0;
  }
  if (!any) {  // Iterator::max_by keeps the last of several equal maxima
    size_t best = idx[0];
    for (size_t j = 1; j < idx.size(); ++j) if (!(l[idx[j]] < l[best])) best = idx[j];
    return (long long)best;
  }
  // WeightedIndex::new: cumulative weights of all but the last item, total, Uniform::new(0, total)
  std::vector<double> cum;
  cum.reserve(w.size());
  double total = w[0];
  for (size_t j = 1; j < w.size(); ++j) { cum.push_back(total); total += w[j]; }
  if (total == 0.0) return -1;
  const double low = 0.0, high = total;
  const double max_rand = 1.0 - std::ldexp(1.0, -52);  // (u64::MAX >> 12) as a [1, 2) mantissa, minus 1
  double scale = high - low;
  while (scale * max_rand + low >= high) {  // UniformFloat::new: shrink until the largest draw stays below `high`
    uint64_t bits;
    memcpy(&bits, &scale, 8);
    --bits;
    memcpy(&scale, &bits, 8);
  }
  // UniformFloat::sample: 52 random mantissa bits -> [1, 2) -> [0, 1) -> scale
  const uint64_t bits = (rng.next_u64() >> 12) | (1023ull << 52);
  double v12;
  memcpy(&v12, &bits, 8);
  const double chosen = (v12 - 1.0) * scale + low;
  // first item whose cumulative weight is above the draw (binary_search_by ... unwrap_err)
  const size_t pos = (size_t)(std::partition_point(cum.begin(), cum.end(), [&](double c) { return c <= chosen; }) - cum.begin());
  return (long long)idx[pos];
}
}  // namespace

int64_t select_token_id(const float* logits, size_t V, const SamplingParams& p, const int64_t* context, size_t n_context,
                        StdRng& rng) {
  if (V == 0) throw std::runtime_error("logits tensor is empty");
  std::vector<float> raw(logits, logits + V), adjusted(raw);
  // apply_repetition_penalty (sampling.rs:120-139)
  if (!(p.repetition_penalty <= 0.f) && std::fabs(p.repetition_penalty - 1.0f) > std::numeric_limits<float>::epsilon()) {
    const float pen = std::max(p.repetition_penalty, std::numeric_limits<float>::min());
    std::unordered_set<size_t> seen;
    for (size_t i = 0; i < n_context; ++i) {
      if (context[i] < 0 || (size_t)context[i] >= V) continue;
      if (!seen.insert((size_t)context[i]).second) continue;
      float& e = adjusted[(size_t)context[i]];
      if (e > 0.f) e /= pen; else e *= pen;
    }
  }
  std::vector<float> filtered(adjusted);
  // banned_ngram_tokens (sampling.rs:141-158)
  if (p.no_repeat_ngram > 1 && n_context + 1 >= p.no_repeat_ngram) {
    const size_t n = p.no_repeat_ngram, pre = n - 1;
    const int64_t* tail = context + n_context - pre;
    for (size_t i = 0; i + n <= n_context; ++i) {
      if (memcmp(context + i, tail, pre * sizeof(int64_t)) != 0) continue;
      const int64_t t = context[i + pre];
      if (t >= 0 && (size_t)t < V) filtered[(size_t)t] = -std::numeric_limits<float>::infinity();
    }
  }
  bool valid = false;
  for (float v : filtered) if (finite32(v)) { valid = true; break; }
  if (!valid) filtered = adjusted;

  if (p.do_sample && p.temperature > 0.0) {
    std::vector<double> l64(V);
    for (size_t i = 0; i < V; ++i) l64[i] = (double)filtered[i] / p.temperature;
    if (p.has_top_k && p.top_k > 0 && p.top_k < V) apply_top_k(l64, p.top_k);
    if (p.has_top_p && p.top_p >= 0.0 && p.top_p < 1.0) apply_top_p(l64, p.top_p);
    const long long s = sample_from_logits(l64, rng);
    if (s >= 0) return s;
  }
  long long best = argmax_index(filtered);
  if (best < 0) best = argmax_index(adjusted);
  if (best < 0) best = argmax_index(raw);
  return best < 0 ? 0 : best;
}

}  // namespace dsocr
