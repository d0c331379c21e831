// DSQ snapshot writer (host only): the on-disk producer of the format the engine reads.  Mirrors DsqWriter of
// crates/dsq-writer/src/lib.rs: tensors are appended to a payload, `finalize` emits header | records | payload with the
// record offsets rebased by the metadata length (:410-527), the output path gets the `.dsq` extension (:104), duplicate
// names / unaligned in_dim / wrong byte lengths are rejected with the reference's wording.
//   Q8_0 from f32  : quantize_q8_0 (:555-598) - d = amax / 127 in f32 (stored as f16), q = round-half-away(v * (1 / d)),
//                    clamp [-128, 127], all-zero block -> d = 0, q = 0.  Bit-exact with gguf-py (tests/golden/dsq_blocks.npz).
//   F32 / F16 / BF16: raw little-endian payloads (:281-364); f16 / bf16 are rounded to nearest-even from f32 here.
//   Q4_K / Q6_K    : `add_quantized_bytes` takes blocks quantised elsewhere (:366-409).  From f32 the reference calls
//                    candle's BlockQ4K / BlockQ6K::from_float (:600-664); candle is not available offline, so the
//                    quantisers below follow ggml's reference routines (make_qkx1_quants / make_qx_quants search, 6-bit
//                    sub-scales) as published and are NOT byte-pinned to candle - they emit valid blocks whose
//                    dequantisation error is tested, nothing more is claimed.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "dsq.h"

namespace dsocr {

namespace {

uint16_t f16_bits(float v) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
float f16_val(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
uint16_t bf16_bits(float v) { __nv_bfloat16 h = __float2bfloat16_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
int nearest_int(float v) { return (int)lrintf(v); }  // ggml nearest_int: round half to even

void quantize_q8_0(const float* w, size_t rows, size_t cols, uint8_t* dst) {
  for (size_t b = 0; b < rows * cols / 32; ++b) {
    const float* x = w + b * 32;
    uint8_t* o = dst + b * 34;
    float amax = 0.f;
    for (int i = 0; i < 32; ++i) amax = std::max(amax, fabsf(x[i]));
    const float scale = amax > 0.f ? amax / 127.0f : 0.f;
    const uint16_t d = f16_bits(scale);
    o[0] = (uint8_t)(d & 0xFF); o[1] = (uint8_t)(d >> 8);
    if (scale == 0.f) { memset(o + 2, 0, 32); continue; }
    const float inv = 1.0f / scale;
    for (int i = 0; i < 32; ++i) {
      float q = roundf(x[i] * inv);  // f32::round: half away from zero
      q = std::min(127.f, std::max(-128.f, q));
      o[2 + i] = (uint8_t)(int8_t)(int)q;
    }
  }
}

// ggml make_qkx1_quants: asymmetric (scale, min) for n values on levels 0..nmax, `ntry` refinement rounds.
float make_qkx1_quants(int n, int nmax, const float* x, uint8_t* L, float* the_min, int ntry) {
  float mn = x[0], mx = x[0];
  for (int i = 1; i < n; ++i) { mn = std::min(mn, x[i]); mx = std::max(mx, x[i]); }
  if (mn > 0.f) mn = 0.f;  // clamp first (as ggml's make_qkx2_quants does): a constant positive block keeps a scale,
  if (mx == mn) { for (int i = 0; i < n; ++i) L[i] = 0; *the_min = -mn; return 0.f; }  // a constant negative one lives in the min
  float iscale = (float)nmax / (mx - mn);
  float scale = 1.f / iscale;
  for (int t = 0; t < ntry; ++t) {
    float sumlx = 0.f; int suml2 = 0; bool changed = false;
    for (int i = 0; i < n; ++i) {
      int l = std::max(0, std::min(nmax, nearest_int(iscale * (x[i] - mn))));
      if (l != L[i]) { L[i] = (uint8_t)l; changed = true; }
      sumlx += (x[i] - mn) * l; suml2 += l * l;
    }
    if (suml2 > 0) scale = sumlx / suml2;
    float sum = 0.f;
    for (int i = 0; i < n; ++i) sum += x[i] - scale * L[i];
    mn = sum / n;
    if (mn > 0.f) mn = 0.f;
    iscale = scale != 0.f ? 1.f / scale : 0.f;
    if (!changed) break;
  }
  *the_min = -mn;
  return scale;
}

// ggml make_qx_quants (rmse_type 1): symmetric scale for n values on levels -nmax..nmax-1, 19-point search around
// nmax / max|x| weighted by x^2.
float make_qx_quants(int n, int nmax, const float* x, int8_t* L) {
  float mx = 0.f, amax = 0.f;
  for (int i = 0; i < n; ++i) { const float ax = fabsf(x[i]); if (ax > amax) { amax = ax; mx = x[i]; } }
  if (amax == 0.f) { for (int i = 0; i < n; ++i) L[i] = 0; return 0.f; }
  float iscale = -(float)nmax / mx;
  float sumlx = 0.f, suml2 = 0.f;
  for (int i = 0; i < n; ++i) {
    int l = std::max(-nmax, std::min(nmax - 1, nearest_int(iscale * x[i])));
    L[i] = (int8_t)(l + nmax);
    const float w = x[i] * x[i];
    sumlx += w * x[i] * l; suml2 += w * l * l;
  }
  float scale = suml2 > 0.f ? sumlx / suml2 : 0.f;
  float best = scale * sumlx;
  for (int is = -9; is <= 9; ++is) {
    if (is == 0) continue;
    iscale = -((float)nmax + 0.1f * is) / mx;
    sumlx = suml2 = 0.f;
    for (int i = 0; i < n; ++i) {
      int l = std::max(-nmax, std::min(nmax - 1, nearest_int(iscale * x[i])));
      const float w = x[i] * x[i];
      sumlx += w * x[i] * l; suml2 += w * l * l;
    }
    if (suml2 > 0.f && sumlx * sumlx > best * suml2) {
      for (int i = 0; i < n; ++i) L[i] = (int8_t)(nmax + std::max(-nmax, std::min(nmax - 1, nearest_int(iscale * x[i]))));
      scale = sumlx / suml2; best = scale * sumlx;
    }
  }
  return scale;
}

void scale_min_k4(int j, const uint8_t* q, uint8_t* d, uint8_t* m) {  // ggml get_scale_min_k4
  if (j < 4) { *d = q[j] & 63; *m = q[j + 4] & 63; }
  else { *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); *m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

void quantize_q4k(const float* w, size_t rows, size_t cols, uint8_t* dst) {
  for (size_t b = 0; b < rows * cols / 256; ++b) {
    const float* x = w + b * 256;
    uint8_t L[256]; float mins[8], scales[8];
    float max_scale = 0.f, max_min = 0.f;
    for (int j = 0; j < 8; ++j) {
      for (int i = 0; i < 32; ++i) L[32 * j + i] = 0;
      scales[j] = make_qkx1_quants(32, 15, x + 32 * j, L + 32 * j, &mins[j], 5);
      max_scale = std::max(max_scale, scales[j]); max_min = std::max(max_min, mins[j]);
    }
    const float inv_scale = max_scale > 0.f ? 63.f / max_scale : 0.f;
    const float inv_min = max_min > 0.f ? 63.f / max_min : 0.f;
    uint8_t blk[144] = {0};
    uint8_t* sc = blk + 4;
    for (int j = 0; j < 8; ++j) {
      const uint8_t ls = (uint8_t)std::min(63, nearest_int(inv_scale * scales[j]));
      const uint8_t lm = (uint8_t)std::min(63, nearest_int(inv_min * mins[j]));
      if (j < 4) { sc[j] = ls; sc[j + 4] = lm; }
      else { sc[j + 4] = (ls & 0xF) | ((lm & 0xF) << 4); sc[j - 4] |= (uint8_t)((ls >> 4) << 6); sc[j] |= (uint8_t)((lm >> 4) << 6); }
    }
    const uint16_t d16 = f16_bits(max_scale / 63.f), m16 = f16_bits(max_min / 63.f);
    memcpy(blk, &d16, 2); memcpy(blk + 2, &m16, 2);
    const float dd = f16_val(d16), dmin = f16_val(m16);
    for (int j = 0; j < 8; ++j) {
      uint8_t s, m; scale_min_k4(j, sc, &s, &m);
      const float d = dd * s;
      if (d == 0.f) continue;
      const float dm = dmin * m;
      for (int i = 0; i < 32; ++i) L[32 * j + i] = (uint8_t)std::max(0, std::min(15, nearest_int((x[32 * j + i] + dm) / d)));
    }
    uint8_t* q = blk + 16;
    for (int g = 0; g < 4; ++g)
      for (int l = 0; l < 32; ++l) q[32 * g + l] = L[64 * g + l] | (uint8_t)(L[64 * g + 32 + l] << 4);
    memcpy(dst + b * 144, blk, 144);
  }
}

void quantize_q6k(const float* w, size_t rows, size_t cols, uint8_t* dst) {
  for (size_t b = 0; b < rows * cols / 256; ++b) {
    const float* x = w + b * 256;
    int8_t L[256]; float scales[16];
    float max_scale = 0.f, max_abs = 0.f;
    for (int ib = 0; ib < 16; ++ib) {
      scales[ib] = make_qx_quants(16, 32, x + 16 * ib, L + 16 * ib);
      if (fabsf(scales[ib]) > max_abs) { max_abs = fabsf(scales[ib]); max_scale = scales[ib]; }
    }
    uint8_t blk[210] = {0};
    if (max_abs == 0.f) { memcpy(dst + b * 210, blk, 210); continue; }  // all-zero block
    const float iscale = -128.f / max_scale;
    const uint16_t d16 = f16_bits(1.f / iscale);
    memcpy(blk + 208, &d16, 2);
    int8_t* sc = reinterpret_cast<int8_t*>(blk + 192);
    for (int ib = 0; ib < 16; ++ib) sc[ib] = (int8_t)std::min(127, nearest_int(iscale * scales[ib]));
    const float dd = f16_val(d16);
    for (int j = 0; j < 16; ++j) {
      const float d = dd * sc[j];
      if (d == 0.f) { for (int i = 0; i < 16; ++i) L[16 * j + i] = 32; continue; }
      for (int i = 0; i < 16; ++i) L[16 * j + i] = (int8_t)(std::max(-32, std::min(31, nearest_int(x[16 * j + i] / d))) + 32);
    }
    uint8_t* ql = blk; uint8_t* qh = blk + 128;
    for (int half = 0; half < 2; ++half)
      for (int l = 0; l < 32; ++l) {
        const uint8_t q1 = (uint8_t)L[128 * half + l], q2 = (uint8_t)L[128 * half + 32 + l], q3 = (uint8_t)L[128 * half + 64 + l],
                      q4 = (uint8_t)L[128 * half + 96 + l];
        ql[64 * half + l] = (q1 & 0xF) | (uint8_t)((q3 & 0xF) << 4);
        ql[64 * half + 32 + l] = (q2 & 0xF) | (uint8_t)((q4 & 0xF) << 4);
        qh[32 * half + l] = (uint8_t)((q1 >> 4) | ((q2 >> 4) << 2) | ((q3 >> 4) << 4) | ((q4 >> 4) << 6));
      }
    memcpy(dst + b * 210, blk, 210);
  }
}

void put_u32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
void put_u64(std::vector<uint8_t>& b, uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
void put_str(std::vector<uint8_t>& b, const std::string& s) { put_u32(b, (uint32_t)s.size()); b.insert(b.end(), s.begin(), s.end()); }

}  // namespace

DsqWriter::DsqWriter(const std::string& path, const std::string& candle_version, const std::string& model_id,
                     const std::string& backend, DsqDType default_dtype)
    : candle_version_(candle_version), model_id_(model_id), backend_(backend), default_dtype_(default_dtype) {
  if (dsq_block_elems(default_dtype) == 0) throw std::runtime_error("value overflow while encoding block_size");  // SnapshotMetadata::block_size
  // Path::with_extension("dsq")
  const size_t slash = path.find_last_of('/');
  const size_t dot = path.find_last_of('.');
  path_ = (dot != std::string::npos && (slash == std::string::npos || dot > slash + 1)) ? path.substr(0, dot) + ".dsq" : path + ".dsq";
}

void DsqWriter::append(const std::string& name, uint32_t out_dim, uint32_t in_dim, DsqDType dt, const uint8_t* q, size_t q_len,
                       const float* bias) {
  Pending r;
  r.name = name; r.out_dim = out_dim; r.in_dim = in_dim; r.dtype = dt;
  r.q_offset = payload_.size(); r.q_len = q_len;
  payload_.insert(payload_.end(), q, q + q_len);
  if (bias) {
    r.bias_offset = payload_.size(); r.bias_len = (uint64_t)out_dim * 4; r.has_bias = true;
    const uint8_t* bb = reinterpret_cast<const uint8_t*>(bias);  // little-endian host
    payload_.insert(payload_.end(), bb, bb + r.bias_len);
  }
  records_.push_back(std::move(r));
}

void DsqWriter::check_new(const std::string& name, DsqDType dt, uint32_t in_dim) const {
  for (const Pending& r : records_)
    if (r.name == name) throw std::runtime_error("tensor `" + name + "` already exists in snapshot");
  const int be = dsq_block_elems(dt);
  if (be && in_dim % be)
    throw std::runtime_error("tensor `" + name + "` in_dim " + std::to_string(in_dim) + " is not divisible by block size " + std::to_string(be));
}

void DsqWriter::add_tensor_f32(const std::string& name, uint32_t out_dim, uint32_t in_dim, DsqDType dt, const float* w, const float* bias) {
  check_new(name, dt, in_dim);
  const size_t n = (size_t)out_dim * in_dim;
  std::vector<uint8_t> q;
  const int be = dsq_block_elems(dt);
  if (be) {
    // rows are independent: quantise row ranges on the host cores (the output offset of a row is known up front, so the
    // bytes do not depend on the thread count)
    const size_t row_bytes = (size_t)(in_dim / be) * dsq_block_bytes(dt);
    q.resize((size_t)out_dim * row_bytes);
    const size_t want = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
    const size_t nthreads = n < (1u << 18) ? 1 : std::min<size_t>(want, out_dim);
    auto work = [&](size_t r0, size_t r1) {
      const float* src = w + r0 * in_dim;
      uint8_t* dst = q.data() + r0 * row_bytes;
      if (dt == DsqDType::Q8_0) quantize_q8_0(src, r1 - r0, in_dim, dst);
      else if (dt == DsqDType::Q4K) quantize_q4k(src, r1 - r0, in_dim, dst);
      else quantize_q6k(src, r1 - r0, in_dim, dst);
    };
    if (nthreads <= 1) work(0, out_dim);
    else {
      std::vector<std::thread> th;
      const size_t chunk = (out_dim + nthreads - 1) / nthreads;
      for (size_t t = 0; t < nthreads; ++t) {
        const size_t r0 = t * chunk, r1 = std::min<size_t>(out_dim, r0 + chunk);
        if (r0 < r1) th.emplace_back(work, r0, r1);
      }
      for (auto& t : th) t.join();
    }
  } else if (dt == DsqDType::F32) {
    q.resize(n * 4); memcpy(q.data(), w, n * 4);
  } else if (dt == DsqDType::F16) {
    q.resize(n * 2);
    for (size_t i = 0; i < n; ++i) { const uint16_t u = f16_bits(w[i]); memcpy(&q[2 * i], &u, 2); }
  } else {
    q.resize(n * 2);
    for (size_t i = 0; i < n; ++i) { const uint16_t u = bf16_bits(w[i]); memcpy(&q[2 * i], &u, 2); }
  }
  append(name, out_dim, in_dim, dt, q.data(), q.size(), bias);
}

void DsqWriter::add_quantized_bytes(const std::string& name, uint32_t out_dim, uint32_t in_dim, DsqDType dt, const uint8_t* q,
                                    size_t q_len, const float* bias) {
  if (dsq_block_elems(dt) == 0) throw std::runtime_error("quantization failed: add_quantized_bytes expects quantized dtype");
  check_new(name, dt, in_dim);
  const size_t expected = (size_t)out_dim * (in_dim / dsq_block_elems(dt)) * dsq_block_bytes(dt);
  if (q_len != expected)
    throw std::runtime_error("tensor `" + name + "` expected " + std::to_string(expected) + " elements but received " + std::to_string(q_len));
  append(name, out_dim, in_dim, dt, q, q_len, bias);
}

void DsqWriter::finalize() {
  std::vector<uint8_t> head;
  const char magic[7] = {'D', 'S', 'Q', 'S', 'N', 'A', 'P'};
  head.insert(head.end(), magic, magic + 7);
  put_u32(head, 1);
  put_str(head, candle_version_); put_str(head, model_id_); put_str(head, backend_);
  put_u32(head, (uint32_t)default_dtype_);
  put_u32(head, (uint32_t)dsq_block_elems(default_dtype_));
  put_u32(head, (uint32_t)records_.size());
  uint64_t meta = head.size();
  for (const Pending& r : records_) meta += 52 + r.name.size();  // record_entry_len
  for (const Pending& r : records_) {
    put_str(head, r.name);
    put_u32(head, r.out_dim); put_u32(head, r.in_dim); put_u32(head, (uint32_t)r.dtype);
    put_u64(head, r.q_offset + meta); put_u64(head, r.q_len);
    if (r.has_bias) { put_u64(head, r.bias_offset + meta); put_u64(head, r.bias_len); put_u32(head, 4 /* DsqBiasDType::F32 */); }
    else { put_u64(head, 0); put_u64(head, 0); put_u32(head, 0); }
  }
  std::ofstream f(path_, std::ios::binary | std::ios::trunc);
  if (!f) throw std::runtime_error("cannot create snapshot " + path_);
  f.write(reinterpret_cast<const char*>(head.data()), (std::streamsize)head.size());
  f.write(reinterpret_cast<const char*>(payload_.data()), (std::streamsize)payload_.size());
  if (!f) throw std::runtime_error("failed to write snapshot " + path_);
}

}  // namespace dsocr
