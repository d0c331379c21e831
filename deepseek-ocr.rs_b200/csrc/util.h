// Small host-side helpers: device buffers, error plumbing, 16-bit conversion.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

#include "kernels.h"

namespace dsocr {

void set_last_error(const std::string& msg);

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return -5;
  } catch (...) {
    set_last_error("unknown error");
    return -5;
  }
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    cuda_check(cudaMalloc(&p, n), "cudaMalloc");
    bytes = n;
  }
  void ensure(size_t n) { if (n > bytes) alloc(n); }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline uint16_t f32_to_16(float v, DType dt) {
  if (dt == DType::BF16) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    uint16_t u; memcpy(&u, &b, 2); return u;
  }
  __half h = __float2half_rn(v);
  uint16_t u; memcpy(&u, &h, 2); return u;
}
inline float f16_to_32(uint16_t u, DType dt) {
  if (dt == DType::BF16) { __nv_bfloat16 b; memcpy(&b, &u, 2); return __bfloat162float(b); }
  __half h; memcpy(&h, &u, 2); return __half2float(h);
}

inline std::vector<uint16_t> to16(const float* src, size_t n, DType dt) {
  std::vector<uint16_t> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = f32_to_16(src[i], dt);
  return out;
}
// hi/lo split: hi = r16(x), lo = r16(x - hi); layout [hi(n) | lo(n)]
inline std::vector<uint16_t> to16_split(const float* src, size_t n, DType dt) {
  std::vector<uint16_t> out(2 * n);
  for (size_t i = 0; i < n; ++i) {
    uint16_t hi = f32_to_16(src[i], dt);
    out[i] = hi;
    out[n + i] = f32_to_16(src[i] - f16_to_32(hi, dt), dt);
  }
  return out;
}

// cudaMemcpy from pageable memory may return before the DMA has landed and only orders against the legacy
// default stream; the engine's kernels run on their own stream, so make the copy globally visible first.
inline void h2d(void* dst, const void* src, size_t bytes) {
  cuda_check(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy H2D");
  cuda_check(cudaDeviceSynchronize(), "cudaMemcpy H2D sync");
}
inline void d2h(void* dst, const void* src, size_t bytes) {
  cuda_check(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy D2H");
}

}  // namespace dsocr
