// Bandwidth-bound kernels of the vision path: image normalise + patch gather, LayerNorm (with the SAM
// window partition folded into its row mapping), im2col for the 3x3 neck convs, CLIP embedding assembly,
// CLIP|SAM concat and the final token-layout scatter.  All use 128-bit accesses and warp-shuffle reductions.
#include "kernels.h"
#include "ptx.cuh"

namespace dsocr {

namespace {

template <typename T> struct Pack8 { uint4 v; };

template <typename T>
__device__ __forceinline__ uint32_t pack2f(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2f<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2f<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T>
__device__ __forceinline__ uint4 pack8f(const float* f) {
  uint4 r;
  r.x = pack2f<T>(f[0], f[1]); r.y = pack2f<T>(f[2], f[3]); r.z = pack2f<T>(f[4], f[5]); r.w = pack2f<T>(f[6], f[7]);
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------
// image_to_tensor (model/mod.rs:2332-2347) fused with the 16x16/s16 patch gather of the SAM patch-embed
// conv (vision/sam.rs:427-456): u8 HWC -> 16-bit [B*g*g, 3*16*16], column = c*256 + ky*16 + kx.
// One thread = one (token, ky): 48 contiguous input bytes (3 x 128-bit loads), 3 x 32 output bytes.
template <typename T>
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, T* __restrict__ out, int B, int G) {
  const int g = G / 16;
  const long long total = (long long)B * g * g * 16;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ky = idx & 15;
  const long long tok = idx >> 4;
  const int px = tok % g;
  const int py = (tok / g) % g;
  const int b = tok / ((long long)g * g);
  const uint8_t* src = img + (((long long)b * G + (py * 16 + ky)) * G + px * 16) * 3;
  uint4 raw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) raw[i] = reinterpret_cast<const uint4*>(src)[i];
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(raw);
  T* dst = out + tok * 768 + ky * 16;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f[16];
#pragma unroll
    for (int kx = 0; kx < 16; ++kx) f[kx] = ((float)bytes[kx * 3 + c] / 255.0f - 0.5f) / 0.5f;
    reinterpret_cast<uint4*>(dst + c * 256)[0] = pack8f<T>(f);
    reinterpret_cast<uint4*>(dst + c * 256)[1] = pack8f<T>(f + 8);
  }
}

// Same gather from the reference's f32 CHW tensor ([B,3,G,G], already normalised).
template <typename T>
__global__ void patchify_f32_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int G) {
  const int g = G / 16;
  const long long total = (long long)B * g * g * 48;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ky = idx % 16;
  const int c = (idx / 16) % 3;
  const long long tok = idx / 48;
  const int px = tok % g;
  const int py = (tok / g) % g;
  const int b = tok / ((long long)g * g);
  const float* src = img + (((long long)b * 3 + c) * G + (py * 16 + ky)) * G + px * 16;
  float f[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(f)[i] = reinterpret_cast<const float4*>(src)[i];
  T* dst = out + tok * 768 + c * 256 + ky * 16;
  reinterpret_cast<uint4*>(dst)[0] = pack8f<T>(f);
  reinterpret_cast<uint4*>(dst)[1] = pack8f<T>(f + 8);
}

// dst[b*T + t, :] = src[t, :]   (absolute position embedding broadcast over the batch)
__global__ void bcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows_per_batch,
                                  int cols4, long long total4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const long long per = rows_per_batch * cols4;
  reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i % per];
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (candle layer_norm: biased variance, eps inside the sqrt), one warp per
// OUTPUT row.  `win` > 0 folds SAM's window_partition (vision/sam.rs:926-955) into the row mapping: output
// row (b, wy, wx, iy, ix) reads token (b, wy*win+iy, wx*win+ix) and is ZERO when that falls in the padding
// (the reference pads with zeros *after* norm1, sam.rs:733-739).
template <typename T, int C>
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bi,
                                 T* __restrict__ out16, float* __restrict__ out32, long long out_rows, float eps,
                                 int win, int g, int nw) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= out_rows) return;
  const int lane = threadIdx.x & 31;
  long long src = row;
  if (win > 0) {
    const int per = win * win;
    const int i = row % per;
    const long long wi = row / per;
    const int wx = wi % nw, wy = (wi / nw) % nw;
    const long long b = wi / ((long long)nw * nw);
    const int ty = wy * win + i / win, tx = wx * win + i % win;
    src = (ty < g && tx < g) ? (b * g + ty) * g + tx : -1;
  }
  constexpr int G8 = C / 256;  // lane owns 8 consecutive channels per 256-channel group
  if (src < 0) {
#pragma unroll
    for (int gq = 0; gq < G8; ++gq) {
      const int c = gq * 256 + lane * 8;
      if (out16) reinterpret_cast<uint4*>(out16 + row * C + c)[0] = make_uint4(0, 0, 0, 0);
      if (out32) {
        reinterpret_cast<float4*>(out32 + row * C + c)[0] = make_float4(0, 0, 0, 0);
        reinterpret_cast<float4*>(out32 + row * C + c)[1] = make_float4(0, 0, 0, 0);
      }
    }
    return;
  }
  float f[G8][8];
  float s = 0.f;
#pragma unroll
  for (int gq = 0; gq < G8; ++gq) {
    const float4* p = reinterpret_cast<const float4*>(x + src * C + gq * 256 + lane * 8);
    float4 a = p[0], b4 = p[1];
    f[gq][0] = a.x; f[gq][1] = a.y; f[gq][2] = a.z; f[gq][3] = a.w;
    f[gq][4] = b4.x; f[gq][5] = b4.y; f[gq][6] = b4.z; f[gq][7] = b4.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[gq][i];
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float vs = 0.f;
#pragma unroll
  for (int gq = 0; gq < G8; ++gq)
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = f[gq][i] - mean; vs += d * d; }
  const float rstd = rsqrtf(warp_sum(vs) * (1.0f / C) + eps);
#pragma unroll
  for (int gq = 0; gq < G8; ++gq) {
    const int c = gq * 256 + lane * 8;
    float o[8];
    const float4 w0 = reinterpret_cast<const float4*>(w + c)[0], w1 = reinterpret_cast<const float4*>(w + c)[1];
    const float4 b0 = reinterpret_cast<const float4*>(bi + c)[0], b1 = reinterpret_cast<const float4*>(bi + c)[1];
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (f[gq][i] - mean) * rstd * ww[i] + bb[i];
    if (out16) reinterpret_cast<uint4*>(out16 + row * C + c)[0] = pack8f<T>(o);
    if (out32) {
      reinterpret_cast<float4*>(out32 + row * C + c)[0] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(out32 + row * C + c)[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// window row -> token row (or -1 for padding); consumed by the proj GEMM epilogue (window_unpartition,
// vision/sam.rs:957-980, becomes an output row remap + residual add).
__global__ void window_row_map_kernel(int* __restrict__ map, long long rows, int win, int g, int nw) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int per = win * win;
  const int i = row % per;
  const long long wi = row / per;
  const int wx = wi % nw, wy = (wi / nw) % nw;
  const long long b = wi / ((long long)nw * nw);
  const int ty = wy * win + i / win, tx = wx * win + i % win;
  map[row] = (ty < g && tx < g) ? (int)((b * g + ty) * g + tx) : -1;
}

// f32 -> 16-bit row cast (8 elements per thread)
template <typename T>
__global__ void cast16_kernel(const float* __restrict__ x, T* __restrict__ out, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float f[8];
  reinterpret_cast<float4*>(f)[0] = reinterpret_cast<const float4*>(x)[2 * i];
  reinterpret_cast<float4*>(f)[1] = reinterpret_cast<const float4*>(x)[2 * i + 1];
  reinterpret_cast<uint4*>(out)[i] = pack8f<T>(f);
}

// im2col for 3x3 / pad 1 convs over NHWC 16-bit activations: out[(b,oy,ox), (ky,kx,c)] ; 8 channels/thread.
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Hin, int Win, int C,
                                 int stride, int Hout, int Wout) {
  const int c8 = C / 8;
  const long long total = (long long)B * Hout * Wout * 9 * c8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cc = idx % c8;
  const int tap = (idx / c8) % 9;
  const long long tok = idx / (9LL * c8);
  const int ox = tok % Wout, oy = (tok / Wout) % Hout;
  const long long b = tok / ((long long)Wout * Hout);
  const int iy = oy * stride - 1 + tap / 3, ix = ox * stride - 1 + tap % 3;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
    v = reinterpret_cast<const uint4*>(in + ((b * Hin + iy) * Win + ix) * C)[cc];
  reinterpret_cast<uint4*>(out + tok * 9 * C + (long long)tap * C)[cc] = v;
}

// CLIP embeddings (vision/clip.rs:165-236): row 0 = cls + pos[0]; row 1+i = sam_out[b,i] + pos[1+i].
__global__ void clip_embed_kernel(const float* __restrict__ sam, const float* __restrict__ cls,
                                  const float* __restrict__ pos, float* __restrict__ out, int B, int n, int C) {
  const int c4 = C / 4;
  const long long total = (long long)B * (n + 1) * c4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % c4;
  const int t = (idx / c4) % (n + 1);
  const long long b = idx / ((long long)c4 * (n + 1));
  float4 a = t == 0 ? reinterpret_cast<const float4*>(cls)[c]
                    : reinterpret_cast<const float4*>(sam + (b * n + (t - 1)) * C)[c];
  const float4 p = reinterpret_cast<const float4*>(pos + (long long)t * C)[c];
  a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
  reinterpret_cast<float4*>(out)[idx] = a;
}

// build_clip_sam_tokens (model/mod.rs:604-650): pre[b*n+i] = [clip[b,1+i,:] | sam[b,i,:]] as 16-bit (+ f32 tap).
template <typename T>
__global__ void concat_clip_sam_kernel(const float* __restrict__ clip, const float* __restrict__ sam,
                                       T* __restrict__ out16, float* __restrict__ out32, int B, int n, int C) {
  const int c8 = C / 8;
  const long long total = (long long)B * n * 2 * c8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = idx % c8;
  const int half = (idx / c8) % 2;
  const long long row = idx / (2LL * c8);
  const long long b = row / n, i = row % n;
  const float* src = half == 0 ? clip + ((b * (n + 1)) + 1 + i) * C : sam + row * C;
  float f[8];
  reinterpret_cast<float4*>(f)[0] = reinterpret_cast<const float4*>(src)[2 * c];
  reinterpret_cast<float4*>(f)[1] = reinterpret_cast<const float4*>(src)[2 * c + 1];
  const long long o = row * 2 * C + (long long)half * C + c * 8;
  reinterpret_cast<uint4*>(out16 + o)[0] = pack8f<T>(f);
  if (out32) {
    reinterpret_cast<float4*>(out32 + o)[0] = reinterpret_cast<float4*>(f)[0];
    reinterpret_cast<float4*>(out32 + o)[1] = reinterpret_cast<float4*>(f)[1];
  }
}

// format_global_tokens / format_local_tokens / assemble_artifacts (model/mod.rs:656-709, 879-923) as one
// row scatter: dst row r takes projected row map[r] (>= 0), the image_newline row (-1) or view_seperator (-2).
__global__ void scatter_tokens_kernel(const float* __restrict__ proj, const float* __restrict__ newline,
                                      const float* __restrict__ sep, const int* __restrict__ map,
                                      float* __restrict__ dst, long long rows, int C) {
  const int c4 = C / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c4) return;
  const long long r = idx / c4;
  const int c = idx % c4;
  const int m = map[r];
  const float* src = m >= 0 ? proj + (long long)m * C : (m == -1 ? newline : sep);
  reinterpret_cast<float4*>(dst)[idx] = reinterpret_cast<const float4*>(src)[c];
}

inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }

}  // namespace

#define DISPATCH_T(dt, ...)                                              \
  do {                                                                   \
    if ((dt) == DType::BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else { using T = __half; __VA_ARGS__; }                              \
  } while (0)

void patchify_u8(const uint8_t* img, void* out, int B, int G, DType dt, cudaStream_t s) {
  const long long n = (long long)B * (G / 16) * (G / 16) * 16;
  DISPATCH_T(dt, (patchify_u8_kernel<T><<<blocks_for(n, 256), 256, 0, s>>>(img, (T*)out, B, G)));
  launch_check("patchify_u8");
}
void patchify_f32(const float* img, void* out, int B, int G, DType dt, cudaStream_t s) {
  const long long n = (long long)B * (G / 16) * (G / 16) * 48;
  DISPATCH_T(dt, (patchify_f32_kernel<T><<<blocks_for(n, 256), 256, 0, s>>>(img, (T*)out, B, G)));
  launch_check("patchify_f32");
}
void bcast_rows(const float* src, float* dst, long long rows_per_batch, int batch, int cols, cudaStream_t s) {
  const long long total4 = rows_per_batch * batch * (cols / 4);
  bcast_rows_kernel<<<blocks_for(total4, 256), 256, 0, s>>>(src, dst, rows_per_batch, cols / 4, total4);
  launch_check("bcast_rows");
}
void layernorm(const float* x, const float* w, const float* b, void* out16, float* out32, long long out_rows, int C,
               float eps, int win, int g, int nw, DType dt, cudaStream_t s) {
  const int wpb = 8;
  const int blocks = blocks_for(out_rows, wpb);
  DISPATCH_T(dt, {
    if (C == 768) layernorm_kernel<T, 768><<<blocks, wpb * 32, 0, s>>>(x, w, b, (T*)out16, out32, out_rows, eps, win, g, nw);
    else if (C == 1024) layernorm_kernel<T, 1024><<<blocks, wpb * 32, 0, s>>>(x, w, b, (T*)out16, out32, out_rows, eps, win, g, nw);
    else if (C == 256) layernorm_kernel<T, 256><<<blocks, wpb * 32, 0, s>>>(x, w, b, (T*)out16, out32, out_rows, eps, win, g, nw);
    else throw std::runtime_error("layernorm: unsupported width " + std::to_string(C));
  });
  launch_check("layernorm");
}
void window_row_map(int* map, long long rows, int win, int g, int nw, cudaStream_t s) {
  window_row_map_kernel<<<blocks_for(rows, 256), 256, 0, s>>>(map, rows, win, g, nw);
  launch_check("window_row_map");
}
void cast16(const float* x, void* out, long long n, DType dt, cudaStream_t s) {
  DISPATCH_T(dt, (cast16_kernel<T><<<blocks_for(n / 8, 256), 256, 0, s>>>(x, (T*)out, n / 8)));
  launch_check("cast16");
}
void im2col3x3(const void* in, void* out, int B, int Hin, int Win, int C, int stride, DType dt, cudaStream_t s) {
  const int Hout = (Hin + 2 - 3) / stride + 1, Wout = (Win + 2 - 3) / stride + 1;
  const long long n = (long long)B * Hout * Wout * 9 * (C / 8);
  DISPATCH_T(dt, (im2col3x3_kernel<T><<<blocks_for(n, 256), 256, 0, s>>>((const T*)in, (T*)out, B, Hin, Win, C, stride, Hout, Wout)));
  launch_check("im2col3x3");
}
void clip_embed(const float* sam, const float* cls, const float* pos, float* out, int B, int n, int C, cudaStream_t s) {
  const long long total = (long long)B * (n + 1) * (C / 4);
  clip_embed_kernel<<<blocks_for(total, 256), 256, 0, s>>>(sam, cls, pos, out, B, n, C);
  launch_check("clip_embed");
}
void concat_clip_sam(const float* clip, const float* sam, void* out16, float* out32, int B, int n, int C, DType dt,
                     cudaStream_t s) {
  const long long total = (long long)B * n * 2 * (C / 8);
  DISPATCH_T(dt, (concat_clip_sam_kernel<T><<<blocks_for(total, 256), 256, 0, s>>>(clip, sam, (T*)out16, out32, B, n, C)));
  launch_check("concat_clip_sam");
}
void scatter_tokens(const float* proj, const float* newline, const float* sep, const int* map, float* dst,
                    long long rows, int C, cudaStream_t s) {
  scatter_tokens_kernel<<<blocks_for(rows * (C / 4), 256), 256, 0, s>>>(proj, newline, sep, map, dst, rows, C);
  launch_check("scatter_tokens");
}

}  // namespace dsocr

// ---------------------------------------------------------------------------------------------------
// Integer bicubic resample on the device, bit-exact with vision/resample.rs:101-160 (22-bit fixed-point
// coefficients computed on the host exactly as compute_resample_coeffs does, i64 accumulation, u8 intermediate
// after the horizontal pass).  The vertical pass writes straight into the destination layout: either the
// 127-grey global canvas at (x_off, y_off) (build_global_view, model/mod.rs:2308-2330) or the row-major
// tile stack of dynamic_preprocess (vision/preprocess.rs:113-127).
namespace dsocr {
namespace {
constexpr int kPrecisionBits = 22;

__global__ void resample_h_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw,
                                  const int* __restrict__ start, const int* __restrict__ len,
                                  const int* __restrict__ coef, int ksize) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)sh * dw) return;
  const int x = idx % dw;
  const int y = idx / dw;
  const int s0 = start[x], n = len[x];
  const int* w = coef + (long long)x * ksize;
  const uint8_t* p = src + ((long long)y * sw + s0) * 3;
  long long a0 = 1ll << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int i = 0; i < n; ++i, p += 3) {
    const long long wi = w[i];
    a0 += (long long)p[0] * wi; a1 += (long long)p[1] * wi; a2 += (long long)p[2] * wi;
  }
  uint8_t* o = dst + idx * 3;
  o[0] = (uint8_t)min(255ll, max(0ll, a0 >> kPrecisionBits));
  o[1] = (uint8_t)min(255ll, max(0ll, a1 >> kPrecisionBits));
  o[2] = (uint8_t)min(255ll, max(0ll, a2 >> kPrecisionBits));
}

// tile == 0: write into a canvas of side `canvas` at (x_off, y_off), clipped; tile > 0: write into the tile
// stack [n][tile][tile][3] of a (tiles_w * tile)-wide image.
__global__ void resample_v_kernel(const uint8_t* __restrict__ horiz, int dw, int dh, uint8_t* __restrict__ dst,
                                  const int* __restrict__ start, const int* __restrict__ len,
                                  const int* __restrict__ coef, int ksize, int canvas, int x_off, int y_off, int tile,
                                  int tiles_w) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)dh * dw) return;
  const int x = idx % dw;
  const int y = idx / dw;
  const int s0 = start[y], n = len[y];
  const int* w = coef + (long long)y * ksize;
  const uint8_t* p = horiz + ((long long)s0 * dw + x) * 3;
  long long a0 = 1ll << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int i = 0; i < n; ++i, p += (long long)dw * 3) {
    const long long wi = w[i];
    a0 += (long long)p[0] * wi; a1 += (long long)p[1] * wi; a2 += (long long)p[2] * wi;
  }
  uint8_t* o;
  if (tile > 0) {
    const int t = (y / tile) * tiles_w + (x / tile);
    o = dst + (((long long)t * tile + (y % tile)) * tile + (x % tile)) * 3;
  } else {
    const int cy = y + y_off, cx = x + x_off;
    if (cy < 0 || cy >= canvas || cx < 0 || cx >= canvas) return;
    o = dst + ((long long)cy * canvas + cx) * 3;
  }
  o[0] = (uint8_t)min(255ll, max(0ll, a0 >> kPrecisionBits));
  o[1] = (uint8_t)min(255ll, max(0ll, a1 >> kPrecisionBits));
  o[2] = (uint8_t)min(255ll, max(0ll, a2 >> kPrecisionBits));
}
}  // namespace

void resample_h(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, const int* start, const int* len,
                const int* coef, int ksize, cudaStream_t s) {
  const long long n = (long long)sh * dw;
  resample_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, sw, sh, dst, dw, start, len, coef, ksize);
  launch_check("resample_h");
}
void resample_v(const uint8_t* horiz, int dw, int dh, uint8_t* dst, const int* start, const int* len, const int* coef,
                int ksize, int canvas, int x_off, int y_off, int tile, int tiles_w, cudaStream_t s) {
  const long long n = (long long)dh * dw;
  resample_v_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(horiz, dw, dh, dst, start, len, coef, ksize, canvas, x_off,
                                                               y_off, tile, tiles_w);
  launch_check("resample_v");
}
}  // namespace dsocr
