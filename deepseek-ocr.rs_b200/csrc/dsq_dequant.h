// Dequantisation of one 64-wide k-block of one weight row from the device plane layouts of QuantWeight (dsq.h) - the
// unit of work of the dequant-fused tensor-core GEMM (one thread per weight row of a 128 x 64 MMA stage, see
// linear_dq.cuh).  __host__ __device__ so that the index arithmetic is checked on the CPU against the
// oracle's dequantisers (tests/test_dsq_dequant_cpu.py) before it ever runs on a GPU.
//   Q8_0 : a = int8 [rows][K], b = f16 d [rows][K/32]
//   Q4_K : a = 144-byte blocks as on disk [rows][K/256]
//   Q6_K : a = ql [rows][K/2], b = qh [rows][K/4], c = int8 scales [rows][K/16], d = f16 d [rows][K/256]
//   F32  : a = float [rows][K]
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define DSQ_HD __host__ __device__ __forceinline__
#else
#define DSQ_HD inline
#endif

namespace dsocr {

DSQ_HD float dsq_f16_to_f32(uint16_t h) {  // IEEE half -> float, no intrinsics (host + device)
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {  // subnormal: normalise
      int e = -1;
      do { ++e; man <<= 1; } while ((man & 0x400u) == 0);
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112u) << 23) | (man << 13);
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

struct DsqPlanes { const uint8_t* a; const uint8_t* b; const uint8_t* c; const uint8_t* d; };

// out[0..63] = dequant(W)[row, kb*64 .. kb*64 + 63];  fmt: 8 / 12 / 14 / 0
DSQ_HD void dsq_dequant64(int fmt, const DsqPlanes& p, long long row, int K, int kb, float* out) {
  const int k0 = kb * 64;
  if (fmt == 8) {
    const int8_t* q = reinterpret_cast<const int8_t*>(p.a) + row * K + k0;
    const uint16_t* d = reinterpret_cast<const uint16_t*>(p.b) + row * (K / 32) + (k0 >> 5);
    const float d0 = dsq_f16_to_f32(d[0]), d1 = dsq_f16_to_f32(d[1]);
    for (int i = 0; i < 32; ++i) { out[i] = d0 * (float)q[i]; out[32 + i] = d1 * (float)q[32 + i]; }
  } else if (fmt == 12) {
    const int sb = k0 >> 8, g = (k0 >> 6) & 3;  // 64-weight group g of super-block sb: sub-blocks 2g (low nibbles), 2g+1 (high)
    const uint8_t* blk = p.a + (row * (K / 256) + sb) * 144;
    uint16_t dh, mh;
    memcpy(&dh, blk, 2); memcpy(&mh, blk + 2, 2);
    const float d = dsq_f16_to_f32(dh), dmin = dsq_f16_to_f32(mh);
    const uint8_t* s = blk + 4;
    int sc[2], mn[2];
    for (int h = 0; h < 2; ++h) {  // ggml get_scale_min_k4
      const int j = 2 * g + h;
      if (j < 4) { sc[h] = s[j] & 63; mn[h] = s[j + 4] & 63; }
      else { sc[h] = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4); mn[h] = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
    }
    const uint8_t* q = blk + 16 + g * 32;
    const float d1 = d * (float)sc[0], m1 = dmin * (float)mn[0], d2 = d * (float)sc[1], m2 = dmin * (float)mn[1];
    for (int l = 0; l < 32; ++l) { out[l] = d1 * (float)(q[l] & 0xF) - m1; out[32 + l] = d2 * (float)(q[l] >> 4) - m2; }
  } else if (fmt == 14) {
    const int sb = k0 >> 8, within = k0 & 255, half = within >> 7, upper = (within >> 6) & 1;  // upper: weights 64..127 of the half
    const uint8_t* ql = p.a + row * (K / 2) + sb * 128 + half * 64;
    const uint8_t* qh = p.b + row * (K / 4) + sb * 64 + half * 32;
    const int8_t* sc = reinterpret_cast<const int8_t*>(p.c) + row * (K / 16) + sb * 16 + half * 8;
    const float d = dsq_f16_to_f32(reinterpret_cast<const uint16_t*>(p.d)[row * (K / 256) + sb]);
    for (int l = 0; l < 32; ++l) {
      const int is = l >> 4;
      int qa, qb;  // weights l and 32 + l of this 64-wide block
      if (!upper) {
        qa = (int)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
        qb = (int)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
        out[l] = d * (float)sc[is + 0] * (float)qa;
        out[32 + l] = d * (float)sc[is + 2] * (float)qb;
      } else {
        qa = (int)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
        qb = (int)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
        out[l] = d * (float)sc[is + 4] * (float)qa;
        out[32 + l] = d * (float)sc[is + 6] * (float)qb;
      }
    }
  } else {
    const float* w = reinterpret_cast<const float*>(p.a) + row * K + k0;
    for (int i = 0; i < 64; ++i) out[i] = w[i];
  }
}

}  // namespace dsocr
