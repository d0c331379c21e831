// Host-side TMA descriptor construction (cuTensorMapEncodeTiled fetched through the runtime's
// driver entry point so the library links against libcudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

namespace tmap {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major 16-bit tensor [rows, cols] with row pitch `ld_elems`; box = [box_rows, box_cols],
// 128-byte swizzle (box_cols * 2 must be 128), zero fill out of bounds.
inline CUtensorMap make_2d_16bit(const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                                 uint32_t box_cols) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled(2d) failed: code " + std::to_string((int)r) + " rows=" +
                             std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld_elems));
  return m;
}

// 2-D 16-bit row-major tensor for TMA *stores* from an unswizzled shared-memory tile [box_rows][box_cols]
inline CUtensorMap make_2d_16bit_store(const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                                       uint32_t box_rows, uint32_t box_cols) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw std::runtime_error("cuTensorMapEncodeTiled(store) failed: code " + std::to_string((int)r) + " rows=" +
                             std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld_elems));
  return m;
}

// 3-D 16-bit tensor [d2, d1, d0] (d0 innermost) with byte strides s1 (dim1) and s2 (dim2).
inline CUtensorMap make_3d_16bit(const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  CUtensorMap m;
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(3d) failed: code " + std::to_string((int)r));
  return m;
}

}  // namespace tmap
