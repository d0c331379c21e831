// Internal launcher interface shared by the engine and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

namespace dsocr {

enum class DType : int { F32 = 0, F16 = 1, BF16 = 2 };  // values match include/dsocr.h

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------------- tensor-core linear
struct LinearTile { int w_row0, x_row0, rows, n0; };  // == lin::Tile

struct LinearCall {
  // weights: [w_rows, K] 16-bit, row pitch ldw (0 = K).  w1 != nullptr selects the dual (SwiGLU) kernel.
  const void* w0 = nullptr;
  const void* w1 = nullptr;
  long long w_rows = 0;  // 0 = N
  long long ldw = 0;
  // activations: [x_rows, K] 16-bit, row pitch ldx (0 = K); x_parts = 2 adds the lo part at row x_lo_row_off
  const void* x = nullptr;
  long long x_rows = 0;
  long long ldx = 0;
  int x_parts = 1;
  int x_lo_row_off = 0;
  // batched X (3-D): nbatch slices, x_batch_stride elements apart (e.g. attention heads inside a qkv row)
  int nbatch = 1;
  long long x_batch_stride = 0;
  long long out_batch_stride = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  void* out = nullptr;
  void* out_lo = nullptr;
  long long ldo = 0;
  const int* row_map = nullptr;
  int act = 0;       // lin::Act
  int out_mode = 0;  // lin::Out
  // grouped problems
  const LinearTile* tiles = nullptr;
  const int* num_tiles_dev = nullptr;
  int max_tiles = 0;
  int tile_rows_hint = 0;  // typical rows per tile, used to pick the token tile
  int bn = 0;              // force the token tile (0 = auto)
};

int linear_pick_bn(long long m, bool dual);
void linear(const LinearCall& c, DType dt, int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------------- vision attention
// qkv: [rows = B*S, 3, H, 64] 16-bit.  grid > 0: tokens form a grid x grid raster per sequence and the
// decomposed rel-pos bias is added from Z = q . [rel_h ; rel_w]^T  ([rows, H, zw] f32, Zw half at zhalf).
struct VAttnCall {
  const void* qkv = nullptr;
  long long rows = 0;
  int B = 0, S = 0, H = 0;
  int grid = 0;
  const float* Z = nullptr;
  int zw = 0, zhalf = 0;
  void* out = nullptr;  // [rows, H*64] 16-bit
  float scale = 0.125f;
};
bool vision_attention_supported(int grid);
void vision_attention(const VAttnCall& c, DType dt, cudaStream_t stream);
// Z = q . table^T through the tensor-core linear kernel.  table: [2*zhalf, 64] 16-bit (rel_h rows then rel_w rows).
void vision_relpos_products(const void* qkv, long long rows, int H, const void* table, int zhalf, float* Z, DType dt,
                            int num_sms, cudaStream_t stream);

}  // namespace dsocr
