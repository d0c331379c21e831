// Host-side table preparation + integer image preprocessing (see hostmath.cpp for reference citations).
#pragma once
#include <stdint.h>
#include <vector>

namespace dsocr {

std::vector<float> resize_table_aa(const float* table, int in_h, int in_w, int C, int out_h, int out_w);
std::vector<float> resize_rel_pos(const float* rel, int orig_len, int hd, int size);
void rope_tables(float theta, int dim, int max_pos, std::vector<float>& cos_t, std::vector<float>& sin_t);

// compute_resample_coeffs (vision/resample.rs:38-99): per output index (start, len) and ksize i32 coefficients
struct ResampleCoeffs { std::vector<int> start, len; std::vector<int32_t> coef; int ksize = 0; };
ResampleCoeffs resample_coeffs_public(int in_size, int out_size);
// aspect-preserving size + centring offsets of build_global_view (model/mod.rs:2308-2330)
void global_view_geometry(int w, int h, int base, int* nw, int* nh, int* x_off, int* y_off);
void resize_bicubic_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void build_global_view_u8(const uint8_t* rgb, int w, int h, int base, uint8_t* out);
void select_tile_grid(int w, int h, int tile, int min_num, int max_num, int* gw, int* gh);
// returns the number of tiles (0 when the image is <= tile in both dims); tiles_out may be null to query.
int dynamic_preprocess_u8(const uint8_t* rgb, int w, int h, int tile, uint8_t* tiles_out, int* gw, int* gh);
int image_token_count(int base_size, int image_size, int crop_mode, int crop_w, int crop_h);

}  // namespace dsocr
