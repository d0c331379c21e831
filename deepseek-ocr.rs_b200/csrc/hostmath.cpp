// Host-side, once-per-resolution table preparation and the integer image preprocessing of the hot path.
//  * antialiased bicubic resize of the SAM / CLIP position tables (vision/sam.rs:1000-1123,
//    vision/clip.rs:486-544): the reference redoes this on the host for EVERY forward; here it runs once
//    per token-grid size and is cached on the device.
//  * linear interpolation of the rel-pos tables (vision/sam.rs:1194-1232)
//  * RoPE cos/sin tables (transformer/rope.rs:172-207)
//  * Pillow-style 22-bit fixed-point bicubic (vision/resample.rs), Gundam tiling (vision/preprocess.rs:67-138),
//    global view (model/mod.rs:2295-2330) - bit-exact integer arithmetic.
#include "hostmath.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <stdexcept>

namespace dsocr {

namespace {
inline float bicubic_filter(float x) {
  const float a = -0.5f;
  x = std::fabs(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}

struct AxisWeights {
  std::vector<int> start, count;
  std::vector<std::vector<float>> w;
};

AxisWeights axis_weights_aa(int in_len, int out_len) {
  const float scale = (float)in_len / (float)out_len;
  const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
  const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  AxisWeights aw;
  aw.start.resize(out_len); aw.count.resize(out_len); aw.w.resize(out_len);
  for (int o = 0; o < out_len; ++o) {
    const float center = scale * ((float)o + 0.5f);
    int xmin = (int)std::floor(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xmax = (int)std::floor(center + support + 0.5f);
    if (xmax > in_len) xmax = in_len;
    const int n = std::max(0, xmax - xmin);
    aw.start[o] = xmin; aw.count[o] = n;
    aw.w[o].resize(n);
    const float xmin_m_center = (float)xmin - center;
    float total = 0.f;
    for (int j = 0; j < n; ++j) {
      const float wv = bicubic_filter(((float)j + xmin_m_center + 0.5f) * invscale);
      aw.w[o][j] = wv; total += wv;
    }
    if (total != 0.f) for (auto& v : aw.w[o]) v /= total;
  }
  return aw;
}
}  // namespace

// table: [in_h, in_w, C] (channels last) -> [out_h, out_w, C]; vertical pass then horizontal (sam.rs:1084-1115).
std::vector<float> resize_table_aa(const float* table, int in_h, int in_w, int C, int out_h, int out_w) {
  if (in_h == out_h && in_w == out_w) return std::vector<float>(table, table + (size_t)in_h * in_w * C);
  AxisWeights wy = axis_weights_aa(in_h, out_h), wx = axis_weights_aa(in_w, out_w);
  std::vector<float> tmp((size_t)out_h * in_w * C, 0.f), out((size_t)out_h * out_w * C, 0.f);
  for (int oh = 0; oh < out_h; ++oh)
    for (int k = 0; k < wy.count[oh]; ++k) {
      const float w = wy.w[oh][k];
      const float* src = table + (size_t)(wy.start[oh] + k) * in_w * C;
      float* dst = tmp.data() + (size_t)oh * in_w * C;
      for (size_t i = 0; i < (size_t)in_w * C; ++i) dst[i] += src[i] * w;
    }
  for (int oh = 0; oh < out_h; ++oh)
    for (int ow = 0; ow < out_w; ++ow) {
      float* dst = out.data() + ((size_t)oh * out_w + ow) * C;
      for (int k = 0; k < wx.count[ow]; ++k) {
        const float w = wx.w[ow][k];
        const float* src = tmp.data() + ((size_t)oh * in_w + wx.start[ow] + k) * C;
        for (int c = 0; c < C; ++c) dst[c] += src[c] * w;
      }
    }
  return out;
}

// rel table [orig_len, hd] -> [2*size-1, hd]  (get_rel_pos_vec, q_size == k_size == size)
std::vector<float> resize_rel_pos(const float* rel, int orig_len, int hd, int size) {
  const int max_rel = 2 * size - 1;
  if (orig_len == max_rel) return std::vector<float>(rel, rel + (size_t)orig_len * hd);
  std::vector<float> out((size_t)max_rel * hd);
  const float scale = (float)orig_len / (float)max_rel;
  for (int i = 0; i < max_rel; ++i) {
    float src = scale * ((float)i + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    const float max_src = (float)(orig_len - 1);
    if (src > max_src) src = max_src;
    const float left_f = std::floor(src);
    const int left = (int)left_f;
    const int right = std::min(left + 1, orig_len - 1);
    float w = src - left_f;
    w = std::min(std::max(w, 0.f), 1.f);
    for (int d = 0; d < hd; ++d)
      out[(size_t)i * hd + d] = rel[(size_t)left * hd + d] * (1.0f - w) + rel[(size_t)right * hd + d] * w;
  }
  return out;
}

void rope_tables(float theta, int dim, int max_pos, std::vector<float>& cos_t, std::vector<float>& sin_t) {
  const int half = dim / 2;
  std::vector<float> inv(half);
  for (int i = 0; i < half; ++i) inv[i] = 1.0f / std::pow(theta, ((float)i * 2.0f) / (float)dim);
  cos_t.resize((size_t)max_pos * half);
  sin_t.resize((size_t)max_pos * half);
  for (int p = 0; p < max_pos; ++p)
    for (int i = 0; i < half; ++i) {
      const float ang = (float)p * inv[i];
      cos_t[(size_t)p * half + i] = std::cos(ang);
      sin_t[(size_t)p * half + i] = std::sin(ang);
    }
}

// ------------------------------------------------------------------------------------------------
// Integer preprocessing
namespace {
constexpr int kPrecisionBits = 22;
constexpr int64_t kRoundingBias = 1ll << (kPrecisionBits - 1);

inline long round_half_towards_zero(double v) { return v >= 0.0 ? (long)std::floor(v + 0.5) : (long)std::ceil(v + 0.5); }

inline double bicubic_kernel_f64(double v) {
  const double a = -0.5;
  const double x = std::fabs(v);
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
  return 0.0;
}

struct Coeffs {
  std::vector<int> start, len;
  std::vector<int32_t> c;
  int ksize = 0;
};

Coeffs resample_coeffs(int in_size, int out_size) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = std::max(scale, 1.0);
  const double support = 2.0 * filterscale;
  Coeffs k;
  k.ksize = (int)std::ceil(support) * 2 + 1;
  k.start.resize(out_size); k.len.resize(out_size);
  k.c.assign((size_t)out_size * k.ksize, 0);
  std::vector<double> row(k.ksize);
  const double ss = 1.0 / filterscale;
  for (int o = 0; o < out_size; ++o) {
    const double center = ((double)o + 0.5) * scale;
    long xmin = round_half_towards_zero(center - support);
    if (xmin < 0) xmin = 0;
    long xmax = round_half_towards_zero(center + support);
    if (xmax > in_size) xmax = in_size;
    if (xmin >= in_size) xmin = std::max(in_size - 1, 0);
    if (xmax <= xmin) xmax = xmin + 1;
    const int length = (int)(xmax - xmin);
    std::fill(row.begin(), row.end(), 0.0);
    double sum = 0.0;
    for (int i = 0; i < length && i < k.ksize; ++i) {
      const double w = bicubic_kernel_f64(((double)xmin + (double)i - center + 0.5) * ss);
      row[i] = w; sum += w;
    }
    if (sum != 0.0) for (int i = 0; i < length && i < k.ksize; ++i) row[i] /= sum;
    for (int i = 0; i < k.ksize; ++i) {
      const double v = row[i];
      k.c[(size_t)o * k.ksize + i] = v < 0.0 ? (int32_t)(-0.5 + v * (double)(1 << kPrecisionBits))
                                             : (int32_t)(0.5 + v * (double)(1 << kPrecisionBits));
    }
    k.start[o] = (int)xmin; k.len[o] = length;
  }
  return k;
}

inline uint8_t clip8(int64_t v) {
  const int64_t s = v >> kPrecisionBits;
  return (uint8_t)std::min<int64_t>(255, std::max<int64_t>(0, s));
}

double round_ties_to_even(double value) {
  const double rounded = std::round(value);
  if (std::fabs(value - rounded) != 0.5) return rounded;
  const double truncated = std::trunc(value);
  if ((long long)truncated % 2 == 0) return truncated;
  return truncated + (value > 0 ? 1.0 : (value < 0 ? -1.0 : 0.0));
}
}  // namespace

ResampleCoeffs resample_coeffs_public(int in_size, int out_size) {
  Coeffs c = resample_coeffs(in_size, out_size);
  ResampleCoeffs r;
  r.start = std::move(c.start); r.len = std::move(c.len); r.coef = std::move(c.c); r.ksize = c.ksize;
  return r;
}

void global_view_geometry(int w, int h, int base, int* nw, int* nh, int* x_off, int* y_off) {
  const double scale = std::min((double)base / (double)w, (double)base / (double)h);
  *nw = (int)std::min(std::max(round_ties_to_even((double)w * scale), 1.0), (double)base);
  *nh = (int)std::min(std::max(round_ties_to_even((double)h * scale), 1.0), (double)base);
  *x_off = (int)round_ties_to_even(((double)base - (double)*nw) * 0.5);
  *y_off = (int)round_ties_to_even(((double)base - (double)*nh) * 0.5);
}

void resize_bicubic_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  if (dw == 0 || dh == 0) return;
  const Coeffs cx = resample_coeffs(sw, dw), cy = resample_coeffs(sh, dh);
  std::vector<uint8_t> horiz((size_t)sh * dw * 3);
  for (int y = 0; y < sh; ++y) {
    const uint8_t* srow = src + (size_t)y * sw * 3;
    uint8_t* hrow = horiz.data() + (size_t)y * dw * 3;
    for (int x = 0; x < dw; ++x) {
      const int32_t* w = &cx.c[(size_t)x * cx.ksize];
      int64_t a0 = kRoundingBias, a1 = kRoundingBias, a2 = kRoundingBias;
      const uint8_t* p = srow + (size_t)cx.start[x] * 3;
      for (int i = 0; i < cx.len[x]; ++i, p += 3) {
        a0 += (int64_t)p[0] * w[i]; a1 += (int64_t)p[1] * w[i]; a2 += (int64_t)p[2] * w[i];
      }
      hrow[x * 3] = clip8(a0); hrow[x * 3 + 1] = clip8(a1); hrow[x * 3 + 2] = clip8(a2);
    }
  }
  std::vector<int64_t> acc((size_t)dw * 3);
  for (int y = 0; y < dh; ++y) {
    std::fill(acc.begin(), acc.end(), kRoundingBias);
    const int32_t* w = &cy.c[(size_t)y * cy.ksize];
    for (int i = 0; i < cy.len[y]; ++i) {
      const uint8_t* hrow = horiz.data() + (size_t)(cy.start[y] + i) * dw * 3;
      const int64_t wi = w[i];
      for (int j = 0; j < dw * 3; ++j) acc[j] += (int64_t)hrow[j] * wi;
    }
    uint8_t* drow = dst + (size_t)y * dw * 3;
    for (int j = 0; j < dw * 3; ++j) drow[j] = clip8(acc[j]);
  }
}

void build_global_view_u8(const uint8_t* rgb, int w, int h, int base, uint8_t* out) {
  std::memset(out, 127, (size_t)base * base * 3);
  if (w == 0 || h == 0) return;
  const double scale = std::min((double)base / (double)w, (double)base / (double)h);
  const int nw = (int)std::min(std::max(round_ties_to_even((double)w * scale), 1.0), (double)base);
  const int nh = (int)std::min(std::max(round_ties_to_even((double)h * scale), 1.0), (double)base);
  std::vector<uint8_t> resized((size_t)nw * nh * 3);
  resize_bicubic_u8(rgb, w, h, resized.data(), nw, nh);
  const int xo = (int)round_ties_to_even(((double)base - (double)nw) * 0.5);
  const int yo = (int)round_ties_to_even(((double)base - (double)nh) * 0.5);
  const int ch = std::min(nh, base - yo), cw = std::min(nw, base - xo);
  for (int y = 0; y < ch; ++y)
    std::memcpy(out + ((size_t)(yo + y) * base + xo) * 3, resized.data() + (size_t)y * nw * 3, (size_t)cw * 3);
}

void select_tile_grid(int w, int h, int tile, int min_num, int max_num, int* gw, int* gh) {
  const double aspect = (double)w / (double)h;
  std::set<std::pair<int, int>> ratios;
  for (int n = min_num; n <= max_num; ++n)
    for (int i = 1; i <= n; ++i)
      for (int j = 1; j <= n; ++j)
        if (i * j <= max_num && i * j >= min_num) ratios.insert({i, j});
  std::pair<int, int> best{1, 1};
  double best_diff = 1.79769313486231570e308;
  const double area = (double)((unsigned)w * (unsigned)h);
  for (auto& r : ratios) {
    const double diff = std::fabs(aspect - (double)r.first / (double)r.second);
    if (diff < best_diff) { best_diff = diff; best = r; }
    else if (std::fabs(diff - best_diff) < 2.220446049250313e-16 &&
             area > 0.5 * (double)((unsigned)tile * (unsigned)tile * (unsigned)r.first * (unsigned)r.second))
      best = r;
  }
  *gw = best.first; *gh = best.second;
}

int dynamic_preprocess_u8(const uint8_t* rgb, int w, int h, int tile, uint8_t* tiles_out, int* gw, int* gh) {
  if (w <= tile && h <= tile) { *gw = 1; *gh = 1; return 0; }
  select_tile_grid(w, h, tile, 2, 9, gw, gh);
  const int n = (*gw) * (*gh);
  if (!tiles_out) return n;
  const int tw = tile * (*gw), th = tile * (*gh);
  std::vector<uint8_t> resized((size_t)tw * th * 3);
  resize_bicubic_u8(rgb, w, h, resized.data(), tw, th);
  for (int i = 0; i < n; ++i) {
    const int x = (i % *gw) * tile, y = (i / *gw) * tile;
    uint8_t* dst = tiles_out + (size_t)i * tile * tile * 3;
    for (int r = 0; r < tile; ++r)
      std::memcpy(dst + (size_t)r * tile * 3, resized.data() + ((size_t)(y + r) * tw + x) * 3, (size_t)tile * 3);
  }
  return n;
}

int image_token_count(int base_size, int image_size, int crop_mode, int crop_w, int crop_h) {
  auto q = [](int sz) { return (int)std::ceil((float)(sz / 16) / 4.0f); };
  if (crop_mode) {
    const int qg = q(base_size), ql = q(image_size);
    int n = 0;
    if (crop_w > 1 || crop_h > 1) n += (ql * crop_h) * (ql * crop_w + 1);
    n += qg * (qg + 1) + 1;
    return n;
  }
  const int qq = q(image_size);
  return qq * (qq + 1) + 1;
}

}  // namespace dsocr
