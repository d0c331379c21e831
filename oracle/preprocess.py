"""Integer image preprocessing oracle (numpy).  Test infrastructure only.

Restates, bit-exactly:
  * crates/infer-deepseek/src/vision/resample.rs:1-160   (22-bit fixed-point separable bicubic)
  * crates/infer-deepseek/src/vision/preprocess.rs:67-138 (Gundam dynamic tiling)
  * crates/infer-deepseek/src/model/mod.rs:2295-2347      (global view, image_to_tensor)
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

PRECISION_BITS = 22  # resample.rs:9
ROUNDING_BIAS = 1 << (PRECISION_BITS - 1)  # resample.rs:11


def _round_half_towards_zero(v: float) -> int:
    """resample.rs:18-24 (note: ceil(v+0.5) for negative v)."""
    if v >= 0.0:
        return int(math.floor(v + 0.5))
    return int(math.ceil(v + 0.5))


def _bicubic_kernel(x: float) -> float:
    """resample.rs:26-37, a = -0.5."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def _trunc_to_i32(v: float) -> int:
    # Rust `as i32` truncates toward zero (saturating; never hit here).
    return int(v)


def compute_resample_coeffs(input_size: int, output_size: int):
    """resample.rs:38-99 -> (bounds[(start,len)], coeffs_int[out, ksize], ksize)."""
    scale = input_size / output_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = []
    coeffs = np.zeros((output_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for out_index in range(output_size):
        center = (out_index + 0.5) * scale
        xmin = _round_half_towards_zero(center - support)
        if xmin < 0:
            xmin = 0
        xmax = _round_half_towards_zero(center + support)
        if xmax > input_size:
            xmax = input_size
        if xmin >= input_size:
            xmin = max(input_size - 1, 0)
        if xmax <= xmin:
            xmax = xmin + 1
        length = xmax - xmin
        row = [0.0] * ksize
        total = 0.0
        for i in range(min(length, ksize)):
            w = _bicubic_kernel((xmin + i - center + 0.5) * ss)
            row[i] = w
            total += w
        if total != 0.0:
            for i in range(min(length, ksize)):
                row[i] /= total
        for i in range(ksize):
            v = row[i]
            coeffs[out_index, i] = _trunc_to_i32(-0.5 + v * (1 << PRECISION_BITS)) if v < 0.0 else _trunc_to_i32(
                0.5 + v * (1 << PRECISION_BITS)
            )
        bounds.append((xmin, length))
    return bounds, coeffs, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    """resample.rs:13-16: arithmetic shift then clamp."""
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis0(src: np.ndarray, out_size: int) -> np.ndarray:
    """Apply the 1-D integer filter along axis 0 of an [n, ...] u8 array."""
    bounds, coeffs, _ = compute_resample_coeffs(src.shape[0], out_size)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    src64 = src.astype(np.int64)
    for o, (start, length) in enumerate(bounds):
        w = coeffs[o, :length]
        acc = np.tensordot(w, src64[start : start + length], axes=(0, 0)) + ROUNDING_BIAS
        out[o] = _clip8(acc)
    return out


def resize_bicubic(rgb: np.ndarray, width: int, height: int) -> np.ndarray:
    """resample.rs:101-160.  rgb: [H, W, 3] u8 -> [height, width, 3] u8.

    Horizontal pass first (u8 intermediate), then vertical pass.
    """
    assert rgb.dtype == np.uint8 and rgb.ndim == 3 and rgb.shape[2] == 3
    if width == 0 or height == 0:
        return np.zeros((height, width, 3), dtype=np.uint8)
    horizontal = _resample_axis0(np.ascontiguousarray(rgb.transpose(1, 0, 2)), width).transpose(1, 0, 2)
    return np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(horizontal), height))


def round_ties_to_even(value: float) -> float:
    """model/mod.rs:2295-2306."""
    rounded = math.floor(abs(value) + 0.5) * (1.0 if value >= 0 else -1.0)  # f64::round = half away from zero
    if abs(value - rounded) != 0.5:
        return rounded
    truncated = float(math.trunc(value))
    if int(truncated) % 2 == 0:
        return truncated
    return truncated + (1.0 if value > 0 else -1.0)


def build_global_view(rgb: np.ndarray, base_size: int) -> np.ndarray:
    """model/mod.rs:2308-2330: aspect-preserving resize, centred on a 127-grey canvas."""
    canvas = np.full((base_size, base_size, 3), 127, dtype=np.uint8)
    orig_h, orig_w = rgb.shape[:2]
    if orig_w == 0 or orig_h == 0:
        return canvas
    scale = min(base_size / orig_w, base_size / orig_h)
    new_w = int(min(max(round_ties_to_even(orig_w * scale), 1.0), float(base_size)))
    new_h = int(min(max(round_ties_to_even(orig_h * scale), 1.0), float(base_size)))
    resized = resize_bicubic(rgb, new_w, new_h)
    x_off = int(round_ties_to_even((base_size - new_w) * 0.5))
    y_off = int(round_ties_to_even((base_size - new_h) * 0.5))
    # imageops::replace clips to the canvas.
    h = min(new_h, base_size - y_off)
    w = min(new_w, base_size - x_off)
    canvas[y_off : y_off + h, x_off : x_off + w] = resized[:h, :w]
    return canvas


def select_tile_grid(orig_w: int, orig_h: int, tile: int, min_num: int = 2, max_num: int = 9) -> Tuple[int, int]:
    """preprocess.rs:82-111 -> (w_ratio, h_ratio).  BTreeSet order = sorted tuples."""
    aspect = orig_w / orig_h
    ratios = sorted(
        {
            (i, j)
            for n in range(min_num, max_num + 1)
            for i in range(1, n + 1)
            for j in range(1, n + 1)
            if min_num <= i * j <= max_num
        }
    )
    best = (1, 1)
    best_diff = float("inf")
    area = float(orig_w * orig_h)
    eps = np.finfo(np.float64).eps
    for (wr, hr) in ratios:
        diff = abs(aspect - wr / hr)
        if diff < best_diff:
            best_diff = diff
            best = (wr, hr)
        elif abs(diff - best_diff) < eps and area > 0.5 * float(tile * tile * wr * hr):
            best = (wr, hr)
    return best


def dynamic_preprocess(rgb: np.ndarray, tile: int = 640, min_num: int = 2, max_num: int = 9,
                       no_crop_threshold: Optional[int] = None) -> Tuple[List[np.ndarray], Tuple[int, int]]:
    """preprocess.rs:67-138 with PreprocessParams::ocr1 (:17-25): threshold = tile size."""
    if no_crop_threshold is None:
        no_crop_threshold = tile
    orig_h, orig_w = rgb.shape[:2]
    if orig_w <= no_crop_threshold and orig_h <= no_crop_threshold:
        return [], (1, 1)
    wr, hr = select_tile_grid(orig_w, orig_h, tile, min_num, max_num)
    resized = resize_bicubic(rgb, tile * wr, tile * hr)
    tiles = []
    for i in range(wr * hr):
        x = (i % wr) * tile
        y = (i // wr) * tile
        tiles.append(np.ascontiguousarray(resized[y : y + tile, x : x + tile]))
    return tiles, (wr, hr)


def image_to_tensor(rgb: np.ndarray) -> np.ndarray:
    """model/mod.rs:2332-2347: HWC u8 -> CHW f32, (v/255 - 0.5)/0.5 computed in f32."""
    v = rgb.astype(np.float32) / np.float32(255.0)
    v = (v - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(v.transpose(2, 0, 1))


def prepare_vision_input(rgb: np.ndarray, base_size: int, image_size: int, crop_mode: bool):
    """model/mod.rs:1707-1758 -> dict(global_u8, tiles_u8, crop_shape)."""
    global_size = base_size if crop_mode else image_size
    global_view = build_global_view(rgb, global_size)
    tiles: List[np.ndarray] = []
    crop_shape = None
    if crop_mode:
        tiles, crop_shape = dynamic_preprocess(rgb, tile=image_size)
    return {"global": global_view, "tiles": tiles, "crop_shape": crop_shape}


def image_token_count(base_size: int, image_size: int, crop_mode: bool, crop_shape) -> int:
    """model/mod.rs:2605-2689 (OCR-1 branch)."""
    def q(sz):
        return int(math.ceil((sz // 16) / 4))
    if crop_mode:
        qg, ql = q(base_size), q(image_size)
        wc, hc = crop_shape or (1, 1)
        n = 0
        if wc > 1 or hc > 1:
            n += (ql * hc) * (ql * wc + 1)
        n += qg * (qg + 1) + 1
        return n
    qq = q(image_size)
    return qq * (qq + 1) + 1


def synthetic_page(width: int, height: int, seed: int) -> np.ndarray:
    """SURVEY.md 8(d) config 2/3 page generator: white page, seeded black 'text line' boxes
    (line pitch 24 px, glyph boxes 8-14 px wide, 85 % row fill)."""
    rng = np.random.RandomState(seed)
    page = np.full((height, width, 3), 255, dtype=np.uint8)
    margin = 48
    y = margin
    while y + 16 < height - margin:
        x = margin
        row_end = margin + int((width - 2 * margin) * (0.55 + 0.45 * rng.rand()))
        while x < row_end:
            w = int(rng.randint(8, 15))
            if rng.rand() < 0.85:
                shade = int(rng.randint(0, 64))
                page[y : y + 14, x : x + w] = shade
            x += w + 3
        y += 24
    return page
