#!/usr/bin/env python
"""Generates the committed golden vectors of tests/golden/ from implementations that are INDEPENDENT of this repo's
oracle and kernels (run in the build container; the vectors travel, the generators need not exist on the GPU box):

  resample_pillow.npz   Pillow's fixed-point BICUBIC on downscales == vision/resample.rs (bit-exact)
  letterbox_pillow.npz  a 2852x1756 page (the aspect of assets/sample_1.png) letterboxed to 1024x630 on the 127 canvas,
                        built with Pillow only (resize + paste) == build_global_view (model/mod.rs:2308-2330)
  dsq_blocks.npz        random valid Q8_0 / Q4_K / Q6_K ggml blocks and their gguf-py dequantisation; a seeded weight and
                        gguf-py's Q8_0 quantisation of it (== dsq-writer/src/lib.rs:555-598)
  aa_bicubic.npz        torch F.interpolate(bicubic, antialias=True, align_corners=False) for the pos-embed resizes
                        64->40 (SAM, sam.rs:1000-1123) and 16->10 (CLIP, clip.rs:486-544)
  sam_relpos_vllm.npz   get_rel_pos / decomposed rel-pos bias from vLLM's PyTorch deepencoder.py (extracted by AST)
  geometry.json         tile grids / token counts / letterbox geometry the reference was probed for

Usage: python tests/golden/make_golden.py   (rewrites the files next to this script)"""
from __future__ import annotations

import ast
import json
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def resample():
    from PIL import Image

    out = {}
    for i, (src, dst) in enumerate([((97, 61), (40, 32)), ((130, 77), (64, 48)), ((211, 160), (128, 100)), ((64, 200), (33, 90))]):
        rng = np.random.RandomState(100 + i)
        img = rng.randint(0, 256, (src[1], src[0], 3), dtype=np.uint8)
        out[f"src{i}"] = img
        out[f"dst{i}"] = np.asarray(Image.fromarray(img).resize(dst, Image.BICUBIC))
    np.savez_compressed(HERE / "resample_pillow.npz", **out)


def letterbox():
    from PIL import Image

    # low-entropy page so that the fixture stays small: grey ramps + a few rectangles, 2852x1756 like sample_1.png
    h, w = 1756, 2852
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx // 23 * 7) % 256, (yy // 17 * 11) % 256, ((xx // 64 + yy // 64) * 37) % 256], -1).astype(np.uint8)
    small = Image.fromarray(img).resize((1024, 630), Image.BICUBIC)   # round_ties_to_even(1756 * 1024 / 2852) = 630
    canvas = Image.new("RGB", (1024, 1024), (127, 127, 127))
    canvas.paste(small, (0, 197))                                       # round_ties_to_even((1024 - 630) / 2) = 197
    got = np.asarray(canvas)
    np.savez_compressed(HERE / "letterbox_pillow.npz", rows=got[190:210].copy(), row_sum=got.astype(np.int64).sum(axis=(1, 2)),
                        col_sum=got.astype(np.int64).sum(axis=(0, 2)))


def dsq_blocks():
    import gguf
    from gguf import quants

    rng = np.random.RandomState(7)
    out = {}
    w = (rng.randn(8, 512) * 0.05).astype(np.float32)
    w[3, 64:96] = 0.0  # an all-zero block: d = 0, q = 0
    out["q8_weight"] = w
    out["q8_bytes"] = np.frombuffer(quants.quantize(w, gguf.GGMLQuantizationType.Q8_0).tobytes(), dtype=np.uint8)
    for name, gg, nb in (("q8_0", gguf.GGMLQuantizationType.Q8_0, 34), ("q4k", gguf.GGMLQuantizationType.Q4_K, 144),
                         ("q6k", gguf.GGMLQuantizationType.Q6_K, 210)):
        rows, blocks = 4, 3
        raw = rng.randint(0, 256, (rows, blocks * nb), dtype=np.uint8)
        view = raw.reshape(rows, blocks, nb)
        scale = np.float16(0.01).view(np.uint16)
        lo, hi = int(scale) & 0xFF, int(scale) >> 8
        if name == "q8_0":
            view[:, :, 0] = lo; view[:, :, 1] = hi
        elif name == "q4k":
            view[:, :, 0] = lo; view[:, :, 1] = hi; view[:, :, 2] = lo; view[:, :, 3] = hi
        else:
            view[:, :, 208] = lo; view[:, :, 209] = hi
        out[f"{name}_raw"] = raw
        out[f"{name}_deq"] = quants.dequantize(raw, gg).astype(np.float32)
    np.savez_compressed(HERE / "dsq_blocks.npz", **out)


def aa_bicubic():
    import torch
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(3)
    out = {}
    for name, (c, s, d) in {"sam_64_40": (8, 64, 40), "clip_16_10": (8, 16, 10), "up_16_24": (4, 16, 24)}.items():
        x = torch.randn(1, c, s, s, generator=g)
        out[name + "_in"] = x.numpy()
        out[name + "_out"] = F.interpolate(x, size=(d, d), mode="bicubic", antialias=True, align_corners=False).numpy()
    np.savez_compressed(HERE / "aa_bicubic.npz", **out)


def sam_relpos():
    import torch

    path = Path(torch.__file__).resolve().parent.parent / "vllm/model_executor/models/deepencoder.py"
    tree = ast.parse(path.read_text())
    want = {"get_rel_pos", "add_decomposed_rel_pos"}
    src = "\n".join(ast.get_source_segment(path.read_text(), n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want)
    ns = {"torch": torch, "F": torch.nn.functional, "math": __import__("math")}
    exec("from typing import *\n" + src, ns)
    g = torch.Generator().manual_seed(5)
    out = {}
    for name, (qs, ks, rows) in {"window14": (14, 14, 27), "global40_from64": (40, 40, 127)}.items():
        table = torch.randn(rows, 64, generator=g)
        out[name + "_table"] = table.numpy()
        out[name + "_relpos"] = ns["get_rel_pos"](qs, ks, table).numpy()
    B, H, W, C = 2, 6, 5, 64
    q = torch.randn(B, H * W, C, generator=g)
    rh, rw = torch.randn(2 * H - 1, C, generator=g), torch.randn(2 * W - 1, C, generator=g)
    rel_h, rel_w = ns["add_decomposed_rel_pos"](q, rh, rw, (H, W), (H, W))
    out["bias_q"], out["bias_rh"], out["bias_rw"] = q.numpy(), rh.numpy(), rw.numpy()
    out["bias_rel_h"], out["bias_rel_w"] = rel_h.numpy(), rel_w.numpy()
    np.savez_compressed(HERE / "sam_relpos_vllm.npz", **out)


def geometry():
    facts = {
        "source": "SURVEY.md section 8 (probes of the reference: model/mod.rs:2605-2689, vision/preprocess.rs:67-138)",
        "tile_grid": [{"w": 1654, "h": 2339, "grid": [2, 3]}, {"w": 2852, "h": 1756, "grid": [3, 2]}],
        "image_tokens": [{"base": 1024, "image": 640, "crop": True, "grid": [2, 3], "tokens": 903},
                         {"base": 1024, "image": 1024, "crop": False, "grid": [1, 1], "tokens": 273},
                         {"base": 1024, "image": 640, "crop": True, "grid": [1, 1], "tokens": 273}],
        "letterbox": {"w": 2852, "h": 1756, "base": 1024, "inner_h": 630, "top": 197, "fill": 127},
        "no_tiles_when_small": {"w": 640, "h": 600, "grid": [1, 1], "tiles": 0},
    }
    (HERE / "geometry.json").write_text(json.dumps(facts, indent=1) + "\n")


if __name__ == "__main__":
    resample(); letterbox(); dsq_blocks(); aa_bicubic(); geometry()
    try:
        sam_relpos()
    except Exception as ex:  # vLLM's source layout may differ in another image; the other vectors do not depend on it
        print("sam_relpos skipped:", ex)
    for f in sorted(HERE.glob("*.np*")) + sorted(HERE.glob("*.json")):
        print(f.name, f.stat().st_size)
