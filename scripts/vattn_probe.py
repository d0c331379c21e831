#!/usr/bin/env python
"""Runs the vision attention kernel (csrc/attention_tc.cuh) once through its test hook on random inputs of a chosen shape,
so that `ncu --set full --import-source on -k regex:vattn_kernel` has a launch that fills the GPU (the parity tests use a
few dozen CTAs).  python scripts/vattn_probe.py --grid 64 --B 4 --H 12   (grid 0 = CLIP, no bias; --S then sets the length)"""
import argparse
import ctypes
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))
from dsocr.binding import check, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--S", type=int, default=257)
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--H", type=int, default=12)
    ap.add_argument("--dtype", default="bf16")
    a = ap.parse_args()
    S = a.grid * a.grid if a.grid else a.S
    rng = np.random.default_rng(0)
    qkv = rng.standard_normal((a.B * S, 3 * a.H * 64), dtype=np.float32)
    out = np.zeros((a.B * S, a.H * 64), dtype=np.float32)
    fp = lambda x: x.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if x is not None else None
    rh = rw = None
    if a.grid:
        rh = (rng.standard_normal((2 * a.grid - 1, 64), dtype=np.float32) * 0.2)
        rw = (rng.standard_normal((2 * a.grid - 1, 64), dtype=np.float32) * 0.2)
    st = lib().dsocr_test_vision_attention(2 if a.dtype == "bf16" else 1, a.B, S, a.H, fp(qkv), a.grid, fp(rh), fp(rw),
                                           rh.shape[0] if rh is not None else 0, fp(out))
    check(st, "dsocr_test_vision_attention")
    print("ok", out.shape, float(np.abs(out).mean()))


if __name__ == "__main__":
    main()
