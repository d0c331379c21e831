#!/usr/bin/env python
"""Runs the CTA-pair GEMM (csrc/linear_pair.cuh) once through its test hook at a GPU-filling size, one epilogue variant per
call, so that `ncu --set full -k regex:linear_pair_kernel -c 1` has something representative to capture without loading a
model:  python scripts/pair_probe.py {fc1|fc2|proj|qkv|relpos}   (SAM shapes, M = 32768 token rows)"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))
from tests.test_linear_gpu import BF16, run_linear  # noqa: E402

SHAPES = {  # name: (N, K, act, out_mode, row_map)
    "fc1": (3072, 768, 1, 0, False),     # GELU, 16-bit TMA-store epilogue
    "fc2": (768, 3072, 0, 3, False),     # residual add (bulk reductions)
    "proj": (768, 768, 0, 3, True),      # residual add through the window row map
    "qkv": (2304, 768, 0, 0, False),     # 16-bit output
    "relpos": (3072, 768, 0, 2, False),  # f32 rows (bulk stores)
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "proj"
    N, K, act, out_mode, mapped = SHAPES[name]
    M = 32768
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.03
    b = torch.randn(N, generator=g)
    kw = {}
    if out_mode == 3:
        kw["out_init"] = np.zeros((M, N), dtype=np.float32)
    if mapped:
        kw["row_map"] = np.random.RandomState(1).permutation(M).astype(np.int32)
        kw["out_rows"] = M
    y = run_linear(BF16, x, w, bias=b, act=act, out_mode=out_mode, **kw)
    print(name, "ok", tuple(y.shape), float(y.abs().mean()))


if __name__ == "__main__":
    main()
