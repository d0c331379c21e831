#!/usr/bin/env python
"""BASELINE.json configs[3]: DSQ q4k / q8_0 dequant-fused decode.  Synthesises a snapshot of the random-init
full-size decoder (dtype assignment of the reference exporter), loads the engine with it and measures batch-1
(and small-batch) greedy decode tok/s plus the achieved HBM GB/s of the dequant GEMV kernels.
Usage: python scripts/bench_dsq.py [--primary q4k|q6k|q8_0] [--tokens 256] [--pages 1]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--primary", default="q4k", choices=["q4k", "q6k", "q8_0", "float"],
                    help="snapshot format, or `float`: the 16-bit engine without a snapshot (batch-1 decode of configs[1]/[2])")
    ap.add_argument("--tokens", type=int, default=256)
    ap.add_argument("--pages", type=int, default=1)
    ap.add_argument("--config", default="full", choices=["full", "tiny"])
    ap.add_argument("--dtype", default="bf16")
    args = ap.parse_args()
    import bench as B  # random-init checkpoint generator shared with bench.py
    from dsocr.export import export_snapshot

    class A: pass
    a = A(); a.config = args.config; a.dtype = args.dtype
    cfg, ckdir = B.ensure_checkpoint(a, 0)
    snap = ckdir / f"model.{args.primary}.lib.dsq"
    if args.primary == "float":
        snap = None
    elif not snap.exists():
        t0 = time.time()  # the library's own exporter (dsocr_dsq_writer_*): adapter tensor list + dtype fallback chain
        export_snapshot(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), str(snap), args.primary)
        print(f"[dsq] wrote {snap} ({snap.stat().st_size / 1e9:.2f} GB) in {time.time() - t0:.1f}s", file=sys.stderr)
    from dsocr.engine import DecodeParameters, load_model

    t0 = time.time()
    eng = load_model(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), None if snap is None else str(snap), 0, args.dtype)
    print(f"[dsq] engine load {time.time() - t0:.1f}s", file=sys.stderr)
    eng.set_option("kv_cache_f16", 1)
    g = torch.Generator().manual_seed(0)
    n_img = 273
    ids = [[0] + [cfg.image_token_id] * n_img + B.prompt_tail(cfg) for _ in range(args.pages)]
    masks = [[0] + [1] * n_img + [0] * len(B.prompt_tail(cfg)) for _ in range(args.pages)]
    rows = [(torch.randn(n_img, cfg.hidden_size, generator=g) * 0.7).numpy() for _ in range(args.pages)]
    params = DecodeParameters(max_new_tokens=args.tokens, eos_token_id=None)
    eng.generate_batch(ids, masks, rows, DecodeParameters(max_new_tokens=8, eos_token_id=None))  # warm-up
    t0 = time.time()
    l0 = eng.launch_count()
    out = eng.generate_batch(ids, masks, rows, params)
    wall = time.time() - t0
    launches = eng.launch_count() - l0
    tm = eng.timings()
    eng.kernel_timing_begin()
    eng.generate_batch(ids, masks, rows, DecodeParameters(max_new_tokens=33, eos_token_id=None))
    kt = [r for r in eng.kernel_timing_end() if r["name"].startswith("decode/")]
    kt.sort(key=lambda r: -r["ms"])
    H, V = cfg.hidden_size, cfg.vocab_size
    bytes_lm = V * H * (2 if args.primary == "float" else 34 / 32)
    lm = next((r for r in kt if r["name"].endswith("fs_lm_head")), None)
    tok_s = args.pages * (args.tokens - 1) / (tm["decode.iterative"] * 1e-3)
    # algorithmic weight bytes per token: q4k 449 MB, q8_0 610 MB at batch 1
    per_tok = {"q4k": 449e6, "q6k": None, "q8_0": 610e6, "float": 1148e6}[args.primary]
    # KV bytes read per token at the mean context of the run (f16 cache: K and V, 12 layers)
    mean_ctx = len(ids[0]) + args.tokens / 2
    kv_per_tok = 2 * cfg.hidden_size * cfg.num_layers * 2 * mean_ctx
    line = {"config": f"deepseek-ocr-{args.primary} ({args.dtype} engine) decode, batch {args.pages}, {args.tokens}-token output, random-init weights",
            "path": "batched kernels / per-linear GEMVs" if (os.environ.get("DSOCR_DSQ_UNFUSED") or os.environ.get("DSOCR_NO_SMALL_FUSED")) else "fused small-batch step (dsq_decode.cu)",
            "launches_per_token": round(launches / max(1, args.tokens), 1),
            "hbm_GBps_weights_plus_kv": ((per_tok + kv_per_tok) * tok_s / args.pages / 1e9) if per_tok else None,
            "decode_tok_s": tok_s, "ms_per_token": tm["decode.iterative"] / (args.tokens - 1), "wall_s": wall,
            "weight_GBps_at_batch1": (per_tok * tok_s / args.pages / 1e9) if per_tok else None,
            "lm_head_gemv": None if lm is None else {"avg_us": lm["ms"] / lm["launches"] * 1e3,
                                                      "GBps": bytes_lm / (lm["ms"] / lm["launches"] * 1e-3) / 1e9},
            "top_kernels": [{"name": r["name"], "avg_us": round(r["ms"] / r["launches"] * 1e3, 2), "launches": r["launches"]} for r in kt[:10]],
            "tokens_head": out[0][:8]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
