#!/usr/bin/env python
"""BASELINE.json configs[4]: expert-parallel MoE decoder on the GPUs of one box, measured separately from the default
(data-parallel) benchmark.  One process, one engine per GPU (dsocr/dispatch.py EnginePool); every engine decodes its own
`--pages` pages with a long output; the same job runs data-parallel (every GPU streams all populated experts) and
expert-parallel (rank r computes experts [r*E/n, (r+1)*E/n) for everybody's tokens; rows move by peer stores / loads over
NVLink with flag barriers, no NCCL on the data path).  Reports decode tok/s (whole box) and ms per step for both.
  python scripts/bench_ep.py --gpus 8 --pages 128 --tokens 1024
`scripts/bench_ep_nccl.py` (under torchrun) times a plain NCCL all-to-all of the same dispatch + combine payload per layer:
the baseline the fused exchange is compared with."""
from __future__ import annotations

import argparse
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--pages", type=int, default=128, help="pages per GPU")
    ap.add_argument("--tokens", type=int, default=1024, help="output tokens per page")
    ap.add_argument("--prompt-image-tokens", type=int, default=903)
    ap.add_argument("--config", default="full", choices=["full", "tiny"])
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--same-device", action="store_true", help="all engines on GPU 0 (functional check only)")
    args = ap.parse_args()
    import bench as B
    from dsocr.dispatch import EnginePool
    from dsocr.engine import DecodeParameters

    class A:
        pass
    a = A(); a.config = args.config; a.dtype = args.dtype
    cfg, ckdir = B.ensure_checkpoint(a, 0)
    devices = [0] * args.gpus if args.same_device else list(range(args.gpus))
    pool = EnginePool.load(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), None, devices, args.dtype,
                           configure=lambda e: e.set_option("kv_cache_f16", 1))
    n_img = args.prompt_image_tokens
    tail = B.prompt_tail(cfg)
    ids = [0] + [cfg.image_token_id] * n_img + tail
    mask = [0] + [1] * n_img + [0] * len(tail)
    rng = np.random.default_rng(0)
    rows = [[(rng.standard_normal((n_img, cfg.hidden_size)) * 0.7).astype(np.float32) for _ in range(args.pages)] for _ in devices]
    params = DecodeParameters(max_new_tokens=args.tokens, no_repeat_ngram_size=20, eos_token_id=None)
    warm = DecodeParameters(max_new_tokens=8, no_repeat_ngram_size=20, eos_token_id=None)

    def run_all(p):
        out, tm = [None] * len(devices), [None] * len(devices)

        def work(r):
            e = pool.engines[r]
            out[r] = e.generate_batch([ids] * args.pages, [mask] * args.pages, rows[r], p)
            tm[r] = e.timings()
        th = [threading.Thread(target=work, args=(r,)) for r in range(len(devices))]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return out, tm, time.perf_counter() - t0

    res = {}
    for mode in ("data_parallel", "expert_parallel"):
        if mode == "expert_parallel":
            pool.enable_expert_parallel(max_pages_per_engine=args.pages)
        run_all(warm)
        out, tm, wall = run_all(params)
        it_ms = max(t["decode.iterative"] for t in tm)   # device time of the token loop, max over the GPUs
        toks = sum(len(o) for r in out for o in r)
        res[mode] = {"decode_tok_s_box": (toks - len(devices) * args.pages) / (it_ms * 1e-3), "ms_per_step": it_ms / (args.tokens - 1),
                     "prefill_ms": max(t["decode.prefill"] for t in tm), "wall_s": wall, "tokens": toks,
                     "first_tokens": out[0][0][:8]}
    pool.disable_expert_parallel()
    pool.close()
    same = res["data_parallel"]["first_tokens"] == res["expert_parallel"]["first_tokens"]
    print(json.dumps({"config": f"deepseek-ocr {args.dtype} decoder, {args.gpus} GPU(s), {args.pages} pages per GPU, prompt {len(ids)} tokens, "
                                f"{args.tokens}-token outputs, random-init weights ({args.config} architecture), f16 KV",
                      "gpus": args.gpus, "same_device": args.same_device, **res, "tokens_equal_first_page": same,
                      "speedup_ep_over_dp": res["expert_parallel"]["decode_tok_s_box"] / res["data_parallel"]["decode_tok_s_box"]}))


if __name__ == "__main__":
    main()
