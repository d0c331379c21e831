#!/usr/bin/env python
"""Per-kernel GPU time of one bench step as it really runs (CUDA-graph replay included), taken with CUPTI through
torch.profiler -- no per-launch events, no host gaps.  Writes a JSON summary: per-kernel totals, the idle time between
kernels, and the kernel timeline of a few consecutive decode steps.
Usage: python scripts/trace_step.py [--pages 64] [--max-new-tokens 128] [--out gpurun_out/trace.json]"""
from __future__ import annotations

import argparse
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

import torch  # noqa: E402


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|dsocr::|lin::|vattn::", "", name)
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0][:90]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=64)
    ap.add_argument("--max-new-tokens", type=int, default=128)
    ap.add_argument("--mode", default="base")
    ap.add_argument("--kv-cache", default="f16")
    ap.add_argument("--out", default="gpurun_out/trace.json")
    a = ap.parse_args()
    import bench as B

    sys.argv = [sys.argv[0]]
    args = B.parse_args()
    args.pages, args.max_new_tokens, args.mode, args.kv_cache = a.pages, a.max_new_tokens, a.mode, a.kv_cache
    cfg, ckdir = B.ensure_checkpoint(args, 0)
    from dsocr.engine import DecodeParameters, VisionSettings, load_model

    eng = load_model(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), None, 0, args.dtype)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.set_option("kv_cache_f16", 1 if a.kv_cache == "f16" else 0)
    base, img, crop = (1024, 1024, False) if a.mode == "base" else (1024, 640, True)
    vs = VisionSettings(base, img, crop)
    params = DecodeParameters(max_new_tokens=a.max_new_tokens, no_repeat_ngram_size=20, eos_token_id=None)
    pages = B.make_pages(args, 0)
    tail = B.prompt_tail(cfg)
    for _ in range(2):
        eng.decode_pages(pages, vs, [], tail, cfg.image_token_id, params)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.decode_pages(pages, vs, [], tail, cfg.image_token_id, params)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    tot = defaultdict(lambda: [0.0, 0])
    for s, e, n in ks:
        t = tot[short(n)]
        t[0] += e - s
        t[1] += 1
    span = ks[-1][1] - ks[0][0] if ks else 0.0
    busy = 0.0
    cur_end = None
    for s, e, _ in ks:  # union of intervals (two streams overlap)
        if cur_end is None or s > cur_end:
            busy += e - s
            cur_end = e
        elif e > cur_end:
            busy += e - cur_end
            cur_end = e
    table = sorted(({"kernel": k, "total_us": v[0], "launches": v[1], "avg_us": v[0] / v[1]} for k, v in tot.items()),
                   key=lambda r: -r["total_us"])
    # timeline of ~3 decode steps from the middle of the run: split at select_token launches
    sel = [i for i, (_, _, n) in enumerate(ks) if "select_token" in n]
    timeline = []
    if len(sel) > 8:
        mid = len(sel) // 2
        i0, i1 = sel[mid] + 1, sel[mid + 2] + 1
        t0 = ks[i0][0]
        timeline = [{"t_us": round(s - t0, 2), "dur_us": round(e - s, 2), "kernel": short(n)} for s, e, n in ks[i0:i1]]
    out = {"span_us": span, "busy_us": busy, "idle_us": span - busy, "n_kernels": len(ks), "stage_ms": eng.timings(),
           "kernels": table, "timeline_two_decode_steps": timeline}
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(out, indent=1))
    print(json.dumps({"span_ms": span / 1e3, "busy_ms": busy / 1e3, "n_kernels": len(ks)}))
    for r in table[:25]:
        print(f"{r['total_us'] / 1e3:9.2f} ms {r['launches']:6d} x {r['avg_us']:8.2f} us  {r['kernel']}")


if __name__ == "__main__":
    main()
