#!/usr/bin/env python
"""Throughput benchmark of the DeepSeek-OCR per-page forward path on B200 (see BASELINE.json).

One "step" = one pass of the whole hot path over the page set:
   host RGB8 pages -> H2D -> integer resample / tiling -> SAM + CLIP + projector -> prompt build -> prefill ->
   greedy decode (no-repeat-ngram 20) to the token budget -> D2H of the generated ids.
Default workload = BASELINE.json configs[2], the configuration the metric is quoted on: bf16, Gundam dynamic tiling
(A4 pages 1654x2339 -> 6 local 640x640 crops + the 1024x1024 global view, 903 image tokens), ONE fixed set of 1024
synthetic pages sharded round-robin over the N GPUs (strong scaling, no data-path collective), 512-token budget,
random-init weights of the exact architecture (no checkpoint offline).  `--mode base` = configs[1] (Base 1024x1024).

  value : pages/s with the (already resized / tiled) page views resident in HBM when the timed region starts
  e2e   : pages/s through the public C-ABI call `dsocr_decode_pages` with pinned HOST page buffers (H2D of the
          pages and D2H of the tokens inside the timed region)
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the CPU oracle (the restatement of the reference's algorithm; the reference itself
cannot be built here: no cargo, candle not vendored) on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

import numpy as np  # noqa: E402

PROMPT_TAIL = [201, 1719, 32, 9120, 270, 4482, 304, 39353, 16, 185]  # stand-in ids for the text after <image>


def prompt_tail(cfg):
    return [t % (cfg.vocab_size - 2) + 2 for t in PROMPT_TAIL] if cfg.vocab_size < 129280 else PROMPT_TAIL


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pages", type=int, default=0, help="pages in the whole job (default: 1024 Gundam / 64 Base), sharded over the GPUs")
    ap.add_argument("--batch", type=int, default=1024, help="pages decoded in lock-step per group on one GPU")
    ap.add_argument("--max-new-tokens", type=int, default=512)
    ap.add_argument("--mode", default="gundam", choices=["base", "gundam"])
    ap.add_argument("--dtype", default="", choices=["", "f16", "bf16"], help="default: bf16 Gundam / f16 Base (BASELINE.json)")
    ap.add_argument("--config", default="full", choices=["full", "tiny"])
    ap.add_argument("--cpu-tokens", type=int, default=64, help="decode tokens in the CPU-baseline sample")
    ap.add_argument("--agree-pages", type=int, default=4, help="pages of the GPU batch compared token by token with the CPU oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the f32-KV / batch-1 / CUPTI / DSQ side measurements")
    ap.add_argument("--dsq-formats", default="q4k", help="comma list of DSQ snapshot formats for the configs[3] side line (q4k, q8_0, q6k; empty = none)")
    ap.add_argument("--dsq-pages", type=int, default=128, help="Gundam pages of the DSQ pages/s measurement")
    ap.add_argument("--dsq-tokens", type=int, default=4096, help="output length of the DSQ batch-1 decode measurement (configs[3]: 4096)")
    ap.add_argument("--profile-json", default="", help="write the per-kernel timing breakdown here")
    ap.add_argument("--kv-cache", default="f16", choices=["f32", "f16"],
                    help="KV cache storage of the headline run: f16 (half the decode-attention bytes; token agreement with the f32 "
                         "oracle is reported) or f32 (the reference's choice; its number is printed beside the headline)")
    a = ap.parse_args()
    if not a.dtype:
        a.dtype = "bf16" if a.mode == "gundam" else "f16"
    if a.pages <= 0:
        a.pages = 1024 if a.mode == "gundam" else 64
    return a


def checkpoint_dir(args) -> Path:
    return Path(os.environ.get("DSOCR_BENCH_DIR", "/tmp")) / f"dsocr_bench_{args.config}_{args.dtype}"


def ensure_checkpoint(args, rank: int):
    """Random-init checkpoint of the exact architecture, written once per box (rank 0)."""
    import torch
    from oracle import config as OC

    d = checkpoint_dir(args)
    cfg = OC.full_config() if args.config == "full" else OC.tiny_config()
    done = d / "DONE"
    if rank == 0 and not done.exists():
        d.mkdir(parents=True, exist_ok=True)
        t0 = time.time()
        ck = OC.random_checkpoint(cfg, seed=1234, storage=torch.float16 if args.dtype == "f16" else torch.bfloat16)
        OC.save_checkpoint(ck, str(d / "model.safetensors"))
        cfg.save_json(str(d / "config.json"))
        done.write_text("ok")
        del ck
        print(f"[bench] wrote random-init checkpoint to {d} in {time.time() - t0:.1f}s", file=sys.stderr)
    return cfg, d


def _page(job):
    from oracle import preprocess as P

    w, h, seed = job
    return P.synthetic_page(w, h, seed=seed)


def make_pages(args, indices, workers: int):
    """Synthetic document pages (SURVEY.md 8d generator), page i seeded with i.  Generated on forked worker processes
    BEFORE anything touches CUDA (70 ms per A4 page on one core)."""
    w, h = (1024, 1024) if args.mode == "base" else (1654, 2339)
    jobs = [(w, h, int(i)) for i in indices]
    if workers <= 1 or len(jobs) < 16:
        return [_page(j) for j in jobs]
    import multiprocessing as mp

    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_page, jobs, chunksize=8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_sample(args, cfg, ckdir: Path, page: np.ndarray, max_new: int, n_tok: int, oracles=None) -> dict:
    """The oracle (CPU restatement of the reference algorithm, f32, torch/MKL threads = all host cores) on ONE
    page: preprocess + vision + prefill + `n_tok` decode steps; the token loop is extrapolated linearly to
    the full budget (per-step cost is dominated by the fixed 574 M active parameters)."""
    import torch
    from oracle import config as OC, decoder as D, preprocess as P, vision as V

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if oracles is None:
        ck = OC.load_checkpoint(str(ckdir / "model.safetensors"))
        oracles = (V.VisionOracle(cfg, ck), D.DecoderOracle(cfg, ck))
    vo, do = oracles
    t = {}
    t0 = time.perf_counter()
    base, img = (1024, 1024) if args.mode == "base" else (1024, 640)
    vi = P.prepare_vision_input(page, base, img, args.mode == "gundam")
    g = torch.from_numpy(P.image_to_tensor(vi["global"]))
    tiles = torch.from_numpy(np.stack([P.image_to_tensor(x) for x in vi["tiles"]])) if vi["tiles"] else None
    t["vision.prepare_inputs"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        rows = vo.encode(g, tiles, vi["crop_shape"])
    t["vision.compute_embeddings"] = time.perf_counter() - t0
    ids, mask = D.build_prompt_tokens([[], prompt_tail(cfg)], [rows.shape[0]], cfg)
    n_tok = min(n_tok, max_new)
    marks = []
    t0 = time.perf_counter()
    with torch.no_grad():
        toks = do.generate(ids, mask, rows, n_tok, 20, None, callback=lambda c, tk: marks.append(time.perf_counter()))
    total = time.perf_counter() - t0
    t["decode.prefill"] = marks[0] - t0
    per_tok = (marks[-1] - marks[0]) / max(1, len(marks) - 1)
    t["decode.iterative"] = per_tok * (max_new - 1)
    page_s = sum(t.values())
    return {"pages_per_s": 1.0 / page_s, "seconds_per_page": page_s, "stages_s": t, "cores": cores,
            "cpu_decode_tok_s": 1.0 / per_tok, "sample_tokens": n_tok,
            "sample_wall_s": total + t["vision.compute_embeddings"] + t["vision.prepare_inputs"],
            "tokens": toks, "oracles": oracles}


# ------------------------------------------------------------------------------------------- DSQ side line
def dsq_side_line(args, cfg, ckdir: Path, fmt: str, pages, vs, device: int) -> dict:
    from dsocr.engine import DecodeParameters, load_model
    from dsocr.export import export_snapshot

    snap = ckdir / f"model.{fmt}.lib.dsq"
    t_export = None
    if not snap.exists():
        t0 = time.time()
        export_snapshot(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), str(snap), fmt)
        t_export = time.time() - t0
    eng = load_model(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), str(snap), device, args.dtype)
    eng.set_option("kv_cache_f16", 1 if args.kv_cache == "f16" else 0)
    eng.set_option("decode_batch", args.batch)
    params = DecodeParameters(max_new_tokens=args.max_new_tokens, no_repeat_ngram_size=20, eos_token_id=None)
    out = {"snapshot_GB": snap.stat().st_size / 1e9, "export_s": t_export}
    # pages/s through the public call on a bounded page set (host pages in, tokens out)
    eng.decode_pages(pages[:8], vs, [], prompt_tail(cfg), cfg.image_token_id, DecodeParameters(max_new_tokens=8, eos_token_id=None))
    t0 = time.time()
    res = eng.decode_pages(pages, vs, [], prompt_tail(cfg), cfg.image_token_id, params)
    wall = time.time() - t0
    tm = eng.timings()
    out.update({"pages": len(pages), "e2e_pages_per_s": len(pages) / wall, "stage_ms": tm,
                "prefill_rows": sum(r.prompt_tokens for r in res),
                "prefill_tok_s": sum(r.prompt_tokens for r in res) / max(1e-9, tm["decode.prefill"] * 1e-3),
                "decode_tok_s": sum(r.response_tokens for r in res) / max(1e-9, tm["decode.iterative"] * 1e-3)})
    # batch-1 decode, long output (the reference's own `generate` shape)
    n_img = 903 if args.mode == "gundam" else 273
    ids1 = [0] + [cfg.image_token_id] * n_img + prompt_tail(cfg)
    mask1 = [0] + [1] * n_img + [0] * len(prompt_tail(cfg))
    rows1 = (np.random.default_rng(0).standard_normal((n_img, cfg.hidden_size)) * 0.7).astype(np.float32)
    eng.generate_batch([ids1], [mask1], [rows1], DecodeParameters(max_new_tokens=8, no_repeat_ngram_size=20, eos_token_id=None))
    o1 = eng.generate_batch([ids1], [mask1], [rows1], DecodeParameters(max_new_tokens=args.dsq_tokens, no_repeat_ngram_size=20, eos_token_id=None))
    it_ms = eng.timings()["decode.iterative"]
    n1 = len(o1[0])
    per_tok = {"q4k": 449e6, "q8_0": 610e6}.get(fmt)  # algorithmic weight bytes per token
    kvb = 2 if args.kv_cache == "f16" else 4
    kv_per_tok = 2 * cfg.hidden_size * cfg.num_layers * kvb * (len(ids1) + n1 / 2.0)
    tok_s = (n1 - 1) / max(1e-9, it_ms * 1e-3)
    out["decode_batch1"] = {"tokens": n1, "prompt_tokens": len(ids1), "tok_s": tok_s, "ms_per_token": it_ms / max(1, n1 - 1),
                            "hbm_GBps_weights_plus_kv": (per_tok + kv_per_tok) * tok_s / 1e9 if per_tok else None}
    eng.close()
    return out


# ------------------------------------------------------------------------------------------- roofline
def kernel_model(cfg, name: str, args, B: int, views: dict, ctx_mean: float, hbm_gbs: float, tf_peak: float,
                 active_experts: float | None = None, launches: int = 1) -> dict | None:
    """Algorithmic bytes / flops of an AVERAGE launch of a named kernel (DESIGN.md 'Kernels').
    B = pages per decode step; views = {'global': (n, T), 'local': (n, T)} of the profiled pass."""
    H, V, E, mi, K = cfg.hidden_size, cfg.vocab_size, cfg.n_routed_experts, cfg.moe_intermediate_size, cfg.num_experts_per_tok
    S = mi * cfg.n_shared_experts
    # routed-expert weight segments one launch really streams: measured (dsocr_moe_stats) when available
    Ea = active_experts if active_experts else min(E, B * K)
    phase, _, k = name.partition("/")
    if phase == "decode":
        act = lambda kk, nn, parts=2: B * kk * 2 * parts + B * nn * 4  # noqa: E731
        kvb = 2 if args.kv_cache == "f16" else 4
        table = {
            "lm_head": V * H * 2 + act(H, V),
            "dec_qkv": 3 * H * H * 2 + act(H, 3 * H),
            "dec_o_proj": H * H * 2 + act(H, H),
            "dec_dense_gate_up": 2 * cfg.intermediate_size * H * 2 + act(H, cfg.intermediate_size),
            "dec_dense_down": H * cfg.intermediate_size * 2 + act(cfg.intermediate_size, H),
            # routed experts actually hit + the shared experts, which run as extra groups of the same grouped GEMM
            "moe_expert_gate_up": (Ea + cfg.n_shared_experts) * 2 * mi * H * 2 + B * (K + cfg.n_shared_experts) * (H * 4 + mi * 4),
            "moe_expert_down": (Ea + cfg.n_shared_experts) * H * mi * 2 + B * (K + cfg.n_shared_experts) * (mi * 4 + H * 4),
            "moe_shared_gate_up": 2 * S * H * 2 + act(H, S),
            "moe_shared_down": H * S * 2 + act(S, H),
            # cached K and V rows of every (page, head) read once + the step's q/k/v row and the output row
            "rope_attn_decode": B * (2 * ctx_mean * H * kvb + 3 * H * 4 + 2 * H * 2),
        }
        if k in table:
            return {"bound": "hbm", "bytes": float(table[k]), "peak": hbm_gbs, "unit": "GB/s"}
    if phase == "vision":
        (ng, Tg), (nl, Tl) = views["global"], views["local"]
        T = ng * Tg + nl * Tl                      # SAM tokens of the pass
        Tc = ng * (Tg // 16 + 1) + nl * (Tl // 16 + 1)   # CLIP tokens (incl. cls)
        glob_layers = 4
        total = {
            "sam_qkv": 12 * 2 * T * 768 * 2304, "sam_proj": 12 * 2 * T * 768 * 768, "sam_fc1_gelu": 12 * 2 * T * 768 * 3072,
            "sam_fc2": 12 * 2 * T * 3072 * 768, "sam_global_attention": glob_layers * 4 * (ng * Tg * Tg + nl * Tl * Tl) * 64 * 12,
            "clip_fc1_quickgelu": 24 * 2 * Tc * 1024 * 4096, "clip_fc2": 24 * 2 * Tc * 1024 * 4096, "clip_qkv": 24 * 2 * Tc * 1024 * 3072,
        }
        if k in total:
            return {"bound": "tensor", "flops": float(total[k]) / max(1, launches), "peak": tf_peak, "unit": "TFLOP/s"}
    return None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    base, img, crop = (1024, 1024, False) if args.mode == "base" else (1024, 640, True)
    workload = (f"deepseek-ocr {args.dtype} {'Base 1024x1024' if args.mode == 'base' else 'Gundam A4 1654x2339 (6 x 640 crops + 1024 global)'}, "
                f"{args.pages} synthetic pages sharded over the GPUs, {args.max_new_tokens}-token budget, greedy, "
                f"no_repeat_ngram_size=20, {args.kv_cache} KV cache, random-init weights ({args.config} architecture)")

    # the workload description both arms print (the reference arm runs "on this arm's config")
    pages_per_rank = (args.pages + world - 1) // world
    run_config = {"workload": workload, "l2": "per-step working set (weights 6.7 GB + KV + activations) exceeds the 126 MB L2",
                  "parallelism": f"one fixed set of {args.pages} pages sharded round-robin over {world} GPU(s), "
                                 f"{min(args.batch, pages_per_rank)} pages per lock-step decode group, no collective"}

    if args.impl == "reference":
        # CPU restatement of the reference's algorithm on the host cores; rank 0 only.
        if rank != 0:
            return
        cfg, ckdir = ensure_checkpoint(args, 0)
        page = make_pages(args, [0], 1)[0]
        vals, oracles = [], None
        for _ in range(max(1, min(args.steps, 2))):
            r = cpu_reference_sample(args, cfg, ckdir, page, args.max_new_tokens, min(args.cpu_tokens, 32), oracles)
            oracles = r.pop("oracles")
            vals.append(r)
        best = max(vals, key=lambda r: r["pages_per_s"])
        sample = (f"1 page through the f32 torch-CPU oracle: preprocess + vision + prefill + {best['sample_tokens']} decode "
                  f"steps, token loop extrapolated linearly to {args.max_new_tokens} tokens")
        line = {
            "impl": "reference", "metric": "pages/sec/box", "value": best["pages_per_s"], "unit": "pages/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": best["seconds_per_page"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": run_config,
            "cpu_baseline": {"value": best["pages_per_s"], "unit": "pages/s", "cores": best["cores"], "kind": "port",
                             "sample": sample, "stages_s": best["stages_s"], "decode_tok_s": best["cpu_decode_tok_s"]},
            "e2e": {"value": best["pages_per_s"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ---- this rank's shard of the fixed page set (pages are independent: round-robin, no data-path collective)
    from dsocr.sharding import shard_indices

    mine = shard_indices(args.pages, rank, world)
    t0 = time.time()
    pages = make_pages(args, mine, max(1, min(16, (os.cpu_count() or 1) // max(1, world))))
    if rank == 0:
        print(f"[bench] {len(pages)} pages per rank generated in {time.time() - t0:.1f}s", file=sys.stderr)

    import torch
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    import __graft_entry__ as ge

    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    cfg, ckdir = ensure_checkpoint(args, rank)
    if world > 1:
        dist.barrier()
    from dsocr.engine import DecodeParameters, VisionSettings, load_model

    eng = load_model(str(ckdir / "config.json"), str(ckdir / "model.safetensors"), None, local_rank, args.dtype)
    stream = torch.cuda.Stream()  # a capturable (non-default) stream that torch events can bracket
    eng.set_stream(stream.cuda_stream)
    eng.set_option("kv_cache_f16", 1 if args.kv_cache == "f16" else 0)
    eng.set_option("decode_batch", args.batch)
    vs = VisionSettings(base, img, crop)
    params = DecodeParameters(max_new_tokens=args.max_new_tokens, no_repeat_ngram_size=20, eos_token_id=None)
    # the e2e arm copies every step's pages host -> device inside the timed region: keep them in pinned host memory
    # (page-locked sources DMA directly; pageable ones are staged through the driver at a third of the rate)
    pinned = [torch.from_numpy(p).pin_memory() for p in pages]
    pages = [t.numpy() for t in pinned]
    h2d_bytes = sum(int(p.nbytes) for p in pages)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        outs = None
        for _ in range(steps):
            outs = fn()
        e1.record(stream)
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), outs

    def step_e2e():
        return eng.decode_pages(pages, vs, [], prompt_tail(cfg), cfg.image_token_id, params)

    def step_resident():
        return eng.decode_staged([], prompt_tail(cfg), cfg.image_token_id, params)

    for _ in range(args.warmup):
        step_e2e()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_e2e, outs = timed(step_e2e, args.steps)
    launches_e2e = eng.launch_count() - launches0
    eng.stage_pages(pages, vs)
    ms_res, outs_res = timed(step_resident, args.steps)
    clocks = sampler.stop()
    stage_ms = eng.timings()
    gen_tokens = sum(o.response_tokens for o in outs_res)
    d2h_bytes = gen_tokens * 8

    # ---- side measurements on ONE lock-step group of this rank's pages (not part of the timed value)
    sub = pages[: min(len(pages), args.batch)]
    nsub = len(sub)
    kv_compare = None
    if not args.no_extras:
        # f32 KV needs 176 MB per Gundam page at a 512-token budget: at most 512 pages, whatever the headline group size
        nkv = min(nsub, 512)
        eng.stage_pages(sub[:nkv], vs)
        res = {}
        for kv in ("f16", "f32"):
            eng.set_option("kv_cache_f16", 1 if kv == "f16" else 0)
            step_resident()  # sizes the KV workspaces of this mode
            ms, _ = timed(step_resident, 1)
            res[kv] = {"pages_per_s_per_gpu": nkv / (ms * 1e-3), "ms": ms, "decode_ms": eng.timings()["decode.iterative"]}
        eng.set_option("kv_cache_f16", 1 if args.kv_cache == "f16" else 0)
        kv_compare = {"pages": nkv, "note": "one lock-step group on one GPU, views resident; the reference stores KV in f32 (model/mod.rs:82-88)", **res}

    # one extra pass with per-kernel CUDA-event timing for the roofline / breakdown
    eng.stage_pages(sub, vs)
    eng.set_option("moe_stats", 1)
    eng.moe_stats()
    eng.kernel_timing_begin()
    outs_sub = step_resident()
    kt = eng.kernel_timing_end()
    seg, nsteps = eng.moe_stats()
    eng.set_option("moe_stats", 0)
    moe_layers = sum(1 for l in range(cfg.num_layers) if l >= cfg.first_k_dense_replace)
    active_experts = seg / max(1.0, nsteps * moe_layers) if nsteps else None
    kt.sort(key=lambda r: -r["ms"])
    total_kernel_ms = sum(r["ms"] for r in kt)
    if args.profile_json and rank == 0:
        Path(args.profile_json).write_text(json.dumps({"kernels": kt, "total_ms": total_kernel_ms, "stage_ms": eng.timings(),
                                                       "pages": nsub}, indent=1))
    prompt_len = outs_sub[0].prompt_tokens
    gen_mean = float(np.mean([o.response_tokens for o in outs_sub]))
    ctx_mean = prompt_len + max(0.0, gen_mean - 1) / 2.0
    n_tiles = 0 if args.mode == "base" else 6
    views = {"global": (nsub, 4096), "local": (nsub * n_tiles, 1600)}

    # the same pass once more as it really runs (CUDA-graph replay), kernel durations from CUPTI activity records:
    # the per-launch events above serialise the step and add the host launch gap to short kernels
    cupti = {}
    if not args.no_extras:
        try:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step_resident()
                torch.cuda.synchronize()
            for e in prof.events():
                if e.device_type == torch.autograd.DeviceType.CUDA:
                    a = cupti.setdefault(e.name, [0.0, 0])
                    a[0] += e.time_range.end - e.time_range.start
                    a[1] += 1
        except Exception as ex:  # diagnostics only
            print(f"[bench] CUPTI pass skipped: {ex}", file=sys.stderr)
    if cupti and args.profile_json and rank == 0:
        # kernel durations inside the production CUDA-graph replay, by (shortened) kernel name, next to the event-timed list
        try:
            pj = json.loads(Path(args.profile_json).read_text())
            ck = sorted(((k.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").split("(")[0][-110:], v[0] * 1e-3, v[1]) for k, v in cupti.items()), key=lambda t: -t[1])
            pj["cupti_graph_kernels"] = [{"name": n, "ms": round(ms, 3), "launches": c, "avg_us": round(ms * 1e3 / max(1, c), 2)} for n, ms, c in ck[:60]]
            Path(args.profile_json).write_text(json.dumps(pj, indent=1))
        except Exception as ex:  # diagnostics only
            print(f"[bench] CUPTI list not written: {ex}", file=sys.stderr)
    cupti_match = {"decode/moe_expert_gate_up": ("linear_sk_kernel", ", 2>"), "decode/moe_expert_down": ("linear_sk_kernel", ", 1>"),
                   "decode/rope_attn_decode": ("rope_attn_decode", "")}
    traffic_db = {}
    try:
        traffic_db = json.loads((ROOT / "profiles" / "roofline_traffic.json").read_text())
    except Exception:
        pass

    roof, roofs = None, []
    for r in kt:
        m = kernel_model(cfg, r["name"], args, nsub, views, ctx_mean, hbm_peak, tf_sustained, active_experts, r["launches"])
        if m is None:
            continue
        avg_s = r["ms"] / r["launches"] * 1e-3
        ach = (m["bytes"] / avg_s / 1e9) if m["bound"] == "hbm" else (m["flops"] / avg_s / 1e12)
        tr = traffic_db.get(f"{args.mode}/{r['name']}") or traffic_db.get(r["name"])
        entry = {"kernel": r["name"], "bound": m["bound"], "achieved": ach, "peak": m["peak"], "unit": m["unit"],
                 "frac": ach / m["peak"], "traffic": tr.get("dram_bytes_per_launch") if tr else None,
                 "traffic_source": tr.get("source") if tr else None,
                 # the captured launch is not an average launch of this run: its own algorithmic bytes and the ratio
                 "traffic_over_algorithmic_of_captured_launch": (tr["dram_bytes_per_launch"] / tr["algorithmic_bytes_of_that_launch"])
                 if tr and tr.get("algorithmic_bytes_of_that_launch") else None,
                 "algorithmic_per_launch": m.get("bytes", m.get("flops")), "launches_per_pass": r["launches"],
                 "avg_launch_us": avg_s * 1e6, "share_of_pass": r["ms"] / total_kernel_ms}
        if r["name"] in cupti_match and cupti:
            a, b = cupti_match[r["name"]]
            us = [v for k, v in cupti.items() if a in k and (not b or k.split("(")[0].rstrip().endswith(b))]
            if us:
                avg_us = sum(v[0] for v in us) / max(1, sum(v[1] for v in us))
                ach_g = (m["bytes"] if m["bound"] == "hbm" else m["flops"] * 1e-3) / (avg_us * 1e-6) / 1e9
                entry.update({"achieved_graph": ach_g, "frac_graph": ach_g / m["peak"], "avg_kernel_us_graph": avg_us})
        roofs.append(entry)
    if roofs:
        roof = dict(roofs[0])  # the dominant kernel (largest share of the pass) that has a closed-form model
        roof.update({"active_experts_per_layer": active_experts, "mean_context": ctx_mean, "pages_per_decode_step": nsub,
                     "peak_source": peak_src,
                     "timed": "CUDA events after every launch on the engine stream, one extra profiled pass over one lock-step group; "
                              "achieved_graph = CUPTI kernel records of the same pass replayed as the production CUDA graph"})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pages_total = args.pages * args.steps
    value = pages_total / (ms_res * 1e-3)
    e2e_val = pages_total / (ms_e2e * 1e-3)
    line = {
        "metric": "pages/sec/box", "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": run_config,
        "e2e": {"value": e2e_val, "unit": "pages/s", "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": d2h_bytes * world,
                "host_memory": "pinned", "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_e2e,
        "clocks": clocks,
        "decode_tok_s_per_gpu": gen_tokens / max(1e-9, stage_ms["decode.iterative"] * 1e-3),
        "stage_ms": stage_ms,
        "kv_cache_compare": kv_compare,
        "top_kernels": [{"name": r["name"], "ms": round(r["ms"], 3), "launches": r["launches"]} for r in kt[:14]],
        "roofline": roof,
        "roofline_others": [{k: e[k] for k in ("kernel", "bound", "achieved", "frac", "share_of_pass", "avg_launch_us")} for e in roofs[1:8]],
    }
    if world == 1 and not args.no_extras:
        # BASELINE's second metric (decode tok/s per GPU) in the reference's own shape: one page per call.  Outside the
        # timed region; steps of <= 4 pages run the fused small-batch decode step (csrc/dsq_decode.cu).
        try:
            n_img = 903 if args.mode == "gundam" else 273
            ids1 = [0] + [cfg.image_token_id] * n_img + prompt_tail(cfg)
            mask1 = [0] + [1] * n_img + [0] * len(prompt_tail(cfg))
            rows1 = (np.random.default_rng(0).standard_normal((n_img, cfg.hidden_size)) * 0.7).astype(np.float32)
            n1 = max(2, min(256, args.max_new_tokens))
            eng.generate_batch([ids1], [mask1], [rows1], DecodeParameters(max_new_tokens=8, no_repeat_ngram_size=20, eos_token_id=None))
            out1 = eng.generate_batch([ids1], [mask1], [rows1], DecodeParameters(max_new_tokens=n1, no_repeat_ngram_size=20, eos_token_id=None))
            it_ms = eng.timings()["decode.iterative"]
            line["decode_batch1"] = {"tok_s": (len(out1[0]) - 1) / max(1e-9, it_ms * 1e-3), "ms_per_token": it_ms / max(1, len(out1[0]) - 1),
                                     "tokens": len(out1[0]), "prompt_tokens": len(ids1),
                                     "path": "fused small-batch decode step (6 launches per layer), CUDA graph + PDL"}
        except Exception as ex:  # diagnostics only: never take the headline number down
            line["decode_batch1"] = {"error": str(ex)}
    if world == 1 and not args.no_extras and args.dsq_formats and args.config == "full":
        # BASELINE configs[3]: the same engine over a DSQ snapshot (exported here with the library's own writer from the
        # random-init checkpoint, dtype assignment of the reference exporter).  Prefill and multi-page decode run the
        # dequant-fused tensor-core GEMM (linear_dq.cuh), batch-1 decode the fused GEMV step (dsq_decode.cu).
        eng.close()
        line["dsq"] = {}
        for fmt in [f for f in args.dsq_formats.split(",") if f]:
            try:
                line["dsq"][fmt] = dsq_side_line(args, cfg, ckdir, fmt, pages[: min(len(pages), args.dsq_pages)], vs, local_rank)
            except Exception as ex:  # diagnostics only
                line["dsq"][fmt] = {"error": str(ex)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_sample(args, cfg, ckdir, pages[0], args.max_new_tokens, args.cpu_tokens)
            oracles = r.pop("oracles")
            line["cpu_baseline"] = {
                "value": r["pages_per_s"], "unit": "pages/s", "cores": r["cores"], "kind": "port",
                "sample": (f"1 page through the f32 torch-CPU oracle: preprocess + vision + prefill + {r['sample_tokens']} "
                           f"decode steps ({r['sample_wall_s']:.1f} s of CPU work), token loop extrapolated linearly to "
                           f"{args.max_new_tokens} tokens"),
                "stages_s": r["stages_s"], "decode_tok_s": r["cpu_decode_tok_s"]}
            # token agreement of pages of the timed GPU run (end-to-end call, the headline KV mode) with the f32 oracle
            agree, total, per_page = 0, 0, []
            refs = [r["tokens"]]
            for i in range(1, min(args.agree_pages, len(pages))):
                refs.append(cpu_reference_sample(args, cfg, ckdir, pages[i], args.max_new_tokens, args.cpu_tokens, oracles)["tokens"])
            for i, ref in enumerate(refs):
                got = outs[i].generated_tokens[: len(ref)]
                a = sum(int(x == y) for x, y in zip(got, ref))
                per_page.append(a / max(1, len(ref)))
                agree += a; total += len(ref)
            line["token_agreement"] = {"pages": len(refs), "tokens_per_page": len(refs[0]), "agreement": agree / max(1, total),
                                       "per_page": per_page, "against": "f32 CPU oracle, same pages, free-running greedy",
                                       "gpu_run": f"timed end-to-end call, {args.kv_cache} KV cache"}
        except Exception as ex:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "pages/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {ex}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
