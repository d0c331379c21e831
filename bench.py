#!/usr/bin/env python
"""Throughput benchmark of the DeepSeek-OCR per-page forward path on B200 (see BASELINE.json).

One "step" = one pass of the whole hot path over a batch of synthetic pages:
   host RGB8 pages -> integer preprocess -> H2D -> SAM + CLIP + projector -> prompt build -> prefill ->
   greedy decode (no-repeat-ngram 20) to the token budget -> D2H of the generated ids.
Workload at N=1: BASELINE.json configs[1]: fp16, Base mode 1024x1024, batch of 64 synthetic pages, 512-token
budget, random-init weights of the exact architecture (no checkpoint offline).

  value : pages/s with the (already resized) page views resident in HBM when the timed region starts
  e2e   : pages/s through the public C-ABI call `dsocr_decode_pages` with HOST page buffers (H2D of the
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
    ap.add_argument("--pages", type=int, default=64, help="pages per GPU per step")
    ap.add_argument("--max-new-tokens", type=int, default=512)
    ap.add_argument("--mode", default="base", choices=["base", "gundam"])
    ap.add_argument("--dtype", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--config", default="full", choices=["full", "tiny"])
    ap.add_argument("--cpu-tokens", type=int, default=24, help="decode tokens in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default="", help="write the per-kernel timing breakdown here")
    ap.add_argument("--kv-cache", default="f16", choices=["f32", "f16"],
                    help="KV cache storage: f16 (matches the fp16 config; token-exact on the fixtures) or f32 (the reference's choice)")
    return ap.parse_args()


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


def make_pages(args, rank: int):
    from oracle import preprocess as P

    w, h = (1024, 1024) if args.mode == "base" else (1654, 2339)
    return [P.synthetic_page(w, h, seed=rank * 100000 + i) for i in range(args.pages)]


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
def cpu_reference_sample(args, cfg, ckdir: Path, page: np.ndarray, max_new: int) -> dict:
    """The oracle (CPU restatement of the reference algorithm, f32, torch/MKL threads = all host cores) on ONE
    page: preprocess + vision + prefill + `cpu_tokens` decode steps; the token loop is extrapolated linearly to
    the full budget (per-step cost is dominated by the fixed 574 M active parameters)."""
    import torch
    from oracle import config as OC, decoder as D, preprocess as P, vision as V

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ck = OC.load_checkpoint(str(ckdir / "model.safetensors"))
    t = {}
    t0 = time.perf_counter()
    base, img = (1024, 1024) if args.mode == "base" else (1024, 640)
    vi = P.prepare_vision_input(page, base, img, args.mode == "gundam")
    g = torch.from_numpy(P.image_to_tensor(vi["global"]))
    tiles = torch.from_numpy(np.stack([P.image_to_tensor(x) for x in vi["tiles"]])) if vi["tiles"] else None
    t["vision.prepare_inputs"] = time.perf_counter() - t0
    vo = V.VisionOracle(cfg, ck)
    t0 = time.perf_counter()
    with torch.no_grad():
        rows = vo.encode(g, tiles, vi["crop_shape"])
    t["vision.compute_embeddings"] = time.perf_counter() - t0
    do = D.DecoderOracle(cfg, ck)
    ids, mask = D.build_prompt_tokens([[], prompt_tail(cfg)], [rows.shape[0]], cfg)
    n_tok = min(args.cpu_tokens, max_new)
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
            "cpu_decode_tok_s": 1.0 / per_tok, "sample_tokens": n_tok, "sample_wall_s": total + t["vision.compute_embeddings"],
            "first_tokens": toks[:8]}


# ------------------------------------------------------------------------------------------- roofline
def kernel_model(cfg, name: str, args, pages: int, hbm_gbs: float, tf_peak: float, active_experts: float | None = None) -> dict | None:
    """Algorithmic bytes / flops of one launch of a named kernel (DESIGN.md 'Kernels')."""
    H, V, E, mi, K = cfg.hidden_size, cfg.vocab_size, cfg.n_routed_experts, cfg.moe_intermediate_size, cfg.num_experts_per_tok
    S = mi * cfg.n_shared_experts
    B = pages
    # routed-expert weight segments one launch really streams: measured (dsocr_moe_stats) when available
    Ea = active_experts if active_experts else min(E, B * K)
    phase, _, k = name.partition("/")
    if phase == "decode":
        act = lambda kk, nn, parts=2: B * kk * 2 * parts + B * nn * 4  # noqa: E731
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
        }
        if k in table:
            return {"bound": "hbm", "bytes": float(table[k]), "peak": hbm_gbs, "unit": "GB/s"}
    if phase == "vision":
        g = 64 if args.mode == "base" else 64
        T = g * g
        flops = {
            "sam_qkv": 2 * T * 768 * 2304, "sam_proj": 2 * T * 768 * 768, "sam_fc1_gelu": 2 * T * 768 * 3072,
            "sam_fc2": 2 * T * 3072 * 768, "sam_global_attention": 4 * T * T * 64 * 12,
            "clip_fc1_quickgelu": 2 * 257 * 1024 * 4096, "clip_fc2": 2 * 257 * 1024 * 4096, "clip_qkv": 2 * 257 * 1024 * 3072,
        }
        if k in flops:
            return {"bound": "tensor", "flops": float(flops[k]) * B, "peak": tf_peak, "unit": "TFLOP/s"}
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
    workload = (f"deepseek-ocr {args.dtype} {'Base 1024x1024' if args.mode == 'base' else 'Gundam A4 1654x2339'}, "
                f"batch of {args.pages} synthetic pages per GPU, {args.max_new_tokens}-token budget, greedy, "
                f"no_repeat_ngram_size=20, {args.kv_cache} KV cache, random-init weights ({args.config} architecture)")

    if args.impl == "reference":
        # CPU restatement of the reference's algorithm on the host cores; rank 0 only.
        if rank != 0:
            return
        cfg, ckdir = ensure_checkpoint(args, 0)
        page = make_pages(args, 0)[0]
        vals = []
        for _ in range(max(1, min(args.steps, 2))):
            r = cpu_reference_sample(args, cfg, ckdir, page, args.max_new_tokens)
            vals.append(r)
        best = max(vals, key=lambda r: r["pages_per_s"])
        sample = (f"1 page through the f32 torch-CPU oracle: preprocess + vision + prefill + {best['sample_tokens']} decode "
                  f"steps, token loop extrapolated linearly to {args.max_new_tokens} tokens")
        line = {
            "impl": "reference", "metric": "pages/sec/box", "value": best["pages_per_s"], "unit": "pages/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": best["seconds_per_page"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload},
            "cpu_baseline": {"value": best["pages_per_s"], "unit": "pages/s", "cores": best["cores"], "kind": "port",
                             "sample": sample, "stages_s": best["stages_s"], "decode_tok_s": best["cpu_decode_tok_s"]},
            "e2e": {"value": best["pages_per_s"], "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

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
    vs = VisionSettings(base, img, crop)
    params = DecodeParameters(max_new_tokens=args.max_new_tokens, no_repeat_ngram_size=20, eos_token_id=None)
    pages = make_pages(args, rank)
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

    # one extra step with per-kernel CUDA-event timing for the roofline / breakdown (not part of the timed value)
    eng.set_option("moe_stats", 1)
    eng.moe_stats()
    eng.kernel_timing_begin()
    step_resident()
    kt = eng.kernel_timing_end()
    seg, nsteps = eng.moe_stats()
    eng.set_option("moe_stats", 0)
    moe_layers = sum(1 for l in range(cfg.num_layers) if l >= cfg.first_k_dense_replace)
    active_experts = seg / max(1.0, nsteps * moe_layers) if nsteps else None
    kt.sort(key=lambda r: -r["ms"])
    total_kernel_ms = sum(r["ms"] for r in kt)
    if args.profile_json and rank == 0:
        Path(args.profile_json).write_text(json.dumps({"kernels": kt, "total_ms": total_kernel_ms, "stage_ms": stage_ms}, indent=1))

    # the same step once more as it really runs (CUDA-graph replay), kernel durations from CUPTI activity records:
    # the per-launch events above serialise the step and add the host launch gap to short kernels
    cupti = {}
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
    cupti_match = {"decode/moe_expert_gate_up": ("linear_sk_kernel", ", 2>"), "decode/moe_expert_down": ("linear_sk_kernel", ", 1>"),
                   "decode/rope_attn_decode": ("rope_attn_decode", "")}

    roof = None
    for r in kt:
        m = kernel_model(cfg, r["name"], args, args.pages, hbm_peak, tf_sustained, active_experts)
        if m is None:
            continue
        avg_s = r["ms"] / r["launches"] * 1e-3
        if m["bound"] == "hbm":
            ach = m["bytes"] / avg_s / 1e9
        else:
            ach = m["flops"] / avg_s / 1e12
        roof = {"kernel": r["name"], "bound": m["bound"], "achieved": ach, "peak": m["peak"], "unit": m["unit"],
                "frac": ach / m["peak"], "traffic": None, "launches_per_step": r["launches"],
                "avg_launch_us": avg_s * 1e6,
                "active_experts_per_layer": active_experts, "share_of_step": r["ms"] / total_kernel_ms, "peak_source": peak_src,
                "timed": "CUDA events after every launch on the engine stream, one extra profiled step"}
        if r["name"] in cupti_match and cupti:
            a, b = cupti_match[r["name"]]
            us = [v for k, v in cupti.items() if a in k and (not b or k.split("(")[0].rstrip().endswith(b))]
            if us:
                avg_us = sum(v[0] for v in us) / max(1, sum(v[1] for v in us))
                ach_g = (m["bytes"] if m["bound"] == "hbm" else m["flops"] * 1e-3) / (avg_us * 1e-6) / 1e9
                roof.update({"achieved_graph": ach_g, "frac_graph": ach_g / m["peak"], "avg_kernel_us_graph": avg_us,
                             "timed_graph": "CUPTI kernel records (torch.profiler) of one more step running as the "
                                            "production CUDA-graph replay"})
        break

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pages_total = args.pages * world * args.steps
    value = pages_total / (ms_res * 1e-3)
    e2e_val = pages_total / (ms_e2e * 1e-3)
    line = {
        "metric": "pages/sec/box", "value": value, "unit": "pages/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload, "l2": "per-step working set (weights 6.7 GB + KV + activations) exceeds the 126 MB L2",
                   "parallelism": f"pages sharded over {world} GPU(s), no collective"},
        "e2e": {"value": e2e_val, "unit": "pages/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "host_memory": "pinned",
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_e2e,
        "clocks": clocks,
        "decode_tok_s_per_gpu": gen_tokens / max(1e-9, stage_ms["decode.iterative"] * 1e-3),
        "stage_ms": stage_ms,
        "top_kernels": [{"name": r["name"], "ms": round(r["ms"], 3), "launches": r["launches"]} for r in kt[:12]],
        "roofline": roof,
    }
    if world == 1:
        # BASELINE's second metric (decode tok/s per GPU) in the reference's own shape: one page per call.  Outside the
        # timed region; steps of <= 4 pages run the fused small-batch decode step (csrc/dsq_decode.cu).
        try:
            import numpy as np
            n_img = 273
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
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_sample(args, cfg, ckdir, pages[0], args.max_new_tokens)
            line["cpu_baseline"] = {
                "value": r["pages_per_s"], "unit": "pages/s", "cores": r["cores"], "kind": "port",
                "sample": (f"1 page through the f32 torch-CPU oracle: preprocess + vision + prefill + {r['sample_tokens']} "
                           f"decode steps ({r['sample_wall_s']:.1f} s of CPU work), token loop extrapolated linearly to "
                           f"{args.max_new_tokens} tokens"),
                "stages_s": r["stages_s"], "decode_tok_s": r["cpu_decode_tok_s"]}
            # token agreement of the GPU batch's page 0 with the oracle over the sampled prefix
            n = len(r["first_tokens"])
            line["token_agreement_page0_first_tokens"] = sum(int(a == b) for a, b in zip(outs_res[0].generated_tokens[:n], r["first_tokens"])) / max(1, n)
        except Exception as ex:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "pages/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {ex}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
