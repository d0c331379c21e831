#!/usr/bin/env python
"""NCCL baseline for the expert-parallel exchange (BASELINE.json configs[4] names "NCCL all-to-all over NVLink"): times
torch.distributed.all_to_all_single of exactly the payload one MoE layer of one decode step moves per GPU - dispatch of
pages*6 token rows (hi + lo 16-bit, 2 x 1280 x 2 B each) and combine of as many f32 result rows - for the same pages per
GPU as scripts/bench_ep.py, plus the per-step figure (11 MoE layers x 2 all-to-alls).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_ep_nccl.py --pages 128"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pages", type=int, default=128)
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    H, topk = 1280, 6
    rows_per_peer = (a.pages * topk + world - 1) // world          # balanced routing: rows each rank sends to each peer
    disp = torch.empty(world * rows_per_peer, 2 * H, dtype=torch.bfloat16, device="cuda")   # hi + lo
    comb = torch.empty(world * rows_per_peer, H, dtype=torch.float32, device="cuda")
    d_out, c_out = torch.empty_like(disp), torch.empty_like(comb)
    for _ in range(20):
        dist.all_to_all_single(d_out, disp); dist.all_to_all_single(c_out, comb)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        dist.all_to_all_single(d_out, disp); dist.all_to_all_single(c_out, comb)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        per_layer_us = float(ms.item()) * 1e3
        print(json.dumps({"world": world, "pages_per_gpu": a.pages, "rows_per_peer": rows_per_peer,
                          "dispatch_bytes_per_gpu": disp.numel() * 2, "combine_bytes_per_gpu": comb.numel() * 4,
                          "nccl_all_to_all_pair_us_per_layer": per_layer_us, "per_step_ms_11_layers": per_layer_us * 11 / 1e3}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
