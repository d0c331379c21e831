"""Multi-GPU host logic on CPU: pages are sharded round-robin over ranks with no data-path collective; results are
gathered page-ordered on rank 0 (gloo, world_size 2)."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

from dsocr.sharding import merge_shards, shard_indices  # noqa: E402


def test_shard_indices_partition():
    for n in (0, 1, 7, 64, 1024):
        for w in (1, 2, 4, 8):
            seen = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def test_merge_shards_restores_order_and_checks_counts():
    shards = [[f"p{i}" for i in shard_indices(7, r, 3)] for r in range(3)]
    assert merge_shards(shards, 7) == [f"p{i}" for i in range(7)]
    with pytest.raises(ValueError):
        merge_shards([["a"], []], 3)


def _worker(rank, world, port, n_pages, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dsocr.sharding import gather_results, shard_indices

    mine = [(i, [i * 10 + k for k in range(3)]) for i in shard_indices(n_pages, rank, world)]  # fake token ids per page
    out = gather_results(mine, n_pages)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(i, [i * 10, i * 10 + 1, i * 10 + 2]) for i in range(9)]
