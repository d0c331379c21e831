"""Page sharding across the GPUs of one box.

Pages are independent (`generate` is batch-1 in the reference and `PromptCacheGuard` clears all per-prompt state:
crates/core/src/cache.rs:347-382), so the multi-GPU mode is one engine replica per GPU and NO data-path
collective: rank r of w processes pages r, r+w, r+2w, ...  Only the host-side result gather uses
torch.distributed (object gather of token ids, off the timed path)."""
from __future__ import annotations

from typing import List, Sequence


def shard_indices(n_pages: int, rank: int, world: int) -> List[int]:
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"invalid rank {rank} / world {world}")
    return list(range(rank, n_pages, world))


def merge_shards(shards: Sequence[Sequence], n_pages: int) -> list:
    """Inverse of shard_indices: shards[r][i] is the result of page r + i*world."""
    world = len(shards)
    out = [None] * n_pages
    for r, sh in enumerate(shards):
        idx = shard_indices(n_pages, r, world)
        if len(idx) != len(sh):
            raise ValueError(f"rank {r} returned {len(sh)} results for {len(idx)} pages")
        for i, v in zip(idx, sh):
            out[i] = v
    return out


def gather_results(local_results: list, n_pages: int) -> list | None:
    """All ranks call this; rank 0 gets the page-ordered list, others None."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local_results)
    world, rank = dist.get_world_size(), dist.get_rank()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(local_results), gathered, dst=0)
    return merge_shards(gathered, n_pages) if rank == 0 else None
