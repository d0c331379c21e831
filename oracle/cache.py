"""KV-cache bookkeeping and the padding-mask builder of the reference, restated (test infrastructure only):
crates/core/src/cache.rs (KvCacheChunk :15-57, KvCacheEntry :60-238, LayerKvCache :241-337, DynamicCache +
PromptCacheGuard :340-471) and transformer/block.rs:1476-1495 (lengths_to_padding_mask).  The GPU engine keeps one
preallocated cache per generate call (created and dropped inside the call, like PromptCacheGuard), so these rules have no
API of their own there; they are pinned here because the reference's own unit tests pin them
(crates/infer-deepseek/tests/transformer_cache.rs:20-101, transformer_block.rs:59-67) and the oracle's decode loop relies on
the same growth rule (positions = arange(past, past + q))."""
from __future__ import annotations

from contextlib import contextmanager
from typing import Callable, List, Optional

import torch


def lengths_to_padding_mask(lengths, seq_len: int) -> torch.Tensor:
    out = torch.zeros(len(lengths), seq_len, dtype=torch.float32)
    for b, n in enumerate(lengths):
        if n > seq_len:
            raise ValueError(f"length {n} exceeds sequence dimension {seq_len}")
        out[b, :n] = 1.0
    return out


class KvCacheChunk:
    """key_t [batch, heads, dim, seq], value [batch, heads, seq, dim]."""

    def __init__(self, key_t: torch.Tensor, value: torch.Tensor):
        if key_t.dim() != 4:
            raise ValueError(f"expected key chunk tensor with rank 4 [batch, heads, dim, seq], got rank {key_t.dim()}")
        if value.dim() != 4:
            raise ValueError(f"expected value chunk tensor with rank 4 [batch, heads, seq, dim], got rank {value.dim()}")
        kb, kh, _, ks = key_t.shape
        vb, vh, vs, _ = value.shape
        if kb != vb:
            raise ValueError(f"chunk batch mismatch between key ({kb}) and value ({vb})")
        if kh != vh:
            raise ValueError(f"chunk heads mismatch between key ({kh}) and value ({vh})")
        if ks != vs:
            raise ValueError(f"chunk sequence mismatch between key ({ks}) and value ({vs})")
        self.key_t, self.value = key_t, value

    def seq_len(self) -> int:
        return self.key_t.shape[-1]


class KvCacheEntry:
    def __init__(self, chunk: KvCacheChunk):
        self.chunks: List[KvCacheChunk] = [chunk]

    def append(self, chunk: KvCacheChunk) -> None:
        first = self.chunks[0]
        b, h, kd, _ = first.key_t.shape
        cb, ch, ckd, _ = chunk.key_t.shape
        if cb != b:
            raise ValueError(f"chunk batch {cb} does not match cache batch {b}")
        if ch != h:
            raise ValueError(f"chunk heads {ch} does not match cache heads {h}")
        if ckd != kd:
            raise ValueError(f"chunk key dim {ckd} does not match cache key dim {kd}")
        if chunk.key_t.dtype != first.key_t.dtype:
            raise ValueError(f"chunk dtype {chunk.key_t.dtype} does not match cache dtype {first.key_t.dtype}")
        if chunk.value.shape[-1] != first.value.shape[-1]:
            raise ValueError(f"chunk value dim {chunk.value.shape[-1]} does not match cache value dim {first.value.shape[-1]}")
        self.chunks.append(chunk)

    def seq_len(self) -> int:
        return sum(c.seq_len() for c in self.chunks)

    def key_view(self) -> torch.Tensor:    # cat of all chunks along seq (cache.rs:204-213) - the O(S) copy per step
        return torch.cat([c.key_t for c in self.chunks], dim=-1)

    def value_view(self) -> torch.Tensor:
        return torch.cat([c.value for c in self.chunks], dim=-2)


class LayerKvCache:
    def __init__(self, num_layers: int = 0):
        self.entries: List[Optional[KvCacheEntry]] = [None] * num_layers

    def __len__(self) -> int:
        return len(self.entries)

    def get(self, layer: int) -> Optional[KvCacheEntry]:
        return self.entries[layer] if layer < len(self.entries) else None

    def append_chunk(self, layer: int, chunk: KvCacheChunk) -> None:
        if layer >= len(self.entries):
            self.entries += [None] * (layer + 1 - len(self.entries))
        if self.entries[layer] is not None:
            self.entries[layer].append(chunk)
        else:
            self.entries[layer] = KvCacheEntry(chunk)

    def clear(self) -> None:
        self.entries = [None] * len(self.entries)

    def seq_len(self) -> Optional[int]:
        lens = [e.seq_len() for e in self.entries if e is not None]
        return max(lens) if lens else None


class DynamicCache:
    def __init__(self, num_layers: int = 0):
        self.layers = LayerKvCache(num_layers)
        self._seq_len: Optional[int] = None

    def get(self, layer: int) -> Optional[KvCacheEntry]:
        return self.layers.get(layer)

    def append(self, layer: int, chunk: KvCacheChunk) -> None:
        cur = self.layers.get(layer)
        new_len = (cur.seq_len() if cur is not None else 0) + chunk.seq_len()
        if self._seq_len is not None and new_len < self._seq_len:
            raise ValueError(f"cache seq_len decreased for layer {layer}: {new_len} < {self._seq_len}")
        if self._seq_len is None or new_len > self._seq_len:
            self._seq_len = new_len
        self.layers.append_chunk(layer, chunk)

    def seq_len(self) -> Optional[int]:
        return self._seq_len

    def clear(self) -> None:
        self.layers.clear()
        self._seq_len = None

    @contextmanager
    def prompt_guard(self, reset: Optional[Callable[[], None]] = None):
        """PromptCacheGuard: the cache is cleared (and the optional RoPE reset hook run) when the guard goes out of scope."""
        self.clear()
        try:
            yield self
        finally:
            self.clear()
            if reset is not None:
                reset()
