"""Single-process host dispatcher over the GPUs of one box (SURVEY.md 8e "host dispatcher assigns pages round-robin /
work-stealing").  The reference's server is ONE process that owns its engine behind `Arc<Mutex<..>>`
(crates/server/src/state.rs:210-224); the drop-in equivalent on an 8-GPU box is one process that owns one engine
replica per GPU and drives them from worker threads (every C-ABI call releases the GIL; an engine is used by one thread
at a time, which is the contract the reference's mutex gives).  Pages are independent, so there is no data-path
collective: the page list is cut into groups that the workers pull from a shared queue - a GPU that finishes early (short
outputs, EOS) takes the next group instead of idling behind a static split."""
from __future__ import annotations

import ctypes as C
import queue
import threading
from typing import Any, Callable, List, Optional, Sequence

from .engine import DecodeOutcome, DecodeParameters, OcrEngine, VisionSettings, load_model


def plan_groups(n_items: int, n_workers: int, max_group: int) -> List[range]:
    """Contiguous groups of at most `max_group` items, at least one per worker when there are enough items, sized so that
    the last round is not a single straggler group (ceil split of every round)."""
    if n_items <= 0:
        return []
    n_workers = max(1, n_workers)
    rounds = max(1, -(-n_items // (n_workers * max_group)))       # ceil
    n_groups = min(n_items, rounds * n_workers)
    base, extra = divmod(n_items, n_groups)
    out, start = [], 0
    for g in range(n_groups):
        size = base + (1 if g < extra else 0)
        out.append(range(start, start + size))
        start += size
    return out


class EnginePool:
    def __init__(self, engines: Sequence[Any], max_group: int = 512):
        if not engines:
            raise ValueError("EnginePool needs at least one engine")
        self.engines = list(engines)
        self.max_group = max_group
        self.last_assignment: List[List[int]] = []   # per engine: sizes of the groups it processed (observability / tests)

    @classmethod
    def load(cls, config_path: str, weights_path: str, snapshot_path: Optional[str], devices: Sequence[int],
             dtype: str = "bf16", max_group: int = 512, configure: Optional[Callable[[OcrEngine], None]] = None) -> "EnginePool":
        """`load_model` once per device ordinal, in parallel (weights upload + re-tiling is per GPU)."""
        engines: List[Optional[OcrEngine]] = [None] * len(devices)
        errors: List[BaseException] = []

        def work(i: int, dev: int):
            try:
                e = load_model(config_path, weights_path, snapshot_path, dev, dtype)
                if configure:
                    configure(e)
                engines[i] = e
            except BaseException as ex:  # noqa: BLE001 - reported to the caller below
                errors.append(ex)

        th = [threading.Thread(target=work, args=(i, d)) for i, d in enumerate(devices)]
        [t.start() for t in th]
        [t.join() for t in th]
        if errors:
            for e in engines:
                if e is not None:
                    e.close()
            raise errors[0]
        return cls(engines, max_group)

    def close(self):
        self.disable_expert_parallel()
        for e in self.engines:
            e.close()

    # ---- expert-parallel decode (BASELINE configs[4]; include/dsocr.h dsocr_ep_group_create) -------------------------
    _ep_group = None
    _ep_max_pages = 0

    def enable_expert_parallel(self, max_pages_per_engine: int = 128):
        """Links the pool's engines into one expert-parallel group: from now on decode_pages / decode_requests give every
        engine one equally sized shard per round and run the rounds in lock-step (the group's cross-GPU barriers pair up)."""
        from .binding import check, lib

        if self._ep_group is not None:
            return
        n = len(self.engines)
        handles = (C.c_void_p * n)(*[e._h for e in self.engines])
        group = C.c_void_p()
        check(lib().dsocr_ep_group_create(handles, n, int(max_pages_per_engine), C.byref(group)), "ep_group_create")
        self._ep_group, self._ep_max_pages = group, int(max_pages_per_engine)

    def disable_expert_parallel(self):
        if self._ep_group is not None:
            from .binding import lib

            lib().dsocr_ep_group_destroy(self._ep_group)
            self._ep_group = None

    def _run_lockstep(self, n_items: int, call: Callable[[Any, range], Sequence[Any]]) -> List[Any]:
        n = len(self.engines)
        results: List[Any] = [None] * n_items
        self.last_assignment = [[] for _ in self.engines]
        per_round = n * self._ep_max_pages
        start = 0
        while start < n_items:
            count = min(per_round, n_items - start)
            base, extra = divmod(count, n)
            if base + (1 if extra else 0) > self._ep_max_pages or base <= 4:
                raise ValueError(f"expert-parallel rounds need 5..{self._ep_max_pages} pages per engine, got {base} (+{extra})")
            shards, s0 = [], start
            for w in range(n):
                size = base + (1 if w < extra else 0)
                shards.append(range(s0, s0 + size))
                s0 += size
            errors: List[BaseException] = []

            def worker(w: int):
                try:
                    out = call(self.engines[w], shards[w])
                    for i, o in zip(shards[w], out):
                        results[i] = o
                    self.last_assignment[w].append(len(shards[w]))
                except BaseException as ex:  # noqa: BLE001
                    errors.append(ex)

            th = [threading.Thread(target=worker, args=(w,), name=f"dsocr-ep{w}") for w in range(n)]
            [t.start() for t in th]
            [t.join() for t in th]
            if errors:
                raise errors[0]
            start += count
        return results

    def _run(self, n_items: int, call: Callable[[Any, range], Sequence[Any]]) -> List[Any]:
        if self._ep_group is not None:
            return self._run_lockstep(n_items, call)
        groups = plan_groups(n_items, len(self.engines), self.max_group)
        q: "queue.Queue[range]" = queue.Queue()
        for g in groups:
            q.put(g)
        results: List[Any] = [None] * n_items
        errors: List[BaseException] = []
        self.last_assignment = [[] for _ in self.engines]

        def worker(w: int):
            eng = self.engines[w]
            while not errors:
                try:
                    g = q.get_nowait()
                except queue.Empty:
                    return
                try:
                    out = call(eng, g)
                    if len(out) != len(g):
                        raise RuntimeError(f"engine returned {len(out)} results for {len(g)} pages")
                    for i, o in zip(g, out):
                        results[i] = o
                    self.last_assignment[w].append(len(g))
                except BaseException as ex:  # noqa: BLE001
                    errors.append(ex)

        th = [threading.Thread(target=worker, args=(w,), name=f"dsocr-gpu{w}") for w in range(len(self.engines))]
        [t.start() for t in th]
        [t.join() for t in th]
        if errors:
            raise errors[0]
        return results

    def decode_pages(self, pages_rgb: Sequence[Any], vs: VisionSettings, seg0: Sequence[int], seg1: Sequence[int],
                     image_token_id: int, params: DecodeParameters) -> List[DecodeOutcome]:
        """`OcrEngine::decode` over a page list, results in page order."""
        return self._run(len(pages_rgb), lambda eng, g: eng.decode_pages([pages_rgb[i] for i in g], vs, seg0, seg1,
                                                                         image_token_id, params))

    def decode_requests(self, requests: Sequence[tuple], vs: VisionSettings, image_token_id: int,
                        params: DecodeParameters) -> List[DecodeOutcome]:
        return self._run(len(requests), lambda eng, g: eng.decode_requests([requests[i] for i in g], vs, image_token_id, params))
