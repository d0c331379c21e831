"""Request batching in front of the engine.  The reference serves one request at a time: its engine sits
behind `Arc<Mutex<..>>` and every `/v1/chat/completions` call runs `decode` under that lock (server/src/state.rs:210-224,
generation.rs:84-103), because `generate` is batch-1.  This engine decodes many pages in lock step, so concurrent requests
are collected for a few milliseconds and run as one batch; tokens stream back per request through the engine's per-page
callback.  Requests are only batched together when they agree on what one engine call shares (vision settings, decode
parameters, the <image> token id); prompts and image counts may differ per request (dsocr_decode_requests)."""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Sequence, Tuple


@dataclass
class PageRequest:
    page: Any                                  # RGB8 H x W x 3 (the first image; None for a text-only prompt)
    seg0: Tuple[int, ...]                      # prompt token ids before the first <image>
    seg1: Tuple[int, ...]                      # ... and after it
    image_token_id: int
    vision: Tuple[int, int, bool]              # base_size, image_size, crop_mode
    params: Tuple                              # hashable decode parameters (max_new_tokens, ngram, eos, sampling ...)
    on_tokens: Optional[Callable[[int, List[int]], None]] = None   # (count, all generated ids) after every accepted token
    more: Tuple = ()                           # further (image, token ids after its <image>) pairs of a multi-image prompt
    future: Future = field(default_factory=Future)
    enqueued: float = field(default_factory=time.perf_counter)

    def key(self):
        return (self.image_token_id, self.vision, self.params)

    def images(self) -> list:
        return ([] if self.page is None else [self.page]) + [m[0] for m in self.more]

    def segments(self) -> list:
        if self.page is None:
            return [list(self.seg0)]
        return [list(self.seg0), list(self.seg1)] + [list(m[1]) for m in self.more]


class PageBatcher:
    """`run_batch(requests)` is called on the worker thread with 1..max_batch compatible requests and returns one result
    per request (anything; it is handed to the request's future).  It may call `request.on_tokens` while it runs."""

    def __init__(self, run_batch: Callable[[List[PageRequest]], Sequence[Any]], max_batch: int = 64, max_wait_ms: float = 5.0):
        self._run = run_batch
        self.max_batch = max_batch
        self.max_wait = max_wait_ms * 1e-3
        self._q: "queue.Queue[Optional[PageRequest]]" = queue.Queue()
        self._held: List[PageRequest] = []     # requests pulled from the queue that did not fit the batch being formed
        self.batches: List[int] = []           # sizes of the batches run so far (observability / tests)
        self._thread = threading.Thread(target=self._loop, name="dsocr-batcher", daemon=True)
        self._closed = False
        self._thread.start()

    def submit(self, req: PageRequest) -> Future:
        if self._closed:
            raise RuntimeError("batcher is closed")
        self._q.put(req)
        return req.future

    def close(self) -> None:
        self._closed = True
        self._q.put(None)
        self._thread.join()

    # ------------------------------------------------------------------------------------------------
    def _loop(self) -> None:
        while True:
            first = self._held.pop(0) if self._held else self._q.get()
            if first is None:
                if self._held:
                    continue
                return
            batch, skipped = [first], []
            # compatible requests that an earlier round had to set aside join this batch first
            for item in self._held:
                (batch if item.key() == first.key() and len(batch) < self.max_batch else skipped).append(item)
            self._held = []
            deadline = time.perf_counter() + self.max_wait
            stop = False
            while len(batch) < self.max_batch:
                remaining = deadline - time.perf_counter()
                try:
                    item = self._q.get(timeout=remaining) if remaining > 0 else self._q.get_nowait()
                except queue.Empty:
                    break
                if item is None:
                    stop = True
                    break
                (batch if item.key() == first.key() else skipped).append(item)
            self._held = skipped + self._held
            self.batches.append(len(batch))
            try:
                results = self._run(batch)
                if len(results) != len(batch):
                    raise RuntimeError(f"engine returned {len(results)} results for {len(batch)} requests")
                for r, res in zip(batch, results):
                    r.future.set_result(res)
            except BaseException as ex:  # every waiting request learns about the failure
                for r in batch:
                    if not r.future.done():
                        r.future.set_exception(ex)
            if stop:
                self._q.put(None)


def engine_runner(engine, make_params, make_vision) -> Callable[[List[PageRequest]], Sequence[Any]]:
    """run_batch over a dsocr.engine.OcrEngine: one `decode_requests` call per batch; the per-request callback of the
    engine (count, all generated ids) is routed to the request that owns it."""
    def run(batch: List[PageRequest]):
        first = batch[0]

        def cb(page: int, count: int, tokens: List[int]):
            r = batch[page]
            if r.on_tokens is not None:
                r.on_tokens(count, tokens)

        want_cb = any(r.on_tokens is not None for r in batch)
        return engine.decode_requests([(r.images(), r.segments()) for r in batch], make_vision(first.vision),
                                      first.image_token_id, make_params(first.params), cb if want_cb else None)
    return run
