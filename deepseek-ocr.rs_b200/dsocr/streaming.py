"""Token-by-token text streaming that never emits half a UTF-8 sequence: crates/core/src/streaming.rs restated
(`extract_delta` :4-18, `DeltaTracker` :21-68).  The decoded text of all tokens so far is compared with what was already
sent; while the stream is running, a trailing U+FFFD (a byte-level token that is not a complete character yet) is held back."""
from __future__ import annotations

REPLACEMENT = "�"


def extract_delta(previous: str, current: str) -> str:
    if current.startswith(previous):
        return current[len(previous):]
    n = 0
    for a, b in zip(previous, current):
        if a != b:
            break
        n += 1
    return current[n:]


class DeltaTracker:
    def __init__(self) -> None:
        self.previous = ""

    def reset(self) -> None:
        self.previous = ""

    def advance(self, current: str, is_final: bool) -> str:
        raw = extract_delta(self.previous, current)
        if not raw:
            self.previous = current
            return raw
        if not is_final:
            idx = raw.find(REPLACEMENT)
            if idx == 0:
                return ""
            if idx > 0:
                raw = raw[:idx]
                self.previous += raw
                return raw
        self.previous = current
        return raw

    def snapshot(self) -> str:
        return self.previous
