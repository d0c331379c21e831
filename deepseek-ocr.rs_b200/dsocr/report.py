"""Run reports in the reference CLI's own JSON schemas, so that the reference's `benchsuite` (strict token gate and
perf tables) can consume runs of this engine unchanged.

  * `write_output_json`  == `--output-json`  (crates/cli/src/debug.rs:100-157, struct CliOutputJson, schema_version 1);
    parsed by benchsuite/schemas.py:40-61 (RustDecodeOutput: `tokens`, `prompt_tokens`, `generated_len`, `rendered_prompt`)
  * `BenchRecorder.write` == `--bench-output` (crates/cli/src/bench.rs:138-249: `events` + `stage_totals` with
    count / total / min / max per stage); parsed by benchsuite/schemas.py:64-84 (StageTotals.stage_ms)
The stage names are the reference Timer's (core/src/benchmark.rs; benchsuite/models/base.py:60-63)."""
from __future__ import annotations

import json
import os
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional, Sequence

STAGE_LOAD = "model.load"
STAGE_PROMPT = "prompt.render"
STAGE_VISION_PREPARE = "vision.prepare_inputs"
STAGE_VISION_EMBED = "vision.compute_embeddings"
STAGE_PREFILL = "decode.prefill"
STAGE_ITERATIVE = "decode.iterative"
STAGE_GENERATE = "decode.generate"


@dataclass
class CliOutput:
    """Field for field the reference's CliOutputJson (debug.rs:108-131)."""
    model_id: str
    weights: str
    tokenizer: str
    device: str
    dtype: str
    template: str
    base_size: int
    image_size: int
    crop_mode: bool
    max_new_tokens: int
    repetition_penalty: float
    no_repeat_ngram_size: Optional[int]
    use_cache: bool
    prompt: str
    rendered_prompt: str
    image_paths: List[str]
    prompt_tokens: int
    generated_len: int
    tokens: List[int]
    decoded: str
    normalized: str
    schema_version: int = 1

    def to_json(self) -> Dict[str, Any]:
        d = asdict(self)
        return {"schema_version": d.pop("schema_version"), **d}  # serde keeps declaration order: schema_version first


def write_output_json(path: str, out: CliOutput) -> None:
    parent = os.path.dirname(path)
    if parent:
        os.makedirs(parent, exist_ok=True)
    with open(path, "w") as f:
        json.dump(out.to_json(), f, indent=2)


@dataclass
class BenchEvent:
    stage: str
    duration_ns: int
    fields: Dict[str, Any] = field(default_factory=dict)


class BenchRecorder:
    """Collector of crates/cli/src/bench.rs: events in the order they were recorded, totals per stage."""

    def __init__(self) -> None:
        self.events: List[BenchEvent] = []

    def record(self, stage: str, seconds: float, **fields: Any) -> None:
        self.events.append(BenchEvent(stage, int(round(seconds * 1e9)), dict(fields)))

    def record_ms(self, stage: str, ms: float, **fields: Any) -> None:
        self.record(stage, ms * 1e-3, **fields)

    def stage_totals(self) -> List[Dict[str, Any]]:
        acc: Dict[str, Dict[str, int]] = {}
        for e in self.events:  # totals_for_events (bench.rs:138-170)
            a = acc.setdefault(e.stage, {"count": 0, "total": 0, "min": None, "max": None})
            a["count"] += 1
            a["total"] += e.duration_ns
            a["min"] = e.duration_ns if a["min"] is None or e.duration_ns < a["min"] else a["min"]
            a["max"] = e.duration_ns if a["max"] is None or e.duration_ns > a["max"] else a["max"]
        out = []
        for stage, a in acc.items():  # the reference iterates a HashMap: consumers index by stage, order is not part of the schema
            out.append({"stage": stage, "count": a["count"], "total_ms": a["total"] / 1e6, "total_ns": str(a["total"]),
                        "avg_ms": (a["total"] / 1e6 / a["count"]) if a["count"] else 0.0,
                        "min_ms": (a["min"] or 0) / 1e6, "max_ms": (a["max"] or 0) / 1e6})
        return out

    def to_json(self) -> Dict[str, Any]:
        return {
            "events": [{"stage": e.stage, "duration_ms": e.duration_ns / 1e6, "duration_ns": str(e.duration_ns),
                        "fields": [{"key": k, "value": v} for k, v in e.fields.items()]} for e in self.events],
            "stage_totals": self.stage_totals(),
        }

    def write(self, path: str) -> None:
        parent = os.path.dirname(path)
        if parent:
            os.makedirs(parent, exist_ok=True)
        with open(path, "w") as f:
            json.dump(self.to_json(), f, indent=2)


def record_engine_timings(rec: BenchRecorder, timings: Dict[str, float], prompt_tokens: int, generated: int) -> None:
    """dsocr_last_timings (same stage names as the reference's Timer) -> events, with the fields the reference attaches."""
    for stage in (STAGE_VISION_PREPARE, STAGE_VISION_EMBED, STAGE_PREFILL, STAGE_ITERATIVE, STAGE_GENERATE):
        if stage in timings:
            extra = {}
            if stage == STAGE_PREFILL:
                extra = {"prompt_tokens": prompt_tokens}
            elif stage in (STAGE_ITERATIVE, STAGE_GENERATE):
                extra = {"generated_tokens": generated}
            rec.record_ms(stage, float(timings[stage]), **extra)


def split_prompt_on_image(rendered_prompt: str, image_token: str = "<image>") -> List[str]:
    """build_prompt_tokens splits the rendered prompt on the literal `<image>` (model/mod.rs:2536-2560): n images give
    n + 1 text segments (possibly empty)."""
    return rendered_prompt.split(image_token)


def tokenize_segments(tokenizer: Any, segments: Sequence[str]) -> List[List[int]]:
    """Text segments -> ids without special tokens (the reference encodes each segment with add_special_tokens = false
    and prepends BOS itself).  `tokenizer` is anything with `.encode(text, add_special_tokens=False)` returning either
    ids or an object with `.ids` (tokenizers.Tokenizer, a transformers tokenizer, or a test double)."""
    out = []
    for s in segments:
        if not s:
            out.append([])
            continue
        enc = tokenizer.encode(s, add_special_tokens=False)
        out.append(list(enc.ids if hasattr(enc, "ids") else enc))
    return out
