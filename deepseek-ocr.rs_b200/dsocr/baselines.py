"""Golden-baseline fixtures in the reference's own schema (crates/infer-deepseek/tests/long_generation_baseline.rs:30-150):
a directory with `baseline.json` {variant, prompt, image, base_size, image_size, crop_mode, ...}, `prompt.json`
{rendered_prompt, input_ids, images_seq_mask, image_token_ranges, vision_token_total, bos_token_id, image_token_id,
prefill_len, ...} and `output_tokens.json` {tokens, prefill_len, generated_len, eos_token_id}.  The reference's test decodes
the page greedily (no_repeat_ngram_size 20, use_cache) and requires the generated ids to equal `tokens[prefill_len:]` with
a trailing EOS dropped.  `load_baseline` applies the same consistency checks; `run_baseline` drives an engine from the
fixture's own token ids (no tokenizer needed) so that real `baselines/long/<case>` directories can be dropped in unchanged."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

from .gate import earliest_divergence


@dataclass
class Baseline:
    directory: str
    variant: str
    prompt: str
    image: str
    base_size: int
    image_size: int
    crop_mode: bool
    input_ids: List[int]
    images_seq_mask: List[int]
    image_token_id: int
    vision_token_total: int
    expected: List[int]          # expected_generated_tokens (:143-151)
    requested_tokens: int        # out.generated_len
    eos_token_id: Optional[int]

    def segments(self) -> Tuple[List[int], List[int], int]:
        """(ids before the image placeholders without BOS, ids after them, number of placeholders)."""
        idx = [i for i, m in enumerate(self.images_seq_mask) if m]
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise ValueError("fixture must contain exactly one contiguous run of image placeholders")
        return self.input_ids[1:idx[0]], self.input_ids[idx[-1] + 1:], len(idx)


def load_baseline(directory: str) -> Baseline:
    def read(name: str) -> Dict[str, Any]:
        with open(os.path.join(directory, name)) as f:
            return json.load(f)

    meta = read("baseline.json")
    p = meta.get("prompt_assets_path") or "prompt.json"
    o = meta.get("output_tokens_path") or "output_tokens.json"
    prompt = read(p) if "/" not in p else json.load(open(p))
    out = read(o) if "/" not in o else json.load(open(o))
    ids, mask = [int(v) for v in prompt["input_ids"]], [int(v) for v in prompt["images_seq_mask"]]
    n = len(ids)
    # the checks of run_one_baseline (:196-228), same wording
    if int(prompt["prefill_len"]) != n:
        raise ValueError(f"prefill_len {prompt['prefill_len']} != input_ids len {n}")
    if int(out["prefill_len"]) != n:
        raise ValueError(f"output prefill_len {out['prefill_len']} != prompt len {n}")
    if len(mask) != n:
        raise ValueError(f"images_seq_mask len {len(mask)} != prompt len {n}")
    if not prompt.get("image_token_ranges"):
        raise ValueError("expected at least one image token range")
    if int(prompt.get("vision_token_total", 0)) <= 0:
        raise ValueError("expected non-zero vision_token_total")
    tokens = [int(v) for v in out["tokens"]]
    eos = out.get("eos_token_id")
    expected = tokens[n:]
    if eos is not None and expected and expected[-1] == int(eos):
        expected = expected[:-1]
    image = meta["image"]
    if not os.path.isabs(image):
        image = os.path.join(directory, image) if os.path.exists(os.path.join(directory, image)) else image
    return Baseline(directory=directory, variant=str(meta.get("variant", "")), prompt=str(meta.get("prompt", "")), image=image,
                    base_size=int(meta.get("base_size") or 1024), image_size=int(meta.get("image_size") or 640),
                    crop_mode=bool(True if meta.get("crop_mode") is None else meta.get("crop_mode")), input_ids=ids,
                    images_seq_mask=mask, image_token_id=int(prompt["image_token_id"]),
                    vision_token_total=int(prompt["vision_token_total"]), expected=expected,
                    requested_tokens=int(out["generated_len"]), eos_token_id=None if eos is None else int(eos))


def compare(expected: List[int], got: List[int]) -> Dict[str, Any]:
    """first_mismatch (:92-107) plus the agreement ratio BASELINE.json's target is stated in."""
    d = earliest_divergence(expected, got)
    n = max(len(expected), len(got), 1)
    return {"match": d is None, "first_mismatch": None if d is None else d[0],
            "agreement": sum(int(a == b) for a, b in zip(expected, got)) / n, "expected_len": len(expected), "got_len": len(got)}


def run_baseline(engine, baseline: Baseline) -> Dict[str, Any]:
    """Greedy decode of the fixture page on `engine` (dsocr.engine.OcrEngine) with the reference test's parameters."""
    import numpy as np
    from PIL import Image

    from .engine import DecodeParameters, VisionSettings

    if baseline.variant != "ocr1":
        return {"skipped": f"variant `{baseline.variant}` is not served by this engine"}
    seg0, seg1, n_img = baseline.segments()
    page = np.asarray(Image.open(baseline.image).convert("RGB"))
    params = DecodeParameters(max_new_tokens=baseline.requested_tokens, no_repeat_ngram_size=20, repetition_penalty=1.0,
                              eos_token_id=baseline.eos_token_id if baseline.eos_token_id is not None else 1, use_cache=True)
    out = engine.decode_pages([page], VisionSettings(baseline.base_size, baseline.image_size, baseline.crop_mode), seg0, seg1,
                              baseline.image_token_id, params)[0]
    res = compare(baseline.expected, out.generated_tokens)
    res["prompt_tokens"] = out.prompt_tokens
    res["prompt_tokens_expected"] = len(baseline.input_ids)
    res["image_tokens_expected"] = n_img
    return res
