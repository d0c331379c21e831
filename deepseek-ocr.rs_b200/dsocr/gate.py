"""The reference benchsuite's strict gate (benchsuite/orchestrator.py:455-521, benchsuite/common.py:99-107): generated
token ids are compared after trimming trailing stop tokens (id 1); the earliest divergence is reported with both values;
prompts must be string-equal.  Restated here so that runs of this engine are judged by the same rule
(tests/golden/gate_cases.json holds outputs of the reference's own function for a set of inputs)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

STOP_TOKEN = 1


def earliest_divergence(a: List[int], b: List[int]) -> Optional[Tuple[int, Optional[int], Optional[int]]]:
    upto = min(len(a), len(b))
    for idx in range(upto):
        if a[idx] != b[idx]:
            return idx, a[idx], b[idx]
    if len(a) != len(b):
        idx = upto
        return idx, a[idx] if idx < len(a) else None, b[idx] if idx < len(b) else None
    return None


def trim_trailing_stop_tokens(tokens: List[int]) -> List[int]:
    end = len(tokens)
    while end > 0 and tokens[end - 1] == STOP_TOKEN:
        end -= 1
    return tokens[:end]


def strict_compare(py_metrics: Dict[str, Any], rs_metrics: Dict[str, Any]) -> Dict[str, Any]:
    """`py_metrics` = the Python (HF) run, `rs_metrics` = the engine's run; keys `generated_token_ids`, `rendered_prompt`."""
    py_tokens = py_metrics.get("generated_token_ids", [])
    rs_tokens = rs_metrics.get("generated_token_ids", [])
    if not isinstance(py_tokens, list) or not isinstance(rs_tokens, list):
        return {"token_match": False, "prompt_match": False, "token_diff": {"reason": "missing generated_token_ids"},
                "prompt_diff": {"reason": "missing rendered_prompt"}}
    raw = earliest_divergence(py_tokens, rs_tokens)
    py_norm, rs_norm = trim_trailing_stop_tokens(py_tokens), trim_trailing_stop_tokens(rs_tokens)
    diff = earliest_divergence(py_norm, rs_norm)
    py_prompt, rs_prompt = py_metrics.get("rendered_prompt"), rs_metrics.get("rendered_prompt")
    prompt_match = isinstance(py_prompt, str) and isinstance(rs_prompt, str) and py_prompt == rs_prompt
    prompt_diff = None
    if not prompt_match:
        prompt_diff = {"python_len": len(py_prompt) if isinstance(py_prompt, str) else None,
                       "rust_len": len(rs_prompt) if isinstance(rs_prompt, str) else None}

    def payload(d):
        return None if d is None else {"index": d[0], "python": d[1], "rust": d[2]}

    return {
        "token_match": diff is None,
        "prompt_match": prompt_match,
        "token_diff": payload(diff),
        "token_diff_raw": payload(raw),
        "token_counts": {"python_raw": len(py_tokens), "rust_raw": len(rs_tokens), "python_normalized": len(py_norm),
                         "rust_normalized": len(rs_norm)},
        "trailing_stop_normalized": bool(raw is not None and diff is None),
        "prompt_diff": prompt_diff,
    }


def token_agreement(ref: List[int], got: List[int]) -> float:
    """Fraction of positions that agree after the trim (BASELINE.json target: >= 0.95 on fixture pages)."""
    a, b = trim_trailing_stop_tokens(list(ref)), trim_trailing_stop_tokens(list(got))
    n = max(len(a), len(b))
    if n == 0:
        return 1.0
    return sum(int(x == y) for x, y in zip(a, b)) / n
