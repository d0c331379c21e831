#!/usr/bin/env python
"""Golden vectors for the token gate: outputs of the REFERENCE's own `BenchOrchestrator._strict_compare`
(benchsuite/orchestrator.py:455-521) and of its schema parsers (benchsuite/schemas.py) imported from /root/reference in
the build container.  tests/test_report_cpu.py holds dsocr/gate.py and dsocr/report.py to them."""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")
from benchsuite.orchestrator import BenchOrchestrator  # noqa: E402

CASES = [
    {"py": [5, 6, 7, 8], "rs": [5, 6, 7, 8], "pp": "<image>\nFree OCR.", "rp": "<image>\nFree OCR."},
    {"py": [5, 6, 7, 8, 1], "rs": [5, 6, 7, 8], "pp": "a", "rp": "a"},            # trailing EOS on one side only
    {"py": [5, 6, 7, 8, 1, 1], "rs": [5, 6, 7, 8, 1], "pp": "a", "rp": "a"},
    {"py": [5, 6, 9, 8], "rs": [5, 6, 7, 8], "pp": "a", "rp": "a"},               # divergence at index 2
    {"py": [5, 6, 7], "rs": [5, 6, 7, 8, 9], "pp": "a", "rp": "a"},               # one side longer
    {"py": [], "rs": [], "pp": "a", "rp": "b"},                                   # prompt mismatch
    {"py": [1, 1], "rs": [], "pp": "a", "rp": "a"},                               # only stop tokens
    {"py": [5, 1, 6], "rs": [5, 1, 6, 1], "pp": "a", "rp": None},                 # inner stop token kept, missing prompt
    {"py": "oops", "rs": [1], "pp": "a", "rp": "a"},                              # malformed input
]

out = []
for c in CASES:
    py = {"generated_token_ids": c["py"], "rendered_prompt": c["pp"]}
    rs = {"generated_token_ids": c["rs"], "rendered_prompt": c["rp"]}
    out.append({"python": py, "rust": rs, "expected": BenchOrchestrator._strict_compare(py, rs)})
(HERE / "gate_cases.json").write_text(json.dumps(out, indent=1) + "\n")
print(len(out), "gate cases written")
