#!/usr/bin/env python
"""Runs reference-schema golden baselines (e.g. the reference's `baselines/long/<case>` directories, or
tests/golden/fixture_tiny) through the engine and reports token agreement the way the reference's
long_generation_baseline test judges it.  Usage:
  python scripts/run_baseline.py --model-config config.json --weights model.safetensors [--snapshot x.dsq] DIR [DIR ...]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--model-config", required=True)
    ap.add_argument("--weights", required=True)
    ap.add_argument("--snapshot")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--dtype", default="bf16", choices=["f16", "bf16"])
    ap.add_argument("dirs", nargs="+")
    a = ap.parse_args(argv)
    from dsocr.baselines import load_baseline, run_baseline
    from dsocr.engine import load_model

    eng = load_model(a.model_config, a.weights, a.snapshot, a.device, a.dtype)
    ok = True
    for d in a.dirs:
        res = run_baseline(eng, load_baseline(d))
        print(json.dumps({"baseline": d, **res}))
        ok &= bool(res.get("match", True))
    eng.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
