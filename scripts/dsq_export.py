#!/usr/bin/env python
"""`deepseek-ocr-cli weights snapshot` equivalent (crates/cli/src/args.rs:36-58): --in <safetensors> --out <path>
--dtype q8_0|q4k|q6k --targets text|text+projector, plus --config (the checkpoint's config.json)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "deepseek-ocr.rs_b200"))

from dsocr.export import export_snapshot  # noqa: E402


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--config", required=True)
    ap.add_argument("--in", dest="input", required=True)
    ap.add_argument("--out", dest="output", required=True)
    ap.add_argument("--dtype", default="q8_0")
    ap.add_argument("--targets", default="text", choices=["text", "text+projector"])
    a = ap.parse_args(argv)
    written = export_snapshot(a.config, a.input, a.output, a.dtype, a.targets)
    by = {}
    for code in written.values():
        by[code] = by.get(code, 0) + 1
    print(f"wrote {Path(a.output).with_suffix('.dsq')}: {len(written)} tensors, dtype code counts {by}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
