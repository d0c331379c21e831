#!/usr/bin/env python
"""Counts, per kernel of libdsocr.so, the SASS mnemonics that prove which hardware path a kernel uses (the PTX names never
appear in SASS): UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor copies, UBLKCP = 1-D bulk copies, UBLKRED = bulk reductions (cp.reduce.async.bulk), UTCBAR /
SYNCS = tcgen05.commit / mbarrier traffic, LDTM / STTM = tcgen05.ld / st (TMEM), MUFU = special-function unit, PRMT =
byte permutes (the byte->float path of the GEMVs), I2F = the conversion unit they avoid.  Needs no GPU (cuobjdump).
Usage: python scripts/sass_evidence.py [lib] > profiles/r01_sass_evidence.csv"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "deepseek-ocr.rs_b200" / "lib" / "libdsocr.so"
KEYS = ["UTC.*MMA", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "UTCBAR", "SYNCS", "LDTM", "STTM", "MUFU", "PRMT", "I2F", "HMMA", "FFMA", "LDG", "LDS", "total"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in KEYS[:-1]:
                if re.match(k, op):
                    counts[cur][k] += 1
    names = demangle(list(counts))
    print("kernel," + ",".join(k.replace(".*", "x") for k in KEYS))
    for fn, c in counts.items():
        full = names.get(fn, fn).replace("(anonymous namespace)::", "").replace("dsocr::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", full)
        print('"' + short + '",' + ",".join(str(c[k]) for k in KEYS))


if __name__ == "__main__":
    main()
