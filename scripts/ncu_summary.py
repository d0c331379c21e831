#!/usr/bin/env python
"""Condenses ncu outputs brought back in gpurun_out/ into the small tracked files under profiles/:
  python scripts/ncu_summary.py raw  <report.ncu-rep> <out.csv>     selected metrics of every captured launch
  python scripts/ncu_summary.py list <launches.csv>   <out.csv>     per-kernel totals of a `--metrics gpu__time_duration.sum` pass"""
import collections
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [k for k in KEEP if k in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])


def launch_list(src, out):
    rows = list(csv.reader(open(src, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < 15 or r[12] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[4].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r[14])
    tot = sum(a[1] for a in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_listed_time"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], round(a[1] / 1e3, 1), round(a[1] / a[0] / 1e3, 2), round(a[1] / tot, 4)])


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
