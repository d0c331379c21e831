#!/usr/bin/env python
"""Turns one `.ncu-rep` capture (ncu --set full --import-source on) into the two small CSVs kept under profiles/:
  <out>_raw.csv     selected raw metrics of every captured launch (duration, DRAM bytes, pipe utilisation, occupancy)
  <out>_stalls.csv  warp-stall samples per SASS instruction, the N most sampled instructions with their top reasons
Needs no GPU.  Usage: python scripts/ncu_extract.py gpurun_out/x.ncu-rep profiles/r02_x [top_n]"""
import csv
import io
import re
import subprocess
import sys

PAT = (r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
       r"|sm__inst_executed_pipe_(xu|alu|fma|lsu|tmem|uniform)\.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct|sm__warps_active.avg.per_cycle_active"
       r"|smsp__inst_executed.sum$|sm__cycles_elapsed.avg$|launch__(grid_size|block_size|registers_per_thread$|shared_mem_per_block_dynamic|occupancy_limit)"
       r"|smsp__average_warps_issue_stalled_.*_per_issue_active|l1tex__data_pipe_(lsu|tc)_wavefronts_mem_shared.sum.pct|lts__t_sector_hit_rate.pct|sm__throughput.avg.pct")


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or re.search(PAT, h)]
    with open(out + "_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch {r[0]}" for r in data])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in data])
    src = list(csv.reader(io.StringIO(page(rep, "source"))))
    # the source page repeats a "Kernel Name" line per launch; keep the first launch
    start = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    end = next((i for i in range(start + 1, len(src)) if src[i] and src[i][0] == "Kernel Name"), len(src))
    h = src[start]
    ix = {n: i for i, n in enumerate(h)}
    stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    body = [r for r in src[start + 1:end] if len(r) == len(h)]
    total = sum(int(r[ix["# Samples"]]) for r in body)
    body.sort(key=lambda r: -int(r[ix["# Samples"]]))
    with open(out + "_stalls.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", src[0][1] if src and len(src[0]) > 1 else ""])
        w.writerow(["total_samples", total])
        agg = {n: sum(int(r[ix[n]]) for r in body) for n in stall}
        w.writerow(["all_instructions"] + [f"{n[6:]}={v}" for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v])
        w.writerow(["samples", "executed", "sass", "top_reasons"])
        for r in body[:top_n]:
            st = sorted(((n[6:], int(r[ix[n]])) for n in stall if int(r[ix[n]])), key=lambda kv: -kv[1])[:3]
            w.writerow([r[ix["# Samples"]], r[ix["Instructions Executed"]], " ".join(r[ix["Source"]].split()), " ".join(f"{a}={b}" for a, b in st)])
    print("wrote", out + "_raw.csv", out + "_stalls.csv", "samples", total)


if __name__ == "__main__":
    main()
