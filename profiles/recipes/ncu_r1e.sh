#!/bin/bash
# launch list of a shortened bench step (4 decode tokens instead of 512: the full command launches 93k kernels)
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --max-new-tokens 4"
$CMD > gpurun_out/ncu_plain_e.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_e.log; exit 1; }
timeout 1100 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r1_final.csv $CMD > gpurun_out/ncu_e1.log 2>&1; echo "e1 rc=$?"
wc -l gpurun_out/launches_r1_final.csv
