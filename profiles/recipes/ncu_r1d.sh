#!/bin/bash
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --pages 16 --max-new-tokens 2"
$CMD > gpurun_out/ncu_plain_d.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_d.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:linear_pair_kernel -s 3 -c 2 -o gpurun_out/prof_r1_pair_fc1_fc2 $CMD > gpurun_out/ncu_d1.log 2>&1; echo "d1 rc=$?"
ls -la gpurun_out/prof_r1_pair*.ncu-rep
