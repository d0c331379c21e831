#!/bin/bash
# targeted captures of the current decode kernels (run only after the plain command exited 0)
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --max-new-tokens 24"
$CMD > gpurun_out/ncu_plain_c.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_c.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:linear_sk_kernel -s 60 -c 2 -o gpurun_out/prof_r1_expert_sk $CMD > gpurun_out/ncu_c1.log 2>&1; echo "c1 rc=$?"
ls -la gpurun_out/*.ncu-rep
