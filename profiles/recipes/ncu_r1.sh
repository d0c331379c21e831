#!/bin/bash
# Round-1 ncu captures (run under gpurun).  Same command line plain first, then under ncu.
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --max-new-tokens 64"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
tail -c 600 gpurun_out/ncu_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r1.csv
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"linear_kernel<__half, 32, 2, 2>" -s 20 -c 2 -o gpurun_out/prof_r1_expert_gate_up $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full1 rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:kv_attention -s 100 -c 2 -o gpurun_out/prof_r1_kv_attention $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vattn_kernel -s 2 -c 2 -o gpurun_out/prof_r1_vattn $CMD > gpurun_out/ncu_full3.log 2>&1
echo "full3 rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null
