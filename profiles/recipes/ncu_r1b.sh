#!/bin/bash
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --max-new-tokens 6"
$CMD > gpurun_out/ncu_plain_b.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_b.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"linear_kernel<__half, .{0,8}64, .{0,8}2, .{0,8}2>" -s 4 -c 2 -o gpurun_out/prof_r1_expert_gate_up $CMD > gpurun_out/ncu_b1.log 2>&1; echo "b1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rope_attn_decode -s 30 -c 1 -o gpurun_out/prof_r1_rope_attn_decode $CMD > gpurun_out/ncu_b2.log 2>&1; echo "b2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:vattn_kernel -s 2 -c 1 -o gpurun_out/prof_r1_vattn_v2_global $CMD > gpurun_out/ncu_b3.log 2>&1; echo "b3 rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"linear_kernel<__half, .{0,8}256, .{0,8}1, .{0,8}1>" -s 3 -c 1 -o gpurun_out/prof_r1_sam_qkv_gemm $CMD > gpurun_out/ncu_b4.log 2>&1; echo "b4 rc=$?"
ls -la gpurun_out/*.ncu-rep
