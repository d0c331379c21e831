#!/bin/bash
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --max-new-tokens 260"
$CMD > gpurun_out/ncu_plain_f.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_f.log; exit 1; }
# a launch late in the first timed step: context ~ 283 + 250 tokens
ncu --set full --clock-control none --import-source on -k regex:rope_attn_decode_bulk -s 3000 -c 1 -o gpurun_out/prof_r1_attn_bulk $CMD > gpurun_out/ncu_f1.log 2>&1; echo "f1 rc=$?"
ls -la gpurun_out/prof_r1_attn_bulk.ncu-rep
