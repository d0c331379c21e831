# final-code ncu pass: (1) quick parity of the prefill router change, (2) plain run of the profiled command, (3) launch
# list, (4) --set full captures of the final vision GEMM / attention / prefill-attention / router kernels
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_decoder_batched_gpu.py -q -m gpu -x > gpurun_out/r2c20_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c20_tests.log
CMD="python bench.py --steps 1 --warmup 0 --pages 48 --batch 48 --max-new-tokens 6 --no-cpu-baseline --no-extras --agree-pages 0"
timeout 600 $CMD > gpurun_out/r2c20_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2c20_launches.csv $CMD > gpurun_out/r2c20_ncu_list.log 2>&1; echo "launch list rc=$?"
for K in "linear_pair_kernel" "vattn_kernel" "pattn_kernel" "post_attn_kernel|router_kernel" "linear_sk_kernel"; do
  N=$(echo "$K" | tr -c 'a-z_' '_' | cut -c1-24)
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 40 -c 12 -o gpurun_out/r2c20_$N $CMD > gpurun_out/r2c20_ncu_$N.log 2>&1; echo "ncu $K rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -8
