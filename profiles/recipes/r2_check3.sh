# round 2, call: dequant-fused GEMM (kernel-level + engine), CLI/server on the engine, full suite, batch-1024 experiment
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_linear_dq_gpu.py -x -q -m gpu -s > gpurun_out/r2c3_dq.log 2>&1; echo "linear_dq rc=$?"; grep -c "parity" gpurun_out/r2c3_dq.log; tail -4 gpurun_out/r2c3_dq.log
timeout 900 python -m pytest tests/test_dsq_gpu.py -x -q -m gpu -s > gpurun_out/r2c3_dsq.log 2>&1; echo "dsq rc=$?"; grep "timing\|passed\|failed" gpurun_out/r2c3_dsq.log | tail -5
timeout 900 python -m pytest tests/test_cli_server_gpu.py -x -q -m gpu > gpurun_out/r2c3_cli.log 2>&1; echo "cli/server rc=$?"; tail -4 gpurun_out/r2c3_cli.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c3_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c3_tests.log
timeout 900 python bench.py --steps 1 --warmup 1 --pages 1024 --batch 1024 --no-cpu-baseline --no-extras > gpurun_out/r2c3_bench_1024.log 2> gpurun_out/r2c3_bench_1024.err; echo "bench 1024 rc=$?"; tail -2 gpurun_out/r2c3_bench_1024.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c3_bench_1024.log").read().strip().splitlines()[-1])
    print("batch 1024:", round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
except Exception as ex:
    print("not parsed:", ex)
PY
