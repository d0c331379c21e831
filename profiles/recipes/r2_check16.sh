# incremental asynchronous staging (issued chunk by chunk from the vision loop): parity through the public calls, e2e vs resident
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_boundary_gpu.py tests/test_vision_gpu.py tests/test_cli_server_gpu.py tests/test_dispatch_gpu.py -q -m gpu > gpurun_out/r2c16_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c16_tests.log
python - <<'PY' > gpurun_out/r2c16_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c16_smoke.log
import __graft_entry__ as g
g.smoke()
PY
timeout 900 python bench.py --steps 2 --warmup 1 --pages 512 --batch 512 --max-new-tokens 64 --no-cpu-baseline --no-extras --agree-pages 0 > gpurun_out/r2c16_bench_async.log 2> gpurun_out/r2c16_bench_async.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c16_bench_async.log").read().strip().splitlines()[-1])
    print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms", round(d["ms_per_step"], 1), round(d["e2e"]["ms_per_step"], 1), d["stage_ms"])
except Exception as ex:
    print("not parsed:", ex)
PY
