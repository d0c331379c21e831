# residual-add epilogue of the CTA-pair GEMM as bulk reductions (x += y in the memory system): parity + timing
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_vision_gpu.py tests/test_full_arch_gpu.py -q -m gpu > gpurun_out/r2c13_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c13_tests.log
timeout 600 python bench.py --steps 1 --warmup 1 --pages 256 --batch 256 --max-new-tokens 64 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c13_profile.json > gpurun_out/r2c13_bench.log 2> gpurun_out/r2c13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c13_bench.log").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
    print([(k["name"], round(k["ms"], 1), k["launches"]) for k in d["top_kernels"] if k["name"].startswith("vision")])
except Exception as ex:
    print("not parsed:", ex)
PY
