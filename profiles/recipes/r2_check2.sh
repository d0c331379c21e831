# round 2, call: tcgen05 prefill attention (parity + timing A/B), fused decode schedule beyond 256 pages (experiment)
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c2_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c2_tests.log
DSOCR_FUSED_MAX_ROWS=512 timeout 600 python -m pytest tests/test_decoder_batched_gpu.py -x -q -m gpu -k "large_decode" -s > gpurun_out/r2c2_fused512.log 2>&1; echo "fused512 rc=$?"; grep "parity\|passed\|failed\|Error" gpurun_out/r2c2_fused512.log | tail -8
timeout 900 python bench.py --steps 1 --warmup 1 --pages 256 --no-cpu-baseline --no-extras --profile-json gpurun_out/r2c2_profile.json > gpurun_out/r2c2_bench.log 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c2_bench.err
python - <<'PY'
import json
for f in ("r2c2_bench",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"], [(k["name"], k["ms"]) for k in d["top_kernels"][:8]])
    except Exception as ex:
        print(f, "not parsed:", ex)
PY
DSOCR_PREFILL_SIMT=1 timeout 900 python bench.py --steps 1 --warmup 1 --pages 256 --no-cpu-baseline --no-extras > gpurun_out/r2c2_bench_simt.log 2> gpurun_out/r2c2_bench_simt.err; echo "bench simt rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c2_bench_simt.log").read().strip().splitlines()[-1])
    print("simt prefill:", round(d["value"], 2), "pages/s", d["stage_ms"])
except Exception as ex:
    print("not parsed:", ex)
PY
DSOCR_FUSED_MAX_ROWS=512 timeout 900 python bench.py --steps 1 --warmup 1 --pages 512 --batch 512 --no-cpu-baseline --no-extras > gpurun_out/r2c2_bench_512.log 2> gpurun_out/r2c2_bench_512.err; echo "bench 512 rc=$?"; tail -2 gpurun_out/r2c2_bench_512.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c2_bench_512.log").read().strip().splitlines()[-1])
    print("batch 512:", round(d["value"], 2), "pages/s", d["stage_ms"], [(k["name"], k["ms"]) for k in d["top_kernels"][:8]])
except Exception as ex:
    print("not parsed:", ex)
PY
