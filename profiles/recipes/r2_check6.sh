# round 2: vectorised dequant producer + split-K for 257..1024-row decode steps: parity, then timing
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_linear_dq_gpu.py tests/test_dsq_gpu.py tests/test_decoder_batched_gpu.py tests/test_decoder_gpu.py -x -q -m gpu > gpurun_out/r2c6_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c6_tests.log
timeout 900 python bench.py --steps 1 --warmup 1 --pages 512 --no-cpu-baseline --dsq-formats q4k,q8_0 --profile-json gpurun_out/r2c6_profile.json > gpurun_out/r2c6_bench.log 2> gpurun_out/r2c6_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c6_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c6_bench.log").read().strip().splitlines()[-1])
    print("512 pages:", round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
    print([(k["name"], k["ms"], k["launches"]) for k in d["top_kernels"]])
    for f, v in (d.get("dsq") or {}).items():
        print("dsq", f, {k: v.get(k) for k in ("export_s", "e2e_pages_per_s", "prefill_tok_s", "decode_tok_s", "decode_batch1", "error")}, v.get("stage_ms"))
except Exception as ex:
    print("not parsed:", ex)
PY
