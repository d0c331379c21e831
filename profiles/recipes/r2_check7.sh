# round 2, session 2: single-pass vision attention (tail-specialised, TMEM loads one chunk ahead, optional FMA-pipe exp2):
# parity first, then per-kernel timing A/B (DSOCR_VATTN_POLY=0/3), then the 1024-page lock-step group
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_vision_attention_gpu.py tests/test_vision_gpu.py -x -q -m gpu > gpurun_out/r2c7_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c7_tests.log
for POLY in 3 0; do
  DSOCR_VATTN_POLY=$POLY timeout 600 python bench.py --steps 1 --warmup 1 --pages 256 --batch 256 --max-new-tokens 64 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c7_profile_poly$POLY.json > gpurun_out/r2c7_bench_poly$POLY.log 2> gpurun_out/r2c7_bench_poly$POLY.err; echo "bench poly$POLY rc=$?"; tail -2 gpurun_out/r2c7_bench_poly$POLY.err
done
timeout 900 python bench.py --steps 1 --warmup 1 --pages 1024 --batch 1024 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c7_profile_b1024.json > gpurun_out/r2c7_bench_b1024.log 2> gpurun_out/r2c7_bench_b1024.err; echo "bench b1024 rc=$?"; tail -2 gpurun_out/r2c7_bench_b1024.err
python - <<'PY'
import json
for tag in ("poly3", "poly0", "b1024"):
    try:
        d = json.loads(open(f"gpurun_out/r2c7_bench_{tag}.log").read().strip().splitlines()[-1])
        print(tag, round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
        print([(k["name"], round(k["ms"], 1), k["launches"]) for k in d["top_kernels"]])
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
