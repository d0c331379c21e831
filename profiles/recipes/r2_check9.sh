# round 2, session 2: pipelined vision attention (two score stages in TMEM, K/V ring of 3-4, Zh terms prefetched, hinted
# mbarrier waits): parity, per-kernel timing A/B (DSOCR_VATTN_POLY=0/3), ncu source capture of a GPU-filling launch
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_vision_attention_gpu.py tests/test_vision_gpu.py -x -q -m gpu > gpurun_out/r2c9_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c9_tests.log
for POLY in 0 3; do
  DSOCR_VATTN_POLY=$POLY timeout 600 python bench.py --steps 1 --warmup 1 --pages 256 --batch 256 --max-new-tokens 64 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c9_profile_poly$POLY.json > gpurun_out/r2c9_bench_poly$POLY.log 2> gpurun_out/r2c9_bench_poly$POLY.err; echo "bench poly$POLY rc=$?"; tail -2 gpurun_out/r2c9_bench_poly$POLY.err
done
python - <<'PY'
import json
for tag in ("poly0", "poly3"):
    try:
        d = json.loads(open(f"gpurun_out/r2c9_bench_{tag}.log").read().strip().splitlines()[-1])
        print(tag, round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
        print([(k["name"], round(k["ms"], 1), k["launches"]) for k in d["top_kernels"]])
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
timeout 300 python scripts/vattn_probe.py --grid 64 --B 4 --H 12 > gpurun_out/r2c9_probe.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vattn_kernel -c 1 -o gpurun_out/r2c9_vattn64 python scripts/vattn_probe.py --grid 64 --B 4 --H 12 > gpurun_out/r2c9_ncu64.log 2>&1; echo "ncu64 rc=$?"; tail -2 gpurun_out/r2c9_ncu64.log
