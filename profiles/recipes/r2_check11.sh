# pipelined vision attention after the epilogue-wait fix: parity (all attention + vision tests), then per-kernel timing
# with two score stages (default) and with the single-stage 128-key blocks (DSOCR_VATTN_KV128=1)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_vision_attention_gpu.py tests/test_vision_gpu.py -q -m gpu > gpurun_out/r2c11_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c11_tests.log
for V in ns2 kv128 ns2poly3; do
  export DSOCR_VATTN_POLY=0; unset DSOCR_VATTN_KV128
  [ $V = kv128 ] && export DSOCR_VATTN_KV128=1
  [ $V = ns2poly3 ] && export DSOCR_VATTN_POLY=3
  timeout 600 python bench.py --steps 1 --warmup 1 --pages 256 --batch 256 --max-new-tokens 64 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c11_profile_$V.json > gpurun_out/r2c11_bench_$V.log 2> gpurun_out/r2c11_bench_$V.err; echo "bench $V rc=$?"
done
python - <<'PY'
import json
for tag in ("ns2", "kv128", "ns2poly3"):
    try:
        d = json.loads(open(f"gpurun_out/r2c11_bench_{tag}.log").read().strip().splitlines()[-1])
        print(tag, round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
        print([(k["name"], round(k["ms"], 1), k["launches"]) for k in d["top_kernels"] if k["name"].startswith("vision")])
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
