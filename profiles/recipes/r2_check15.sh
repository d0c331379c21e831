# pages staged on a side stream under the vision tower (decode_pages / decode_requests), post_attn rows per block 1/2/4/8:
# parity through the public calls, then e2e vs resident with the overlap on (default) and off (DSOCR_SYNC_STAGING=1)
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_boundary_gpu.py tests/test_vision_gpu.py tests/test_cli_server_gpu.py tests/test_dispatch_gpu.py tests/test_full_arch_gpu.py tests/test_decoder_batched_gpu.py -q -m gpu > gpurun_out/r2c15_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c15_tests.log
python - <<'PY' > gpurun_out/r2c15_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c15_smoke.log
import __graft_entry__ as g
g.smoke()
PY
for V in async sync; do
  unset DSOCR_SYNC_STAGING; [ $V = sync ] && export DSOCR_SYNC_STAGING=1
  timeout 900 python bench.py --steps 2 --warmup 1 --pages 512 --batch 512 --max-new-tokens 64 --no-cpu-baseline --no-extras --agree-pages 0 > gpurun_out/r2c15_bench_$V.log 2> gpurun_out/r2c15_bench_$V.err; echo "bench $V rc=$?"
done
python - <<'PY'
import json
for tag in ("async", "sync"):
    try:
        d = json.loads(open(f"gpurun_out/r2c15_bench_{tag}.log").read().strip().splitlines()[-1])
        print(tag, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "ms", round(d["ms_per_step"], 1), round(d["e2e"]["ms_per_step"], 1), d["stage_ms"])
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
