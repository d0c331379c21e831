# round 2 final check on a fresh box: full GPU test suite, smoke, the default bench line (Gundam bf16, 1024 pages) with the
# per-kernel breakdown, and the Base line (configs[1]) for reference
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2f_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2f_clocks.csv &
SMI=$!
timeout 1500 python bench.py --profile-json gpurun_out/r2f_profile.json > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2f_bench.err
kill $SMI
timeout 300 python bench.py --mode base --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_base.log 2> gpurun_out/r2f_bench_base.err; echo "bench base rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2f_bench.log", "gpurun_out/r2f_bench_base.log"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["stage_ms"], d.get("clocks"))
        print("  roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "frac_graph", "avg_launch_us", "share_of_pass")})
        print("  b1", d.get("decode_batch1"), "agree", (d.get("token_agreement") or {}).get("agreement"), "kv", d.get("kv_cache_compare"))
        print("  dsq", {q: {k: v.get(k) for k in ("e2e_pages_per_s", "prefill_tok_s", "decode_tok_s", "decode_batch1", "error")} for q, v in (d.get("dsq") or {}).items()})
    except Exception as ex:
        print(f, "not parsed:", ex)
PY
