# round 2, session 2: full GPU test suite, then the default bench line (Gundam bf16, 1024 pages, one 1024-page lock-step group)
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2c12_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c12_tests.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2c12_clocks.csv &
SMI=$!
timeout 1500 python bench.py --profile-json gpurun_out/r2c12_profile.json > gpurun_out/r2c12_bench.log 2> gpurun_out/r2c12_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c12_bench.err
kill $SMI
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2c12_bench.log") if l.startswith("{")][-1])
    print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), d["stage_ms"], d.get("clocks"))
    print("roofline", {k: d["roofline"].get(k) for k in ("kernel", "frac", "frac_graph", "avg_launch_us", "share_of_pass")})
    print("kv", d.get("kv_cache_compare")); print("b1", d.get("decode_batch1")); print("agree", d.get("token_agreement"))
    print("dsq", {f: {k: v.get(k) for k in ("e2e_pages_per_s", "prefill_tok_s", "decode_tok_s", "decode_batch1", "error")} for f, v in (d.get("dsq") or {}).items()})
    print("cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as ex:
    print("not parsed:", ex)
PY
