# One gpurun call that re-establishes the state of the repo on a fresh B200 (about 6 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/recipes/round_check.sh'
# Keeps gpurun_out/ small (logs only; an ncu report of more than a handful of kernels exceeds the 64 MiB return limit).
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/check_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/check_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/check_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/check_smoke.log
timeout 600 python bench.py > gpurun_out/check_bench.log 2> gpurun_out/check_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/check_bench.log").read().strip().splitlines()[-1])
    print("bench:", round(d["value"], 1), "pages/s, e2e", round(d["e2e"]["value"], 1), "| roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
          "| batch-1", d.get("decode_batch1"))
except Exception as ex:
    print("bench line not parsed:", ex)
PY
for p in q4k q8_0 float; do
  timeout 300 python scripts/bench_dsq.py --primary $p --tokens 512 > gpurun_out/check_b1_$p.log 2>&1
  python - "$p" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/check_b1_{sys.argv[1]}.log") if l.startswith("{")][-1])
    print("batch-1", sys.argv[1], round(d["decode_tok_s"]), "tok/s", d["launches_per_token"], "launches/token")
except Exception as ex:
    print("batch-1", sys.argv[1], "failed:", ex)
PY
done
