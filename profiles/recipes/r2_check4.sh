# round 2, call: default bench (Gundam 1024 pages, 512-page groups) with profile + DSQ side line, expert token-tile A/B,
# dispatcher test, ncu launch list + full captures of the two dominant decode kernels at steady state
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_dispatch_gpu.py -x -q -m gpu > gpurun_out/r2c4_dispatch.log 2>&1; echo "dispatch rc=$?"; tail -3 gpurun_out/r2c4_dispatch.log
timeout 1500 python bench.py --steps 1 --warmup 1 --profile-json gpurun_out/r2c4_profile.json > gpurun_out/r2c4_bench.log 2> gpurun_out/r2c4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c4_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c4_bench.log").read().strip().splitlines()[-1])
    print("default:", round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["stage_ms"])
    print("roofline:", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3), [(o["kernel"], round(o["frac"], 3)) for o in d["roofline_others"]])
    print("kv:", d.get("kv_cache_compare")); print("dsq:", d.get("dsq")); print("agree:", d.get("token_agreement")); print("b1:", d.get("decode_batch1"))
except Exception as ex:
    print("not parsed:", ex)
PY
for f in 64 128; do
DSOCR_FBN=$f timeout 600 python bench.py --steps 1 --warmup 1 --pages 512 --no-cpu-baseline --no-extras > gpurun_out/r2c4_fbn$f.log 2> gpurun_out/r2c4_fbn$f.err; echo "fbn $f rc=$?"
python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2c4_fbn{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print("fbn", sys.argv[1], round(d["value"], 2), "pages/s", d["stage_ms"]["decode.iterative"], [(k["name"], k["ms"]) for k in d["top_kernels"] if "expert" in k["name"]])
except Exception as ex:
    print("not parsed:", ex)
PY
done
# ncu: launch list of a small pass (plain run first, as the recipe asks), then one full capture per dominant decode kernel
SMALL="--steps 1 --warmup 0 --pages 64 --batch 64 --max-new-tokens 24 --no-cpu-baseline --no-extras"
DSOCR_NO_GRAPH=1 timeout 600 python bench.py $SMALL > gpurun_out/r2c4_small_plain.log 2>&1; echo "small plain rc=$?"
DSOCR_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 6000 --csv --log-file gpurun_out/r2c4_launches.csv python bench.py $SMALL > gpurun_out/r2c4_ncu_list.log 2>&1; echo "ncu list rc=$?"
MID="--steps 1 --warmup 0 --pages 256 --batch 256 --max-new-tokens 160 --no-cpu-baseline --no-extras"
DSOCR_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rope_attn_decode_bulk -s 1500 -c 1 -o gpurun_out/r2c4_attn python bench.py $MID > gpurun_out/r2c4_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
DSOCR_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_sk_kernel -s 1500 -c 2 -o gpurun_out/r2c4_sk python bench.py $MID > gpurun_out/r2c4_ncu_sk.log 2>&1; echo "ncu sk rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
