# persistent vision-attention CTAs (items walked by 2 x 148 CTAs, pipelines on a global key-block counter): parity, then
# per-kernel timing against one CTA per item (DSOCR_VATTN_ONE_ITEM=1), same build
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_vision_attention_gpu.py tests/test_vision_gpu.py -q -m gpu > gpurun_out/r2c18_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c18_tests.log
for V in persistent one_item; do
  unset DSOCR_VATTN_ONE_ITEM; [ $V = one_item ] && export DSOCR_VATTN_ONE_ITEM=1
  timeout 600 python bench.py --steps 1 --warmup 1 --pages 256 --batch 256 --max-new-tokens 32 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c18_profile_$V.json > gpurun_out/r2c18_bench_$V.log 2> gpurun_out/r2c18_bench_$V.err; echo "bench $V rc=$?"
done
python - <<'PY'
import json
for tag in ("persistent", "one_item"):
    try:
        d = json.load(open(f"gpurun_out/r2c18_profile_{tag}.json"))
        ks = {k["name"]: k for k in d["kernels"]}
        print(tag, d["stage_ms"])
        for n in ("vision/sam_global_attention", "vision/sam_window_attention", "vision/clip_attention"):
            if n in ks: print("  ", n, round(ks[n]["ms"], 1), ks[n]["launches"], round(ks[n]["ms"] / ks[n]["launches"] * 1000, 1), "us")
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
