# post_attn with 4 rows per block + bulk stores for the f32 rel-pos products: parity (decoder, EP, full architecture,
# linear, vision), then event timing at 1024 rows per decode step with R = 4 (default) and R = 1
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_linear_gpu.py tests/test_vision_gpu.py tests/test_decoder_batched_gpu.py tests/test_decoder_gpu.py tests/test_expert_parallel_gpu.py tests/test_full_arch_gpu.py -q -m gpu > gpurun_out/r2c14_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c14_tests.log
for R in 4 1; do
  DSOCR_POST_ATTN_ROWS=$R timeout 600 python bench.py --steps 1 --warmup 1 --pages 1024 --batch 1024 --max-new-tokens 48 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c14_profile_r$R.json > gpurun_out/r2c14_bench_r$R.log 2> gpurun_out/r2c14_bench_r$R.err; echo "bench R=$R rc=$?"
done
python - <<'PY'
import json
for tag in ("r4", "r1"):
    try:
        d = json.load(open(f"gpurun_out/r2c14_profile_{tag}.json"))
        ks = {k["name"]: k for k in d["kernels"]}
        print(tag, d["stage_ms"])
        for n in ("decode/post_attn_norm_router_dispatch", "decode/dec_qkv", "decode/dec_o_proj", "decode/moe_expert_gate_up", "decode/moe_expert_down", "decode/moe_combine_norm", "vision/sam_relpos_products", "vision/sam_proj", "vision/sam_fc2"):
            if n in ks: print("  ", n, round(ks[n]["ms"], 1), ks[n]["launches"], round(ks[n]["ms"] / ks[n]["launches"] * 1000, 1), "us")
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
