# prefill router with 16 tokens per block: parity (prefill + decode tests, full architecture); expert token tile 64 vs 128 at
# 1024 rows per step (DSOCR_FBN), event timing
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_decoder_gpu.py tests/test_decoder_batched_gpu.py tests/test_full_arch_gpu.py tests/test_boundary_gpu.py -q -m gpu > gpurun_out/r2c17_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c17_tests.log
for F in 0 64; do
  unset DSOCR_FBN; [ $F != 0 ] && export DSOCR_FBN=$F
  timeout 600 python bench.py --steps 1 --warmup 1 --pages 1024 --batch 1024 --max-new-tokens 48 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c17_profile_fbn$F.json > gpurun_out/r2c17_bench_fbn$F.log 2> gpurun_out/r2c17_bench_fbn$F.err; echo "bench FBN=$F rc=$?"
done
python - <<'PY'
import json
for tag in ("fbn0", "fbn64"):
    try:
        d = json.load(open(f"gpurun_out/r2c17_profile_{tag}.json"))
        ks = {k["name"]: k for k in d["kernels"]}
        print(tag, d["stage_ms"])
        for n in ("decode/post_attn_norm_router_dispatch", "decode/moe_expert_gate_up", "decode/moe_expert_down", "prefill/moe_router", "prefill/moe_dispatch", "prefill/moe_expert_gate_up"):
            if n in ks: print("  ", n, round(ks[n]["ms"], 1), ks[n]["launches"], round(ks[n]["ms"] / ks[n]["launches"] * 1000, 1), "us")
        l = json.loads(open(f"gpurun_out/r2c17_bench_{tag}.log").read().strip().splitlines()[-1])
        print("   value", round(l["value"], 2), "e2e", round(l["e2e"]["value"], 2))
    except Exception as ex:
        print(tag, "not parsed:", ex)
PY
