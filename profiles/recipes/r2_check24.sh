# prefill router with cp.async token staging and a register ring for the gate weights: parity (prefill + decode, full
# architecture), then event timing of prefill/moe_router
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_decoder_batched_gpu.py tests/test_full_arch_gpu.py -q -m gpu > gpurun_out/r2c24_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c24_tests.log
timeout 600 python bench.py --steps 1 --warmup 1 --pages 128 --batch 128 --max-new-tokens 16 --no-cpu-baseline --no-extras --agree-pages 0 --profile-json gpurun_out/r2c24_profile.json > gpurun_out/r2c24_bench.log 2> gpurun_out/r2c24_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c24_profile.json"))
ks = {k["name"]: k for k in d["kernels"]}
print(d["stage_ms"])
for n in ("prefill/moe_router", "prefill/moe_dispatch", "prefill/moe_combine", "prefill/moe_expert_gate_up"):
    print("  ", n, round(ks[n]["ms"], 1), ks[n]["launches"], round(ks[n]["ms"] / ks[n]["launches"] * 1000, 1), "us")
PY
