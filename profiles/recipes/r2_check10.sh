# bisect: which attention configurations fail with two score stages; the same tests with the single-stage 128-key blocks
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_vision_attention_gpu.py -q -m gpu > gpurun_out/r2c10_tests_ns2.log 2>&1; echo "ns2 rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c10_tests_ns2.log
DSOCR_VATTN_KV128=1 timeout 600 python -m pytest tests/test_vision_attention_gpu.py -q -m gpu > gpurun_out/r2c10_tests_kv128.log 2>&1; echo "kv128 rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c10_tests_kv128.log
