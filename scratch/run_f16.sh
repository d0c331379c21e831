cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_decoder_gpu.py -x -q -m gpu -k "f16_engine" -s > gpurun_out/f16_tests.log 2>&1; echo "pytest rc=$?"; grep "parity\|passed\|failed\|Error" gpurun_out/f16_tests.log | tail -8
timeout 120 python scripts/bench_dsq.py --primary float --dtype f16 --tokens 256 > gpurun_out/small2_f16_256.log 2>&1; tail -1 gpurun_out/small2_f16_256.log | cut -c1-400
