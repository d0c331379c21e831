cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_dsq_gpu.py tests/test_decoder_gpu.py -x -q -m gpu > gpurun_out/small1_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/small1_tests.log
timeout 300 python scripts/bench_dsq.py --primary float --dtype bf16 --tokens 512 > gpurun_out/small1_float_512.log 2>&1; tail -1 gpurun_out/small1_float_512.log | cut -c1-1500
DSOCR_NO_SMALL_FUSED=1 timeout 300 python scripts/bench_dsq.py --primary float --dtype bf16 --tokens 512 > gpurun_out/small1_float_512_old.log 2>&1; tail -1 gpurun_out/small1_float_512_old.log | cut -c1-1200
timeout 300 python scripts/bench_dsq.py --primary q4k --tokens 512 > gpurun_out/small1_q4k_512.log 2>&1; tail -1 gpurun_out/small1_q4k_512.log | cut -c1-420
