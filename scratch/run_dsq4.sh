cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_dsq_gpu.py -x -q -m gpu > gpurun_out/dsqf4_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/dsqf4_tests.log
timeout 300 python scripts/bench_dsq.py --primary q4k --tokens 512 > gpurun_out/dsqf4_q4k_512.log 2>&1; tail -1 gpurun_out/dsqf4_q4k_512.log | cut -c1-1500
timeout 300 python scripts/bench_dsq.py --primary q4k --tokens 4096 > gpurun_out/dsqf4_q4k_4096.log 2>&1; tail -1 gpurun_out/dsqf4_q4k_4096.log | cut -c1-420
timeout 300 python scripts/bench_dsq.py --primary q8_0 --tokens 512 > gpurun_out/dsqf4_q8_512.log 2>&1; tail -1 gpurun_out/dsqf4_q8_512.log | cut -c1-420
