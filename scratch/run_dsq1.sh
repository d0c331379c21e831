set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_dsq_gpu.py -x -q -m gpu > gpurun_out/dsqf_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/dsqf_tests.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_dsq_gpu.py::test_dsq_fused_step_matches_unfused_path" -x -q -m gpu > gpurun_out/dsqf_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/dsqf_sanitizer.log
timeout 600 python scripts/bench_dsq.py --primary q4k --tokens 512 > gpurun_out/dsqf_q4k_512.log 2>&1; tail -1 gpurun_out/dsqf_q4k_512.log | cut -c1-1800
DSOCR_DSQ_UNFUSED=1 timeout 600 python scripts/bench_dsq.py --primary q4k --tokens 512 > gpurun_out/dsqu_q4k_512.log 2>&1; tail -1 gpurun_out/dsqu_q4k_512.log | cut -c1-700
timeout 600 python scripts/bench_dsq.py --primary q4k --tokens 4096 > gpurun_out/dsqf_q4k_4096.log 2>&1; tail -1 gpurun_out/dsqf_q4k_4096.log | cut -c1-1800
timeout 600 python scripts/bench_dsq.py --primary q8_0 --tokens 512 > gpurun_out/dsqf_q8_512.log 2>&1; tail -1 gpurun_out/dsqf_q8_512.log | cut -c1-900
