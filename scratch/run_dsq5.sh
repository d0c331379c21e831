cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_dsq_gpu.py -x -q -m gpu > gpurun_out/dsqf5_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/dsqf5_tests.log
for k in 8 16; do
DSOCR_DSQ_LPR_K=$k timeout 300 python scripts/bench_dsq.py --primary q4k --tokens 512 > gpurun_out/dsqf5_q4k_512_k$k.log 2>&1; echo "LPR_K=$k"; tail -1 gpurun_out/dsqf5_q4k_512_k$k.log | cut -c1-1300
done
for k in 16 32; do
DSOCR_DSQ_LPR_8=$k timeout 300 python scripts/bench_dsq.py --primary q8_0 --tokens 512 > gpurun_out/dsqf5_q8_512_k$k.log 2>&1; echo "LPR_8=$k"; tail -1 gpurun_out/dsqf5_q8_512_k$k.log | cut -c1-1300
done
