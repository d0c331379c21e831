cd $GRAFT_REPO_ROOT
timeout 100 python scripts/bench_dsq.py --primary q4k --tokens 512 > gpurun_out/last_q4k_512.log 2>&1; tail -1 gpurun_out/last_q4k_512.log | cut -c1-1400
