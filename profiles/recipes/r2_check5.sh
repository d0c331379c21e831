# round 2: expert-parallel group on one GPU (functional), dispatcher
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_expert_parallel_gpu.py tests/test_dispatch_gpu.py -x -q -m gpu -s > gpurun_out/r2c5_ep.log 2>&1; echo "ep tests rc=$?"; tail -15 gpurun_out/r2c5_ep.log
timeout 300 python scripts/bench_ep.py --gpus 2 --same-device --config tiny --pages 8 --tokens 64 --prompt-image-tokens 50 > gpurun_out/r2c5_ep_bench_tiny.log 2>&1; echo "ep bench tiny rc=$?"; tail -3 gpurun_out/r2c5_ep_bench_tiny.log | cut -c1-900
