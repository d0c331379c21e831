# round 2, call 1: new parity tests (batched decode, full architecture, boundary), smoke, short Gundam bench
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests -x -q -m gpu -s --durations=15 > gpurun_out/r2c1_tests.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2c1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c1_smoke.log
timeout 900 python bench.py --steps 1 --warmup 1 --pages 256 --profile-json gpurun_out/r2c1_profile.json > gpurun_out/r2c1_bench.log 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c1_bench.err; cat gpurun_out/r2c1_bench.log
