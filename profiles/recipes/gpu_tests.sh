cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/full2_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/full2_tests.log
