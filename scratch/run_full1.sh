cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/full1_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/full1_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/full1_smoke.log
export DSOCR_NO_GRAPH=1
CMD="python scripts/bench_dsq.py --primary q4k --tokens 40"
$CMD > gpurun_out/dsq_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dsq_ -s 1500 -c 400 --csv --log-file gpurun_out/dsq_launches_r1.csv $CMD > gpurun_out/dsq_ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dsq_ -s 1500 -c 80 -o gpurun_out/prof_r1_dsq_step $CMD > gpurun_out/dsq_ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/dsq_ncu_full.log
