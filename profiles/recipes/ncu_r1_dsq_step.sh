cd $GRAFT_REPO_ROOT
export DSOCR_NO_GRAPH=1
CMD="python scripts/bench_dsq.py --primary q4k --tokens 40"
$CMD > gpurun_out/dsq_ncu_plain.log 2>&1 && \
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dsq_fused_gemv -s 1024 -c 5 -o gpurun_out/prof_r1_dsq_gemv $CMD > gpurun_out/dsq_ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/dsq_ncu_full.log | cut -c1-300
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dsq_ -s 1460 -c 150 --csv --log-file gpurun_out/dsq_launches_r1.csv $CMD > gpurun_out/dsq_ncu_list.log 2>&1
echo "list rc=$?"; ls -la gpurun_out | tail -5
