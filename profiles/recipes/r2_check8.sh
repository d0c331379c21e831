# round 2, session 2: where does the single-pass vision attention kernel wait?  ncu source-level capture of a GPU-filling
# launch (64-grid global attention, 4 images x 12 heads = 1536 CTAs), plus the launch / PDL / grid-barrier micro-benchmark
cd $GRAFT_REPO_ROOT
timeout 120 deepseek-ocr.rs_b200/build/microbench_step > gpurun_out/r2c8_microbench.log 2>&1; echo "microbench rc=$?"; cat gpurun_out/r2c8_microbench.log
timeout 300 python scripts/vattn_probe.py --grid 64 --B 4 --H 12 > gpurun_out/r2c8_probe.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vattn_kernel -c 1 -o gpurun_out/r2c8_vattn64 python scripts/vattn_probe.py --grid 64 --B 4 --H 12 > gpurun_out/r2c8_ncu64.log 2>&1; echo "ncu64 rc=$?"; tail -3 gpurun_out/r2c8_ncu64.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vattn_kernel -c 1 -o gpurun_out/r2c8_vattn14 python scripts/vattn_probe.py --grid 14 --B 600 --H 12 > gpurun_out/r2c8_ncu14.log 2>&1; echo "ncu14 rc=$?"; tail -3 gpurun_out/r2c8_ncu14.log
