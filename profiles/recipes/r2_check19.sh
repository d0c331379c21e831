# where does a window-attention item spend its ~10 us?  ncu source capture of the persistent kernel, 600 windows x 12 heads
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/vattn_probe.py --grid 14 --B 600 --H 12 > gpurun_out/r2c19_probe.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vattn_kernel -c 1 -o gpurun_out/r2c19_vattn14 python scripts/vattn_probe.py --grid 14 --B 600 --H 12 > gpurun_out/r2c19_ncu14.log 2>&1; echo "ncu14 rc=$?"; tail -2 gpurun_out/r2c19_ncu14.log
