# what bounds the prefill router kernel (three forms, same ~390 us)?  one ncu source capture from a 40-page bench run
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 1 --warmup 0 --pages 40 --batch 40 --max-new-tokens 3 --no-cpu-baseline --no-extras --agree-pages 0"
mkdir -p /tmp/ncu
timeout 300 $CMD > gpurun_out/r2c23_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:router_kernel -s 2 -c 1 -o /tmp/ncu/router $CMD > gpurun_out/r2c23_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_extract.py /tmp/ncu/router.ncu-rep gpurun_out/r02_final_prefill_router 50 | tail -1
rm -rf /tmp/ncu
