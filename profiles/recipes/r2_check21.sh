# lean final-code ncu pass (reports are converted to small CSVs ON THE BOX and deleted: gpurun_out must stay under 64 MiB):
# decoder parity for the prefill-router change first, then one --set full capture per kernel from GPU-filling probes
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_decoder_gpu.py tests/test_decoder_batched_gpu.py -q -m gpu > gpurun_out/r2c21_tests.log 2>&1; echo "tests rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2c21_tests.log
mkdir -p /tmp/ncu
cap() {  # name regex command...
  local name=$1 rx=$2; shift 2
  timeout 300 "$@" > gpurun_out/r2c21_plain_$name.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -c 1 -o /tmp/ncu/$name "$@" > gpurun_out/r2c21_ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  [ -f /tmp/ncu/$name.ncu-rep ] && python scripts/ncu_extract.py /tmp/ncu/$name.ncu-rep gpurun_out/r02_final_$name 40 | tail -1
}
for v in fc1 fc2 proj qkv relpos; do cap pair_$v linear_pair_kernel python scripts/pair_probe.py $v; done
cap vattn64 vattn_kernel python scripts/vattn_probe.py --grid 64 --B 4 --H 12
cap vattn40 vattn_kernel python scripts/vattn_probe.py --grid 40 --B 12 --H 12
cap vattn14 vattn_kernel python scripts/vattn_probe.py --grid 14 --B 600 --H 12
cap pattn pattn_kernel python bench.py --steps 1 --warmup 0 --pages 8 --batch 8 --max-new-tokens 3 --no-cpu-baseline --no-extras --agree-pages 0
rm -rf /tmp/ncu; du -sh gpurun_out
