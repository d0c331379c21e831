# round 2, multi-GPU call (run with gpurun --gpus N): EP on real peers, dispatcher, NCCL baseline, strong-scaling bench
cd $GRAFT_REPO_ROOT
N=${NGPU:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
timeout 300 python -m pytest tests/test_expert_parallel_gpu.py tests/test_dispatch_gpu.py -x -q -m gpu -s > gpurun_out/r2m_ep_tests_$N.log 2>&1; echo "ep/dispatch tests rc=$?"; tail -4 gpurun_out/r2m_ep_tests_$N.log
timeout 600 python scripts/bench_ep.py --gpus $N --pages 128 --tokens 512 > gpurun_out/r2m_ep_bench_$N.log 2> gpurun_out/r2m_ep_bench_$N.err; echo "ep bench rc=$?"; tail -2 gpurun_out/r2m_ep_bench_$N.err; tail -1 gpurun_out/r2m_ep_bench_$N.log | cut -c1-1200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_ep_nccl.py --pages 128 > gpurun_out/r2m_nccl_$N.log 2>&1; echo "nccl baseline rc=$?"; tail -1 gpurun_out/r2m_nccl_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2m_bench_$N.log 2> gpurun_out/r2m_bench_$N.err; echo "bench N=$N rc=$?"; tail -2 gpurun_out/r2m_bench_$N.err
python - "$N" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f"gpurun_out/r2m_bench_{sys.argv[1]}.log") if l.startswith("{")][-1])
    print("bench N=%s:" % sys.argv[1], round(d["value"], 2), "pages/s e2e", round(d["e2e"]["value"], 2), d["scaling"], d["stage_ms"])
except Exception as ex:
    print("not parsed:", ex)
PY
