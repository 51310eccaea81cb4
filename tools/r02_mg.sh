#!/bin/bash
# round 2, multi-GPU call: usage  bash tools/r02_mg.sh <n_gpus> [ab]
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 10 --warmup 3"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sliced.py -m gpu -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_n$N.log
fi
TGCN_GRID_TIMING=1 timeout 900 $RUN > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "bench n=$N rc=$?"
if [ "$2" = "ab" ]; then
  TGCN_GRID_SIDE_STREAM=0 TGCN_PEER_BARRIER=0 timeout 900 $RUN --no-eval --no-e2e > gpurun_out/bench_c5_n${N}_r01_exchange.json 2> gpurun_out/bench_c5_n${N}_r01_exchange.err; echo "bench (r01 exchange) rc=$?"
fi
tail -c 600 gpurun_out/bench_c5_n$N.err
echo done
