#!/bin/bash
# which switch of the r01 exchange breaks parity at N = 2?  (c2 is quick; c5 reproduces the failure seen in bench_c5_n2_r01_exchange)
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 10 --warmup 3 --no-eval --no-e2e"
for wl in c2 c5; do
for v in "default:" "noside:TGCN_GRID_SIDE_STREAM=0" "nopeerbar:TGCN_PEER_BARRIER=0" "both:TGCN_GRID_SIDE_STREAM=0 TGCN_PEER_BARRIER=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 $RUN --workload $wl > gpurun_out/dbg_${wl}_$name.json 2> gpurun_out/dbg_${wl}_$name.err; echo "$wl $name rc=$?"
done
done
echo done
