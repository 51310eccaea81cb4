#!/bin/bash
# repeat the probe in fresh processes until a failing one shows up (the condition is per process)
mkdir -p gpurun_out
: > gpurun_out/race_loop.jsonl
for i in 1 2 3 4 5 6; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29580 + i)) tools/mg_race_probe.py \
     --users 10000000 --items 2000000 --edges 200000000 --iters 1 --burst 3 --variants default >> gpurun_out/race_loop.jsonl 2> gpurun_out/race_loop_$i.err
  echo "run $i rc=$?" >> gpurun_out/race_loop.jsonl
done
grep -v "^NCCL" gpurun_out/race_loop.jsonl | cut -c1-700
