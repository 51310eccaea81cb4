#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 tools/mg_race_probe.py ${@:2} > gpurun_out/race_probe_n$N.jsonl 2> gpurun_out/race_probe_n$N.err; echo "rc=$?"
cat gpurun_out/race_probe_n$N.jsonl; tail -c 1500 gpurun_out/race_probe_n$N.err
