#!/bin/bash
# round 2, first GPU call: the whole GPU suite, the default bench (N = 1) with its parity block, the reference arm, A/B probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
TGCN_SPMM_LANES_D64=8 timeout 300 python bench.py --workload c2 --steps 20 --no-cpu-baseline --no-torch-ref --no-extras --no-train > gpurun_out/bench_c2_lanes8.json 2> gpurun_out/bench_c2_lanes8.err
timeout 300 python bench.py --workload c2 --steps 20 --no-cpu-baseline --no-torch-ref --no-extras --no-train > gpurun_out/bench_c2_lanes16.json 2> gpurun_out/bench_c2_lanes16.err
TGCN_SPMM_LANES_D128=16 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-torch-ref --no-c2 --no-eval --no-e2e > gpurun_out/bench_c5_lanes16.json 2> gpurun_out/bench_c5_lanes16.err
timeout 600 python tools/hotcold_probe.py 24 48 96 > gpurun_out/hotcold.jsonl 2> gpurun_out/hotcold.err
echo done
