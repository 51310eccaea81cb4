#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu5.log
timeout 600 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2_e.json 2> gpurun_out/bench_c2_e.err; echo "bench c2 rc=$?"
tail -c 400 gpurun_out/bench_c2_e.err
