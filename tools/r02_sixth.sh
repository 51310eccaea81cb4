#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py --workload c2 --no-cpu-baseline --no-train --no-e2e > gpurun_out/bench_c2_f.json 2> gpurun_out/bench_c2_f.err; echo "bench c2 rc=$?"
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c2 --no-e2e > gpurun_out/bench_c5_f.json 2> gpurun_out/bench_c5_f.err; echo "bench c5 rc=$?"
