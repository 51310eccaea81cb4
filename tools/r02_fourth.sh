#!/bin/bash
# round 2, fourth GPU call: GPU suite after the metrics / eval-split / layout changes, default bench (LTR with L2-sized splits, 10M-user metrics)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu4.log
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_debug_asserts.log 2>&1; echo "debug-assert pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_debug_asserts.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_c.json 2> gpurun_out/bench_n1_c.err; echo "bench rc=$?"
echo done
