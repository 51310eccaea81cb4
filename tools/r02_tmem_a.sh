#!/bin/bash
# experiment: user operand of the eval kernel in tensor memory (TGCN_EVAL_TMEM_A=1)
mkdir -p gpurun_out
export TGCN_EVAL_TMEM_A=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model" > gpurun_out/pytest_tmem_a.log 2>&1; echo "pytest (TMEM A) rc=$?"; tail -5 gpurun_out/pytest_tmem_a.log
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c2 --no-e2e > gpurun_out/bench_c5_tmem_a.json 2> gpurun_out/bench_c5_tmem_a.err; echo "bench c5 rc=$?"
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-extras > gpurun_out/bench_c2_tmem_a.json 2> gpurun_out/bench_c2_tmem_a.err; echo "bench c2 rc=$?"
unset TGCN_EVAL_TMEM_A
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c2 --no-e2e > gpurun_out/bench_c5_smem_a.json 2> gpurun_out/bench_c5_smem_a.err; echo "bench c5 (smem A) rc=$?"
