#!/bin/bash
# round 2, third GPU call: SpMM variants after the scheduling fix (packed / unpacked x 1 / 2 quads per step), GPU suite on the
# release and on the debug-assert build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu3.log
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_debug_asserts.log 2>&1; echo "debug-assert pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_debug_asserts.log
C2="--workload c2 --steps 30 --no-cpu-baseline --no-torch-ref --no-extras --no-eval --no-e2e"
for v in "packed1:TGCN_SPMM_QUADS_D64=1" "packed2:TGCN_SPMM_QUADS_D64=2" "unpacked1:TGCN_SPMM_PACKED=0 TGCN_SPMM_QUADS_D64=1" "unpacked2:TGCN_SPMM_PACKED=0 TGCN_SPMM_QUADS_D64=2"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py $C2 > gpurun_out/ab3_c2_$name.json 2> gpurun_out/ab3_c2_$name.err
done
C5="--steps 5 --no-cpu-baseline --no-torch-ref --no-c2 --no-eval --no-e2e"
for v in "packed1:TGCN_SPMM_QUADS_D128=1" "packed2:TGCN_SPMM_QUADS_D128=2" "unpacked1:TGCN_SPMM_PACKED=0 TGCN_SPMM_QUADS_D128=1" "unpacked2:TGCN_SPMM_PACKED=0 TGCN_SPMM_QUADS_D128=2"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py $C5 > gpurun_out/ab3_c5_$name.json 2> gpurun_out/ab3_c5_$name.err
done
echo done
