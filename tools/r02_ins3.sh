#!/bin/bash
# RECORD of an experiment that was not adopted: the 3xTF32 variants with the lists on inserter warps.  The TGCN_EVAL_INS switch this
# script drives was removed from eval_tc.cu together with the instantiations (profiles/r02/README.md has the numbers).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model or screen" > gpurun_out/pytest_ins3.log 2>&1; echo "pytest eval subset rc=$?"; tail -3 gpurun_out/pytest_ins3.log
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py -m gpu -x -q -k "topk or tf32 or screen" > gpurun_out/pytest_ins3_dbg.log 2>&1; echo "pytest (debug build) rc=$?"; tail -3 gpurun_out/pytest_ins3_dbg.log
for ins in 1 0; do
  for wl in c2 c5; do
    TGCN_EVAL_INS=$ins timeout 600 python tools/screen_probe.py $wl > gpurun_out/probe_${wl}_ins3_$ins.json 2> gpurun_out/screen_probe_$wl.err; echo "TGCN_EVAL_INS=$ins $wl rc=$?"; cat gpurun_out/probe_${wl}_ins3_$ins.json
  done
done
