#!/bin/bash
# experiment: screened eval with the lists on dedicated inserter warps (TGCN_EVAL_SCREEN_INS=1)
mkdir -p gpurun_out
TGCN_EVAL_SCREEN_INS=1 timeout 600 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen_ins.log 2>&1; echo "pytest screen (INS) rc=$?"; tail -4 gpurun_out/pytest_screen_ins.log
for ins in 1 0; do
  for wl in c5 c2; do
    TGCN_EVAL_SCREEN_INS=$ins timeout 600 python tools/screen_probe.py $wl > gpurun_out/screen_probe_${wl}_ins$ins.json 2> gpurun_out/screen_probe_$wl.err; echo "INS=$ins $wl rc=$?"; cat gpurun_out/screen_probe_${wl}_ins$ins.json
  done
done
