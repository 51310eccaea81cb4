#!/bin/bash
mkdir -p gpurun_out
TGCN_EVAL_SCREEN_BN=128 timeout 900 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen_bn128.log 2>&1; echo "pytest screen (BN=128, 4 acc stages) rc=$?"; tail -3 gpurun_out/pytest_screen_bn128.log
for bn in 128 256; do
  for wl in c5 c2; do
    TGCN_EVAL_SCREEN_BN=$bn timeout 600 python tools/screen_probe.py $wl > gpurun_out/screen_probe_${wl}_bn$bn.json 2> gpurun_out/screen_probe_$wl.err; echo "BN=$bn $wl rc=$?"; cat gpurun_out/screen_probe_${wl}_bn$bn.json
  done
done
