#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen.log 2>&1; echo "pytest screen rc=$?"; tail -3 gpurun_out/pytest_screen.log
for kl in 40 32 24; do
  for wl in c5 c2; do
    TGCN_EVAL_SCREEN_KL=$kl timeout 600 python tools/screen_probe.py $wl > gpurun_out/screen_probe_${wl}_kl$kl.json 2> gpurun_out/screen_probe_$wl.err; echo "KL=$kl $wl rc=$?"; cat gpurun_out/screen_probe_${wl}_kl$kl.json
  done
done
