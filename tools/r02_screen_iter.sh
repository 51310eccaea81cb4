#!/bin/bash
# iteration loop for the screened eval path: its tests, then the probe on c2 and c5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen.log 2>&1; echo "pytest screen rc=$?"; tail -3 gpurun_out/pytest_screen.log
for wl in c2 c5; do
  timeout 600 python tools/screen_probe.py $wl > gpurun_out/screen_probe_$wl.json 2> gpurun_out/screen_probe_$wl.err; echo "$wl rc=$?"; cat gpurun_out/screen_probe_$wl.json
done
