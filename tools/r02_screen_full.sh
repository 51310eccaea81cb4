#!/bin/bash
# one full ncu capture of the screened eval kernel (first launch) on a bench workload; text exports only (the report stays on the box)
mkdir -p gpurun_out
WL=${1:-c2}; NU=${2:-37888}
timeout 600 python tools/screen_probe.py $WL $NU 1 > gpurun_out/screen_probe_small.json 2>&1; echo "plain rc=$?"; cat gpurun_out/screen_probe_small.json
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'eval_topk_tc_kernel' -c 1 -o /tmp/screen_$WL python tools/screen_probe.py $WL $NU 1 > gpurun_out/screen_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/screen_$WL.ncu-rep --page details > gpurun_out/screen_${WL}_details.txt 2>&1
ncu -i /tmp/screen_$WL.ncu-rep --page source --csv > gpurun_out/screen_${WL}_source.csv 2>&1
ls -la gpurun_out/screen_${WL}_*
