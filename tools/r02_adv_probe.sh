#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/adv_step_probe.py > gpurun_out/adv_probe_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/adv_probe_plain.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/adv_launches.csv python tools/adv_step_probe.py > gpurun_out/adv_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
with open("gpurun_out/adv_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
tot = 0
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", "")); tot += v
        print(f'{v/1e3:9.1f} us  {r["Grid Size"]:>16s}  {r["Kernel Name"][:90]}')
print("sum us", tot / 1e3)
PY
