#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ltr_screen_probe.py > gpurun_out/ltr_screen_probe.json 2> gpurun_out/ltr_screen_probe.err; echo rc=$?; cat gpurun_out/ltr_screen_probe.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'eval_topk|tf32_split|topk_merge|screen_prep|gather_rows|fb_' -c 40 --csv --log-file gpurun_out/ltr_launches.csv python tools/ltr_screen_probe.py > gpurun_out/ltr_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
with open("gpurun_out/ltr_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        print(r["Kernel Name"][:80], r["Grid Size"], r["Block Size"], r["Metric Value"], r["Metric Unit"])
PY
