#!/bin/bash
mkdir -p gpurun_out
WL=${1:-c2}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'eval_topk|tf32_split|topk_merge|max_row|gather_rows|fb_index' -c 40 --csv --log-file gpurun_out/screen_launches_$WL.csv python tools/screen_probe.py $WL ${2:-190000} 1 > gpurun_out/screen_ncu.log 2>&1; echo "ncu rc=$?"
python - $WL <<'PY'
import csv, sys
with open(f"gpurun_out/screen_launches_{sys.argv[1]}.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        print(r["Kernel Name"][:90], r["Grid Size"], r["Block Size"], r["Metric Value"], r["Metric Unit"])
PY
