#!/bin/bash
mkdir -p gpurun_out
for wl in c2 c5; do
  timeout 600 python tools/screen_probe.py $wl > gpurun_out/screen_probe_$wl.json 2> gpurun_out/screen_probe_$wl.err; echo "$wl rc=$?"; cat gpurun_out/screen_probe_$wl.json
  TGCN_EVAL_PAIR=0 timeout 600 python tools/screen_probe.py $wl > gpurun_out/screen_probe_${wl}_single.json 2>> gpurun_out/screen_probe_$wl.err; echo "$wl single rc=$?"; cat gpurun_out/screen_probe_${wl}_single.json
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/screen_launches_c2.csv python tools/screen_probe.py c2 190000 1 > gpurun_out/screen_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = []
with open("gpurun_out/screen_launches_c2.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"][:70], r["Grid Size"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
for r in rows[-40:]:
    print(r)
PY
