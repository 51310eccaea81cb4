#!/bin/bash
# same-box A/B of the CTA-pair eval kernel: c5 and c2 eval legs, pair on / off, twice each
mkdir -p gpurun_out
for rep in 1 2; do
for pair in 1 0; do
  export TGCN_EVAL_PAIR=$pair
  timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c2 --no-e2e --no-extras > gpurun_out/ab_c5_pair${pair}_$rep.json 2> gpurun_out/ab_c5_pair${pair}_$rep.err
  timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-extras > gpurun_out/ab_c2_pair${pair}_$rep.json 2> gpurun_out/ab_c2_pair${pair}_$rep.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_c*_pair*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["eval"]["users_per_s"]), round(d["eval"]["ms"], 3), d["clocks"]["sm_mhz"], d["parity"]["ok"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
