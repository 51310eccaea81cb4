#!/bin/bash
# first run of the screened eval path: its tests, then the eval legs of the bench (auto = screened) for c5 and c2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen.log 2>&1; echo "pytest screen rc=$?"; tail -15 gpurun_out/pytest_screen.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model" > gpurun_out/pytest_screen2.log 2>&1; echo "pytest eval subset rc=$?"; tail -5 gpurun_out/pytest_screen2.log
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c2 --no-e2e --no-extras > gpurun_out/bench_c5_screen.json 2> gpurun_out/bench_c5_screen.err; echo "bench c5 rc=$?"
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-extras > gpurun_out/bench_c2_screen.json 2> gpurun_out/bench_c2_screen.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ("bench_c5_screen", "bench_c2_screen"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e = d["eval"]
        print(f, round(e["users_per_s"]), round(e["ms"], 3), d["clocks"]["sm_mhz"], json.dumps(d.get("parity")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
