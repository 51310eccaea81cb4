#!/bin/bash
# last call of the round: what the driver runs at round end, on the final tree — GPU suite, smoke(), default bench, reference arm (short)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_last.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_last.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_last.log
timeout 900 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_last.json").read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}), d["e2e"]["ms_per_step"], d["eval"]["users_per_s"], d["parity"]["ok"], d["c2"]["parity"]["ok"])
PY
