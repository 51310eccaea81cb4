#!/bin/bash
# after the screened eval path: the whole GPU suite on the release and the debug-assert builds, then the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_release.log 2>&1; echo "pytest release rc=$?"; tail -3 gpurun_out/pytest_gpu_release.log
TGCN_B200_LIB=libtgcn_b200_dbg.so timeout 1500 python -m pytest tests -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model or screen or dropin" > gpurun_out/pytest_gpu_debug.log 2>&1; echo "pytest debug rc=$?"; tail -3 gpurun_out/pytest_gpu_debug.log
timeout 900 python bench.py > gpurun_out/bench_n1_screen.json 2> gpurun_out/bench_n1_screen.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1_screen.json").read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ("metric", "value", "ms_per_step", "gpu_launches", "clocks")}))
print(json.dumps(d["eval"])[:1500])
print(json.dumps(d.get("parity")), json.dumps(d.get("c2", {}).get("eval"))[:600])
PY
