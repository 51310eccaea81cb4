#!/bin/bash
# experiment: CTA pairs (cta_group::2) in the eval kernel (TGCN_EVAL_PAIR=1)
mkdir -p gpurun_out
export TGCN_EVAL_PAIR=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model" > gpurun_out/pytest_pair.log 2>&1; echo "pytest (pair) rc=$?"; tail -5 gpurun_out/pytest_pair.log
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-c2 --no-e2e > gpurun_out/bench_c5_pair.json 2> gpurun_out/bench_c5_pair.err; echo "bench c5 rc=$?"
timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-extras > gpurun_out/bench_c2_pair.json 2> gpurun_out/bench_c2_pair.err; echo "bench c2 rc=$?"
python - <<'PY'
import json
for f in ("bench_c5_pair", "bench_c2_pair"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, json.dumps(d.get("eval")), json.dumps(d.get("parity")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
