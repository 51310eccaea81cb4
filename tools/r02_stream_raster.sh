#!/bin/bash
# streamed eval variant (LTR ranking): split-fastest rasterisation + enough item splits that concurrent CTAs share L2-resident user tiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_dropin.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model or screen" > gpurun_out/pytest_stream_raster.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_stream_raster.log
for r in 1 0; do
  TGCN_EVAL_STREAM_RASTER=$r timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-torch-ref > gpurun_out/bench_c2_stream_raster$r.json 2> gpurun_out/bench_c2_stream_raster$r.err; echo "bench c2 raster=$r rc=$?"
done
python - <<'PY'
import json
for r in (1, 0):
    d = json.loads(open(f"gpurun_out/bench_c2_stream_raster{r}.json").read().strip().splitlines()[-1])
    c = d.get("configs") or {}
    print("raster", r, json.dumps(c.get("ltr_pop"))[:400])
PY
