#!/bin/bash
# CTA pairs in the STREAMED eval variant (LTR ranking, K = 1600 + bias chunk): tests, then the LTR leg of the c2 bench with pairs on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_dropin.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model or screen" > gpurun_out/pytest_stream_pair.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_stream_pair.log
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tf32 or ltr" > gpurun_out/pytest_stream_pair_dbg.log 2>&1; echo "pytest (debug build) rc=$?"; tail -2 gpurun_out/pytest_stream_pair_dbg.log
for pair in 1 0; do
  if [ "$pair" = "0" ]; then export TGCN_EVAL_PAIR=0; fi
  timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-torch-ref > gpurun_out/bench_c2_stream_pair$pair.json 2> gpurun_out/bench_c2_stream_pair$pair.err; echo "bench c2 pair=$pair rc=$?"
done
python - <<'PY'
import json
for p in (1, 0):
    d = json.loads(open(f"gpurun_out/bench_c2_stream_pair{p}.json").read().strip().splitlines()[-1])
    c = d.get("configs") or d.get("c2", {}).get("configs")
    print("pair", p, json.dumps(c)[:900])
PY
