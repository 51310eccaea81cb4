#!/bin/bash
# the streamed screened variant (wide contractions / bias terms: the LTR ranking): tests on both builds, then the LTR leg of the c2 bench, screened (auto) against 3xTF32 (TGCN_EVAL_SCREEN=0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_dropin.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model or screen" > gpurun_out/pytest_stream_screen.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_stream_screen.log
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py -m gpu -x -q -k "screen or tf32 or ltr" > gpurun_out/pytest_stream_screen_dbg.log 2>&1; echo "pytest (debug build) rc=$?"; tail -2 gpurun_out/pytest_stream_screen_dbg.log
for sc in auto 0; do
  if [ "$sc" = "0" ]; then export TGCN_EVAL_SCREEN=0; fi
  timeout 600 python bench.py --workload c2 --steps 5 --no-cpu-baseline --no-train --no-e2e --no-torch-ref --no-eval > gpurun_out/bench_c2_stream_screen_$sc.json 2> gpurun_out/bench_c2_stream_screen_$sc.err; echo "bench c2 screen=$sc rc=$?"
done
python - <<'PY'
import json
for r in ("auto", "0"):
    d = json.loads(open(f"gpurun_out/bench_c2_stream_screen_{r}.json").read().strip().splitlines()[-1])
    c = d.get("configs") or {}
    print("TGCN_EVAL_SCREEN", r, json.dumps(c.get("ltr_pop"))[:260])
PY
