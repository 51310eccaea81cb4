#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "topk or predict or eval or tf32 or ltr or base_model or screen" > gpurun_out/pytest_screen.log 2>&1; echo "pytest eval subset rc=$?"; tail -3 gpurun_out/pytest_screen.log
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen_dbg.log 2>&1; echo "pytest screen (debug build) rc=$?"; tail -3 gpurun_out/pytest_screen_dbg.log
timeout 900 python tools/screen_crossover.py > gpurun_out/screen_crossover.jsonl 2> gpurun_out/screen_crossover.err; echo rc=$?; cat gpurun_out/screen_crossover.jsonl; tail -3 gpurun_out/screen_crossover.err
