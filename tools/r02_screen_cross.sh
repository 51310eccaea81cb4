#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_screen.py -m gpu -x -q > gpurun_out/pytest_screen.log 2>&1; echo "pytest screen rc=$?"; tail -3 gpurun_out/pytest_screen.log
timeout 900 python tools/screen_crossover.py > gpurun_out/screen_crossover.jsonl 2> gpurun_out/screen_crossover.err; echo rc=$?; cat gpurun_out/screen_crossover.jsonl; tail -3 gpurun_out/screen_crossover.err
