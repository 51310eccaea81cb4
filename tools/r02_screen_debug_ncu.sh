#!/bin/bash
# the eval-related GPU tests on the debug-assert build, then one full ncu capture of the final screened kernel at c5
mkdir -p gpurun_out
TGCN_B200_LIB=$PWD/textgcn_b200/libtgcn_b200_dbg.so timeout 1500 python -m pytest tests -m gpu -q -k "topk or predict or eval or tf32 or ltr or base_model or screen or dropin" > gpurun_out/pytest_gpu_debug.log 2>&1; echo "pytest debug rc=$?"; tail -3 gpurun_out/pytest_gpu_debug.log
bash tools/r02_screen_full.sh c5 37888
ncu -i /tmp/screen_c5.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed.sum > gpurun_out/screen_c5_raw.csv 2>&1
