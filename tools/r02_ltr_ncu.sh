#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none -k regex:'eval_topk_tc_kernel' -c 2 -o /tmp/ltr_screen -f python tools/ltr_screen_probe.py > gpurun_out/ltr_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/ltr_screen.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_bytes.sum > gpurun_out/ltr_screen_raw.csv 2>&1
cat gpurun_out/ltr_screen_raw.csv | cut -c1-900
