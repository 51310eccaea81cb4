#!/bin/bash
# round 2: final single-GPU evidence — default bench, reference arm, launch list of the same command, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
SHORT="--steps 2 --warmup 1 --no-cpu-baseline --no-torch-ref --no-extras --no-train"
timeout 600 python bench.py $SHORT > gpurun_out/plain_short.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_n1_short.csv python bench.py $SHORT > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
# dominant kernels: one c5 SpMM layer (4th launch = last layer incl. mean epilogue is launch index 3), the c5 eval kernel
PROP="--steps 1 --warmup 1 --no-cpu-baseline --no-torch-ref --no-c2 --no-e2e --eval-steps 1"
timeout 600 python bench.py $PROP > gpurun_out/plain_prop.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -k regex:'spmm_group_kernel|eval_topk_tc_kernel' -s 4 -c 5 -o /tmp/r02_c5 -f python bench.py $PROP > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
ncu -i /tmp/r02_c5.ncu-rep --page details > gpurun_out/ncu_c5_spmm_eval_details.txt 2>&1
ncu -i /tmp/r02_c5.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread > gpurun_out/ncu_c5_spmm_eval_raw.csv 2>&1
# all kernels at c2 with the final SpMM
timeout 600 python tools/profile_kernels.py > gpurun_out/profile_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none -k regex:'spmm|bpr_|adam_|dropout_|adv_select|ltr_|tf32_split|eval_topk|topk_merge|dense_nt|pairwise_adv|topk_metrics|layer_mean|sample_|permute_mask' -o /tmp/r02_kernels -f python tools/profile_kernels.py > gpurun_out/ncu_kernels.log 2>&1; echo "ncu c2 rc=$?"
ncu -i /tmp/r02_kernels.ncu-rep --page details > gpurun_out/ncu_kernels_details.txt 2>&1
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/ncu_kernels_raw.csv 2>&1
du -sh gpurun_out; echo done
