#!/bin/bash
# round 2, second GPU call: packed-quad SpMM A/B, pipelined host path, per-kernel ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu2.log
TGCN_SPMM_PACKED=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmm or propagate or bpr or fit" > gpurun_out/pytest_unpacked.log 2>&1; echo "unpacked rc=$?"; tail -2 gpurun_out/pytest_unpacked.log
C2="--workload c2 --steps 30 --no-cpu-baseline --no-torch-ref --no-extras --no-eval --no-e2e"
for v in "packed1:" "packed2:TGCN_SPMM_QUADS_D64=2" "unpacked:TGCN_SPMM_PACKED=0" "unpacked2:TGCN_SPMM_PACKED=0 TGCN_SPMM_QUADS_D64=2"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py $C2 > gpurun_out/ab_c2_$name.json 2> gpurun_out/ab_c2_$name.err
done
C5="--steps 5 --no-cpu-baseline --no-torch-ref --no-c2 --no-eval --no-e2e"
for v in "packed1:" "packed2:TGCN_SPMM_QUADS_D128=2" "unpacked:TGCN_SPMM_PACKED=0"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py $C5 > gpurun_out/ab_c5_$name.json 2> gpurun_out/ab_c5_$name.err
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; echo "bench rc=$?"
timeout 600 python tools/profile_kernels.py > gpurun_out/profile_plain.log 2>&1; echo "profile plain rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:'spmm|bpr_|adam_|dropout_|adv_select|ltr_|tf32_split|eval_topk|topk_merge|dense_nt|pairwise_adv|topk_metrics|layer_mean|sample_' -o /tmp/r02_kernels -f python tools/profile_kernels.py > gpurun_out/ncu_kernels.log 2>&1; echo "ncu rc=$?"
ls -la /tmp/r02_kernels.ncu-rep
# the report itself is too large to travel (64 MiB cap): export the text pages here
ncu -i /tmp/r02_kernels.ncu-rep --page details > gpurun_out/ncu_kernels_details.txt 2>&1
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/ncu_kernels_raw.csv 2>&1
du -sh gpurun_out
echo done
