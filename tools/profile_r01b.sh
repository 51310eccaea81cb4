set -x
A="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-torch-ref --eval-steps 1"
B="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-torch-ref --no-train --no-extras --no-e2e --eval-steps 1"
$A > gpurun_out/plainA.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv $A > gpurun_out/ncuA.log 2>&1
$B > gpurun_out/plainB.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/prof_eval_c2_r01b $B > gpurun_out/ncuB1.log 2>&1
$B > gpurun_out/plainB2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_group -s 4 -c 1 -o gpurun_out/prof_spmm_c2_r01b $B > gpurun_out/ncuB2.log 2>&1
ls -la gpurun_out | tail -8
