set -x
B="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-torch-ref --no-train --no-extras --no-e2e --eval-steps 1"
$B > gpurun_out/plainB.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/prof_eval_c2_v6 $B > gpurun_out/ncuB1.log 2>&1
tail -2 gpurun_out/ncuB1.log
