set -x
python tools/eval_c5_probe.py > gpurun_out/eval_c5_probe.json 2> gpurun_out/eval_c5_probe.err && ncu --set full --clock-control none --import-source on -k regex:eval_topk_tc -c 1 -o gpurun_out/prof_eval_c5_v5 python tools/eval_c5_probe.py > gpurun_out/ncuC5.log 2>&1
cat gpurun_out/eval_c5_probe.json
python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --no-torch-ref --no-train --no-extras --no-e2e --eval-steps 3 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('c2 eval', d['eval']['ms'], d['eval']['users_per_s'])"
