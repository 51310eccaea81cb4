import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from bench import build_workload, timed_steps
from textgcn_b200 import ops
dev = torch.device("cuda:0")
w = build_workload("c5", dev)
nu, ni, d, L = w["nu"], w["ni"], w["d"], w["L"]
graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = ops.propagate_fwd(graph, w["uw"], w["iw"], L)
res = {}
CASES = ((16384, "128"), (18944, "128"), (18944, "256"), (37888, "256"))
if os.environ.get("TGCN_PROBE_ONLY_FINAL"):  # one case, for an ncu capture of the shipped configuration
    CASES = ((18944, "256"),)
for n, bn in CASES:
    os.environ["TGCN_EVAL_BN"] = bn
    users = torch.arange(n, dtype=torch.int32, device=dev)
    t = timed_steps(lambda: ops.eval_topk(graph, out[:nu], out[nu:], 20, users=users), 3, 1, flush, torch)
    ms = sum(t) / len(t)
    res[f"c5_eval_{n}_bn{bn}_ms"] = ms
    res[f"c5_eval_{n}_bn{bn}_users_per_s"] = n / ms * 1e3
    res[f"c5_eval_{n}_bn{bn}_tensor_tflops"] = 3 * 2.0 * d * ni * n / (ms * 1e-3) / 1e12
print(json.dumps(res))
