"""Hypothesis probe: does ranking against the item table sorted by descending row norm (thresholds rise early, fewer list updates)
speed the eval kernels up?  Times eval_topk on a bench workload with the item rows as they are and sorted by norm."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from textgcn_b200 import ops  # noqa: E402


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    dev = torch.device("cuda:0")
    w = bench.build_workload(name, dev)
    nu, ni, L = w["nu"], w["ni"], w["L"]
    n_users = nu if name == "c2" else bench.EVAL_USERS_C5
    graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
    out = ops.propagate_fwd(graph, w["uw"], w["iw"], L)
    ue = out[:nu][:n_users].contiguous()
    ie = out[nu:].contiguous()
    order = torch.argsort(ie.norm(dim=1), descending=True)
    ie_sorted = ie[order].contiguous()
    ie_rand = ie[torch.randperm(ni, device=dev)].contiguous()
    res = {"workload": name, "n_users": n_users}
    for label, tab in (("as_is", ie), ("by_norm_desc", ie_sorted), ("shuffled", ie_rand)):
        for prec in ("screen", "3xtf32"):
            res[f"{label}_{prec}_ms"] = round(timed(lambda: ops.eval_topk(None, ue, tab, 20, precision=prec)), 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
