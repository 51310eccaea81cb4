"""Probe of the screened eval path on a bench workload: rows sent to the second pass, time per call against 3xTF32.
usage: python tools/screen_probe.py c2|c5 [n_users] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from textgcn_b200 import ops  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    dev = torch.device("cuda:0")
    w = bench.build_workload(name, dev)
    nu, ni, d, L = w["nu"], w["ni"], w["d"], w["L"]
    n_users = int(sys.argv[2]) if len(sys.argv) > 2 else (nu if name == "c2" else bench.EVAL_USERS_C5)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
    out = ops.propagate_fwd(graph, w["uw"], w["iw"], L)
    users = torch.arange(n_users, dtype=torch.int32, device=dev)
    inorm = out[nu:].norm(dim=1)
    res = {"workload": name, "n_users": n_users, "item_norm_max": float(inorm.max()), "item_norm_median": float(inorm.median()),
           "item_norm_p99": float(inorm.float().quantile(0.99)) if ni <= 16_000_000 else None}
    for prec in ("screen", "3xtf32"):
        stats = {}
        ops.eval_topk(graph, out[:nu], out[nu:], 20, users=users, precision=prec, stats=stats)
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            ops.eval_topk(graph, out[:nu], out[nu:], 20, users=users, precision=prec)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[prec] = {"ms": min(ts), "second_pass_rows": stats.get("second_pass_rows")}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
