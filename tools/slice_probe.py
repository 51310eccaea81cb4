"""Probe (one GPU): time ONE rank's share of the feature-sliced propagation at P = 1, 2, 4, 8 on a bench workload.
All slices store into local tables (n_peers = 1), so this measures the SpMM side only — the per-rank compute time a
P-GPU run would see without NVLink stores — and P·t(P) is what a single GPU would need for P column-blocked passes."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, timed_steps  # noqa: E402
from textgcn_b200 import ops  # noqa: E402
from textgcn_b200.dist import FeatureSlicePartition  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c5")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--only", type=int, default=0, help="time only this slice count")
ap.add_argument("--skip-full", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
w = build_workload(args.workload, dev)
nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty((nu + ni, d), dtype=torch.float32, device=dev)
res = {"workload": args.workload, "nnz": nnz, "d": d, "L": L}
if not args.skip_full:
    t = timed_steps(lambda: ops.propagate_fwd(graph, w["uw"], w["iw"], L, out=out), args.steps, 1, flush, torch)
    res["full_ms"] = sum(t) / len(t)
out_u, out_i = out[:nu], out[nu:]
for P in (2, 4, 8):
    if d % (4 * P) or (args.only and P != args.only):
        continue
    part = FeatureSlicePartition(nu, ni, d, P)
    us, its = part.slice_tables(P - 1, w["uw"], w["iw"])
    pu = [out_u[q * part.per:].data_ptr() for q in range(P)]

    def step():
        ops.propagate_sliced(graph, us, its, L, d, part.cols(P - 1)[0], part.per, pu, [out_i.data_ptr()] * P)

    t = timed_steps(step, args.steps, 1, flush, torch)
    ms = sum(t) / len(t)
    res[f"slice_P{P}_ms"] = ms
    if "full_ms" in res:
        res[f"slice_P{P}_speedup_vs_full"] = res["full_ms"] / ms
print(json.dumps(res))
