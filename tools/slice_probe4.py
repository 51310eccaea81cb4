"""Probe (one GPU): row-order on/off (TGCN_ROW_ORDER) for the full-width propagation and for one rank's share of the
feature-sliced propagation at P = 2, 4, 8; c5 and c2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, timed_steps  # noqa: E402
from textgcn_b200 import ops  # noqa: E402
from textgcn_b200.dist import FeatureSlicePartition  # noqa: E402

dev = torch.device("cuda:0")
res = {}
for name in ("c2", "c5"):
    w = build_workload(name, dev)
    nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
    graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = torch.empty((nu + ni, d), dtype=torch.float32, device=dev)
    out_u, out_i = out[:nu], out[nu:]
    steps = 10 if name == "c2" else 3
    for order in ("0", "1"):
        os.environ["TGCN_ROW_ORDER"] = order
        t = timed_steps(lambda: ops.propagate_fwd(graph, w["uw"], w["iw"], L, out=out), steps, 2, flush, torch)
        res[f"{name}_full_order{order}_ms"] = round(sum(t) / len(t), 4)
    ref = out.clone()
    for P in (2, 4, 8) if name == "c5" else (2, 4):
        part = FeatureSlicePartition(nu, ni, d, P)
        us, its = part.slice_tables(P - 1, w["uw"], w["iw"])
        pu = [out_u[q * part.per:].data_ptr() for q in range(P)]
        c0 = part.cols(P - 1)[0]
        for order in ("0", "1"):
            os.environ["TGCN_ROW_ORDER"] = order
            out[:, c0:c0 + part.ds] = 0
            t = timed_steps(lambda: ops.propagate_sliced(graph, us, its, L, d, c0, part.per, pu, [out_i.data_ptr()] * P), steps, 2, flush, torch)
            res[f"{name}_P{P}_order{order}_ms"] = round(sum(t) / len(t), 4)
            res[f"{name}_P{P}_order{order}_bit_identical"] = bool(torch.equal(out[:, c0:c0 + part.ds], ref[:, c0:c0 + part.ds]))
        print(res, file=sys.stderr, flush=True)
    del graph, out, ref, w
    torch.cuda.empty_cache()
print(json.dumps(res))
