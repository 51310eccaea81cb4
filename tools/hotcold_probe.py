"""VERDICT r01 item 8 — the hot/cold two-phase experiment for the c5 user-row SpMM pass (one gpurun, decide and record).

User rows gather from the 1 GB item table; sigma_i = 1.3 puts ~half of all gathers on the top 10 % of items (~100 MB, L2-sized).
Phase 1 = every user row restricted to its HOT items (gather set L2-resident), phase 2 = the COLD remainder with accumulate.
This probe needs no kernel change: it builds the two restricted CSRs and times them against the single full pass.

    python tools/hotcold_probe.py [hot_mb ...]          # prints one JSON line per hot-set size
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, timed_steps  # noqa: E402
from textgcn_b200 import ops  # noqa: E402


def restrict(rp, col, val, keep):
    counts = (rp[1:] - rp[:-1]).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(rp.numel() - 1, device=rp.device), counts)[keep]
    nrp = torch.zeros(rp.numel(), dtype=torch.int64, device=rp.device)
    nrp[1:] = torch.cumsum(torch.bincount(rows, minlength=rp.numel() - 1), 0)
    return nrp.to(torch.int32).contiguous(), col[keep].contiguous(), val[keep].contiguous()


def main():
    dev = torch.device("cuda:0")
    w = build_workload(os.environ.get("PROBE_WORKLOAD", "c5"), dev)
    nu, ni, d = w["nu"], w["ni"], w["d"]
    n_u = int(w["rowptr"][nu])
    rp = w["rowptr"][:nu + 1].contiguous()
    col = (w["col"][:n_u] - nu).contiguous()
    val = w["val"][:n_u].contiguous()
    deg_i = (w["rowptr"][nu + 1:] - w["rowptr"][nu:-1]).to(torch.int64)
    order = torch.argsort(deg_i, descending=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    x = w["iw"]
    y = torch.empty((nu, d), dtype=torch.float32, device=dev)
    full = ops.Graph(nu, ni, rp, col, val, row_begin=0, block=True)
    t_full = timed_steps(lambda: ops.spmm_ex(full, x, y), 5, 2, flush, torch)
    ref = y.clone()
    base = sum(t_full) / len(t_full)
    print(json.dumps({"pass": "full user rows", "ms": base, "nnz": n_u}), flush=True)
    for mb in [float(a) for a in sys.argv[1:]] or [48.0, 96.0]:
        n_hot = int(mb * (1 << 20) / (d * 4))
        hot = torch.zeros(ni, dtype=torch.bool, device=dev)
        hot[order[:n_hot]] = True
        is_hot = hot[col.long()]
        gh = ops.Graph(nu, ni, *restrict(rp, col, val, is_hot), row_begin=0, block=True)
        gc = ops.Graph(nu, ni, *restrict(rp, col, val, ~is_hot), row_begin=0, block=True)
        t_hot = timed_steps(lambda: ops.spmm_ex(gh, x, y), 5, 2, flush, torch)
        t_cold = timed_steps(lambda: ops.spmm_ex(gc, x, y, accumulate=True), 5, 2, flush, torch)

        def both():
            ops.spmm_ex(gh, x, y)
            ops.spmm_ex(gc, x, y, accumulate=True)

        t_both = timed_steps(both, 5, 2, flush, torch)
        both()
        err = float((y - ref).abs().max() / ref.abs().max())
        print(json.dumps({"hot_mb": mb, "hot_items": n_hot, "hot_nnz_share": float(is_hot.float().mean()),
                          "hot_ms": sum(t_hot) / 5, "cold_ms": sum(t_cold) / 5, "two_phase_ms": sum(t_both) / 5, "full_ms": base,
                          "gain": 1 - (sum(t_both) / 5) / base, "rel_err_vs_full": err}), flush=True)
        del gh, gc


if __name__ == "__main__":
    main()
