"""Probe (one GPU): does column-tiling make the narrow-slice SpMM L2-resident?  One layer of the 16-wide slice at the
200M-edge config, computed as passes over gather-table tiles of T MB (per-tile CSR built with torch ops, physically
contiguous), each pass accumulating into the output rows.  Compared with the untiled layer."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, timed_steps  # noqa: E402
from textgcn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
w = build_workload("c5", dev)
nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
n = nu + ni
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
rowptr, col, val = w["rowptr"], w["col"], w["val"]
graph = ops.Graph(nu, ni, rowptr, col, val)
split = int(rowptr[nu])


def tile_blocks(r0, r1, lo, hi, c_base, n_cols, rows_per_tile):
    """Block handles over rows [r0, r1) (nnz range [lo, hi)), one per tile of `rows_per_tile` gather-table rows."""
    c = col[lo:hi] - c_base
    counts = (rowptr[r0 + 1:r1 + 1] - rowptr[r0:r1]).long()
    rows = torch.repeat_interleave(torch.arange(r1 - r0, device=dev), counts)
    out = []
    for t0 in range(0, n_cols, rows_per_tile):
        keep = (c >= t0) & (c < t0 + rows_per_tile)
        rp = torch.zeros(r1 - r0 + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(torch.bincount(rows[keep], minlength=r1 - r0), 0)
        g = ops.Graph(nu, ni, rp.to(torch.int32).contiguous(), c[keep].contiguous(), val[lo:hi][keep].contiguous(),
                      row_begin=r0, block=True)
        out.append(g)
    return out


for ds in (16,):
    gen = torch.Generator(device=dev).manual_seed(0)
    xu = torch.randn(nu, ds, generator=gen, device=dev)
    xi = torch.randn(ni, ds, generator=gen, device=dev)
    x = torch.cat([xu, xi])
    y_ref = torch.empty((n, ds), device=dev)
    t = timed_steps(lambda: ops.spmm(graph, x, out=y_ref), 3, 1, flush, torch)
    res[f"ds{ds}_untiled_ms"] = round(sum(t) / len(t), 3)
    print(res, file=sys.stderr, flush=True)
    for ph, (r0, r1, lo, hi, cb, nc, xs, ys) in {"user": (0, nu, 0, split, nu, ni, xi, y_ref[:nu]), "item": (nu, n, split, nnz, 0, nu, xu, y_ref[nu:])}.items():
        g1 = tile_blocks(r0, r1, lo, hi, cb, nc, 1 << 30)[0]
        for order in ("0", "1"):
            os.environ["TGCN_ROW_ORDER"] = order
            t = timed_steps(lambda: ops.spmm_ex(g1, xs, ys), 3, 1, flush, torch)
            res[f"ds{ds}_untiled_{ph}phase_order{order}_ms"] = round(sum(t) / len(t), 3)
        del g1
    print(res, file=sys.stderr, flush=True)
    for mb in (48, 64, 96):
        rows_per_tile = (mb << 20) // (ds * 4)
        ub = tile_blocks(0, nu, 0, split, nu, ni, rows_per_tile)        # user rows gather from item tiles
        ib = tile_blocks(nu, n, split, nnz, 0, nu, rows_per_tile)       # item rows gather from user tiles
        y = torch.empty((n, ds), device=dev)

        def layer():
            for k, g in enumerate(ub):
                ops.spmm_ex(g, xi, y[:nu], accumulate=k > 0)
            for k, g in enumerate(ib):
                ops.spmm_ex(g, xu, y[nu:], accumulate=k > 0)

        for order in ("0", "1"):
            os.environ["TGCN_ROW_ORDER"] = order
            t = timed_steps(layer, 3, 1, flush, torch)
            res[f"ds{ds}_tile{mb}MB_{len(ub)}+{len(ib)}passes_order{order}_ms"] = round(sum(t) / len(t), 3)
            # per-phase split
            tu = timed_steps(lambda: [ops.spmm_ex(g, xi, y[:nu], accumulate=k > 0) for k, g in enumerate(ub)], 2, 1, flush, torch)
            res[f"ds{ds}_tile{mb}MB_order{order}_userphase_ms"] = round(sum(tu) / len(tu), 3)
        err = float((y - y_ref).abs().max() / y_ref.abs().max())
        res[f"ds{ds}_tile{mb}MB_relerr"] = err
        print(res, file=sys.stderr, flush=True)
        del ub, ib
print(json.dumps(res))
