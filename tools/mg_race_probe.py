"""Probe for the intermittent N > 1 parity failure the self-verifying bench caught at c5 / 2 GPUs (profiles/r02/README.md).

Run under torchrun.  Builds one workload, computes the single-GPU reference once per rank, then runs the grid propagator
`--iters` times per variant and counts how often a rank's shard differs from the reference by more than 1e-5 (norm-wise):
  default      : torch NCCL all-reduce on the row group (async), side stream, peer-flag barrier
  sync_hops    : the same with a device synchronize after every kernel / collective call (serialised: a race disappears)
  cabi_comm    : the per-hop all-reduce through the C-ABI communicator (tgcn_allreduce_sum_f32 on its own stream)
  normal_prio  : row-group process group without the high-priority stream option
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import norm_rel_err  # noqa: E402
from textgcn_b200 import dist as tdist  # noqa: E402
from textgcn_b200 import ops  # noqa: E402
from textgcn_b200.graph import norm_adj_csr  # noqa: E402
from textgcn_b200.synthetic import interactions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=4_000_000)
    ap.add_argument("--items", type=int, default=800_000)
    ap.add_argument("--edges", type=int, default=80_000_000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--variants", default="default,sync_hops,cabi_comm,normal_prio")
    ap.add_argument("--burst", type=int, default=1, help="propagate calls issued back to back (no host sync / barrier between them, "
                    "like bench.py's timed loop) before each check")
    ap.add_argument("--ref-repeat", type=int, default=0, help="recompute the single-GPU reference this many times and count bitwise mismatches")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    nu, ni, d, L = args.users, args.items, args.d, args.layers
    tu, ti = interactions(nu, ni, args.edges, dev, seed=0)
    rowptr, col, val = (t.contiguous() for t in norm_adj_csr(tu, ti, nu, ni))
    del tu, ti
    gen = torch.Generator(device=dev).manual_seed(0)
    uw = torch.randn(nu, d, generator=gen, device=dev) * 0.1
    iw = torch.randn(ni, d, generator=gen, device=dev) * 0.1
    whole = ops.Graph(nu, ni, rowptr, col, val)
    ref = ops.propagate_fwd(whole, uw, iw, L)
    ref_mismatch = 0
    for _ in range(args.ref_repeat):
        again = ops.propagate_fwd(whole, uw, iw, L)
        ref_mismatch += int(not torch.equal(again, ref))
        del again
    if rank == 0 and args.ref_repeat:
        print(json.dumps({"single_gpu_reference_recomputed": args.ref_repeat, "bitwise_mismatches": ref_mismatch}), flush=True)
    # do the ranks agree on their inputs and on the single-GPU reference?  (every rank builds the workload itself)
    sums = torch.stack([rowptr.double().sum(), col.double().sum(), val.double().sum(), uw.double().sum(), iw.double().sum(),
                        ref.double().sum(), ref.double().abs().sum()])
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"cross_rank_checksum_spread": ((hi - lo) / hi.abs().clamp_min(1e-30)).tolist(),
                          "fields": ["rowptr", "col", "val", "user_w", "item_w", "ref", "abs(ref)"]}), flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    G, R = (2, world // 2) if world >= 8 else (1, world)
    gp = tdist.GridPartition(rowptr, nu, ni, d, G, R)
    gg, rr = gp.coords(rank)
    u0, u1 = gp.rows.users(rr)
    c0, c1 = gp.cols(gg)
    f0, f1 = gp.final_users(rank)
    ug = ops.Graph(nu, ni, *gp.rows.user_block(rr, rowptr, col, val), row_begin=u0, block=True)
    ig = ops.Graph(nu, ni, *gp.rows.item_block(rr, rowptr, col, val), row_begin=nu, block=True)
    e0_u, e0_i = uw[u0:u1, c0:c1].contiguous(), iw[:, c0:c1].contiguous()
    out = {}
    for variant in args.variants.split(","):
        groups = []
        for g_id in range(G):
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=variant != "normal_prio")
            groups.append(dist.new_group(gp.row_group_ranks(g_id), pg_options=opts) if R > 1 else None)
        comm = tdist.CabiComm(gp.row_group_ranks(gg), rank, dev, group=groups[gg]) if (variant == "cabi_comm" and R > 1) else None
        prop = tdist.GridPropagator(gp, rank, ug, ig, L, dev, row_group=groups[gg], comm=comm)
        if variant == "sync_hops":
            inner_spmm, inner_mean = prop._spmm_impl, prop._mean_impl

            def spmm_sync(*a, _f=inner_spmm):
                r = _f(*a)
                torch.cuda.synchronize()
                return r

            def mean_sync(*a, _f=inner_mean):
                torch.cuda.synchronize()
                r = _f(*a)
                torch.cuda.synchronize()
                return r

            prop._spmm_impl, prop._mean_impl = spmm_sync, mean_sync
        bad, worst, worst_i, worst_u = 0, 0.0, 0.0, 0.0
        for it in range(args.iters):
            for _ in range(args.burst):
                flush.zero_()
                o_u, o_i = prop.propagate(e0_u, e0_i)
            torch.cuda.synchronize()
            dist.barrier()
            eu = norm_rel_err(o_u, ref[f0:f1], torch) if f1 > f0 else 0.0
            ei = norm_rel_err(o_i, ref[nu:], torch)
            e = torch.tensor([max(eu, ei), eu, ei], dtype=torch.float64, device=dev)
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            if float(e[0]) > 1e-5 and it == 0:   # where are the wrong rows?  (every rank reports its own view)
                deg = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
                for tag, got, want, base in (("users", o_u, ref[f0:f1], f0), ("items", o_i, ref[nu:], nu)):
                    if got.numel() == 0:
                        continue
                    scale = float(want.abs().max())
                    row_err = torch.cat([(got[s0:s0 + (1 << 20)] - want[s0:s0 + (1 << 20)]).abs().amax(1) for s0 in range(0, got.shape[0], 1 << 20)])
                    wrong = torch.nonzero(row_err > 1e-5 * scale).flatten()
                    if wrong.numel():
                        dg = deg[base + wrong]
                        print(json.dumps({"variant": variant, "rank": rank, "table": tag, "rows": int(got.shape[0]), "wrong_rows": int(wrong.numel()),
                                          "first": wrong[:8].tolist(), "last": wrong[-4:].tolist(), "deg_min": int(dg.min()), "deg_max": int(dg.max()),
                                          "deg_median": int(dg.median()), "n_wrong_deg_le_128": int((dg <= 128).sum()),
                                          "max_row_err_rel": float(row_err.max() / scale)}), flush=True)
            bad += int(float(e[0]) > 1e-5)
            worst, worst_u, worst_i = max(worst, float(e[0])), max(worst_u, float(e[1])), max(worst_i, float(e[2]))
        out[variant] = {"iters": args.iters, "bad": bad, "worst": worst, "worst_users": worst_u, "worst_items": worst_i}
        dist.barrier()
        prop.close()
        if comm is not None:
            comm.close()
        if rank == 0:
            print(json.dumps({variant: out[variant]}), flush=True)
    if rank == 0:
        print(json.dumps({"workload": [nu, ni, args.edges, d, L], "grid": [G, R], "result": out}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
