"""GPU: the feature-sliced multi-GPU propagation (tgcn_propagate_sliced + CUDA IPC peer tables).

* one process, n_peers = 1: P column slices assembled on one GPU must be BIT-identical to the unsliced result (every
  column sees the same non-zeros in the same order whatever the lane layout);
* two processes sharing cuda:0 (gloo for the handle exchange and the barriers — NCCL refuses two ranks on one device):
  the epilogue's stores land in the OTHER process's tables through IPC-mapped pointers.
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import TOL, golden_norm, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.mark.parametrize("case,n_slices", [("small_lgcn_d64", 2), ("small_lgcn_d64", 4), ("small_lgcn_d128_l4", 8),
                                           ("small_lgcn_d32_single", 2), ("small_lgcn_d128_l4", 1)])
def test_slices_assembled_on_one_gpu_are_bit_identical(case, n_slices):
    from textgcn_b200 import ops
    from textgcn_b200.dist import FeatureSlicePartition
    g = load_golden(case)
    nu, ni, L, single = int(g["n_users"]), int(g["n_items"]), int(g["n_layers"]), bool(g["single"])
    gr = ops.Graph.from_norm_matrix(golden_norm(g).to(DEV), nu, ni)
    uw, iw = _cuda(g["user_w"]), _cuda(g["item_w"])
    d = uw.shape[1]
    ref = ops.propagate_fwd(gr, uw, iw, L, single)
    part = FeatureSlicePartition(nu, ni, d, n_slices)
    # emulate the P ranks one after the other: rank q's tables are separate local buffers
    out_u = [torch.full((part.per, d), float("nan"), device=DEV) for _ in range(n_slices)]
    out_i = [torch.full((ni, d), float("nan"), device=DEV) for _ in range(n_slices)]
    for p in range(n_slices):
        us, its = part.slice_tables(p, uw, iw)
        ops.propagate_sliced(gr, us, its, L, d, part.cols(p)[0], part.per, [t.data_ptr() for t in out_u],
                             [t.data_ptr() for t in out_i], single=single)
    torch.cuda.synchronize()
    for q in range(n_slices):
        u0, u1 = part.users(q)
        assert torch.equal(out_u[q][:u1 - u0], ref[u0:u1]), (case, q)
        assert torch.equal(out_i[q], ref[nu:]), (case, q)
    assert rel_err(ref[:nu].cpu().numpy(), g["rep_user"]) < TOL


def test_sliced_with_dropout_mask_and_odd_slice_width():
    from textgcn_b200 import ops
    from textgcn_b200.dist import FeatureSlicePartition
    g = load_golden("small_lgcn_d64")
    nu, ni, L = int(g["n_users"]), int(g["n_items"]), int(g["n_layers"])
    gr = ops.Graph.from_norm_matrix(golden_norm(g).to(DEV), nu, ni)
    gen = torch.Generator(device=DEV).manual_seed(5)
    uw, iw = torch.randn(nu, 48, generator=gen, device=DEV), torch.randn(ni, 48, generator=gen, device=DEV)
    keep = torch.rand(gr.nnz, generator=gen, device=DEV) < 0.6
    ref = ops.propagate_fwd(gr, uw, iw, L, keep=keep, dropout=0.4)
    part = FeatureSlicePartition(nu, ni, 48, 4)  # 12-wide slices: the generic-width kernel
    out_u, out_i = torch.empty((part.per * 4, 48), device=DEV), torch.empty((ni, 48), device=DEV)
    for p in range(4):
        us, its = part.slice_tables(p, uw, iw)
        ops.propagate_sliced(gr, us, its, L, 48, part.cols(p)[0], part.per,
                             [out_u[q * part.per:].data_ptr() for q in range(4)], [out_i.data_ptr()] * 4, keep=keep, dropout=0.4)
    assert rel_err(out_u[:nu].cpu().numpy(), ref[:nu].cpu().numpy()) < 1e-6
    assert rel_err(out_i.cpu().numpy(), ref[nu:].cpu().numpy()) < 1e-6


def _ipc_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from textgcn_b200 import ops
        from textgcn_b200 import dist as tdist
        torch.cuda.set_device(0)
        g = load_golden("small_lgcn_d128_l4")
        nu, ni, L = int(g["n_users"]), int(g["n_items"]), int(g["n_layers"])
        gr = ops.Graph.from_norm_matrix(golden_norm(g).to(DEV), nu, ni)
        uw, iw = _cuda(g["user_w"]), _cuda(g["item_w"])
        d = uw.shape[1]
        ref = ops.propagate_fwd(gr, uw, iw, L)
        part = tdist.FeatureSlicePartition(nu, ni, d, world)
        sp = tdist.SlicedPropagator(part, rank, gr, L, DEV, exchange="p2p")
        us, its = part.slice_tables(rank, uw, iw)
        for _ in range(3):  # reuse of the result tables across steps goes through the leading barrier
            out_u, out_i = sp.propagate(us, its)
        torch.cuda.synchronize()
        u0, u1 = part.users(rank)
        ok = torch.equal(out_u, ref[u0:u1]) and torch.equal(out_i, ref[nu:])
        dist.barrier()
        sp.close()
        # grid scheme, both 2-rank shapes, through the peer-memory exchange (spmm_scatter + layer_mean_scatter)
        rowptr, col, val = gr.rowptr, gr.col, gr.val
        for G_, R_ in ((2, 1), (1, 2)):
            gp = tdist.GridPartition(rowptr, nu, ni, d, G_, R_)
            gg, rr = gp.coords(rank)
            row_group = None
            for g_id in range(G_):
                grp = dist.new_group(gp.row_group_ranks(g_id))
                if g_id == gg:
                    row_group = grp
            u0_, u1_ = gp.rows.users(rr)
            ug = ops.Graph(nu, ni, *gp.rows.user_block(rr, rowptr, col, val), row_begin=u0_, block=True)
            ig = ops.Graph(nu, ni, *gp.rows.item_block(rr, rowptr, col, val), row_begin=nu, block=True)
            gprop = tdist.GridPropagator(gp, rank, ug, ig, L, DEV, row_group=row_group, exchange="p2p")
            c0, c1 = gp.cols(gg)
            for _ in range(2):
                g_u, g_i = gprop.propagate(uw[u0_:u1_, c0:c1].contiguous(), iw[:, c0:c1].contiguous())
            torch.cuda.synchronize()
            h0, h1 = gp.final_users(rank)
            tol = 0.0 if R_ == 1 else 1e-6  # R = 1 sums every row in the single-GPU order; R = 2 adds two partial tables
            eu = float((g_u - ref[h0:h1]).abs().max() / ref.abs().max())
            ei = float((g_i - ref[nu:]).abs().max() / ref.abs().max())
            ok = ok and eu <= tol and ei <= tol
            dist.barrier()
            gprop.close()
        ret[rank] = "ok" if ok else "mismatch"
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_peer_stores_through_cuda_ipc_two_processes():
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    mp.spawn(_ipc_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_sliced_long_rows_take_the_segment_path():
    """Hub items (rows of > 128 non-zeros, cut into segments whose partial sums the last arriver adds) must scatter too."""
    from textgcn_b200 import ops
    from textgcn_b200.dist import FeatureSlicePartition
    from textgcn_b200.graph import graph_from_interactions
    rng = np.random.default_rng(11)
    nu, ni, d, L, P = 3001, 40, 64, 3, 4
    u = np.concatenate([np.arange(nu), rng.integers(nu, size=6000)])
    i = np.concatenate([np.arange(nu) % ni, rng.integers(ni, size=6000)])
    key = np.unique(u.astype(np.int64) * ni + i)
    gr = graph_from_interactions(key // ni, key % ni, nu, ni, DEV)
    assert gr.n_segments > 0
    gen = torch.Generator(device=DEV).manual_seed(1)
    uw, iw = torch.randn(nu, d, generator=gen, device=DEV), torch.randn(ni, d, generator=gen, device=DEV)
    ref = ops.propagate_fwd(gr, uw, iw, L)
    part = FeatureSlicePartition(nu, ni, d, P)
    out_u = [torch.zeros((part.per, d), device=DEV) for _ in range(P)]
    out_i = [torch.zeros((ni, d), device=DEV) for _ in range(P)]
    for p in range(P):
        us, its = part.slice_tables(p, uw, iw)
        ops.propagate_sliced(gr, us, its, L, d, part.cols(p)[0], part.per, [t.data_ptr() for t in out_u],
                             [t.data_ptr() for t in out_i])
    for q in range(P):
        u0, u1 = part.users(q)
        assert torch.equal(out_u[q][:u1 - u0], ref[u0:u1])
        assert torch.equal(out_i[q], ref[nu:])
