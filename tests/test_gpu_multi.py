"""GPU, needs >= 2 devices (skipped on a one-GPU box): the grid and feature-sliced propagators with NCCL and real
peer-to-peer stores between two B200s, against the single-GPU result of rank 0, plus the item-sharded eval merge."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import lightgcn_oracle as O
        from textgcn_b200 import dist as tdist
        from textgcn_b200 import ops
        from textgcn_b200.graph import norm_adj_csr
        nu, ni, ne, d, L, k = 6001, 1500, 60000, 64, 3, 20
        tu, ti = O.synthetic_interactions(nu, ni, ne, seed=5)
        rowptr, col, val = norm_adj_csr(torch.as_tensor(tu).to(dev), torch.as_tensor(ti).to(dev), nu, ni)
        rowptr, col, val = rowptr.contiguous(), col.contiguous(), val.contiguous()
        gen = torch.Generator(device=dev).manual_seed(0)
        uw, iw = torch.randn(nu, d, generator=gen, device=dev) * 0.1, torch.randn(ni, d, generator=gen, device=dev) * 0.1
        whole = ops.Graph(nu, ni, rowptr, col, val)
        ref = ops.propagate_fwd(whole, uw, iw, L)
        errs = {}
        for G_, R_ in ((2, 1), (1, 2)):
            gp = tdist.GridPartition(rowptr, nu, ni, d, G_, R_)
            gg, rr = gp.coords(rank)
            row_group = None
            for g_id in range(G_):
                grp = dist.new_group(gp.row_group_ranks(g_id)) if R_ > 1 else None
                if g_id == gg:
                    row_group = grp
            u0, u1 = gp.rows.users(rr)
            ug = ops.Graph(nu, ni, *gp.rows.user_block(rr, rowptr, col, val), row_begin=u0, block=True)
            ig = ops.Graph(nu, ni, *gp.rows.item_block(rr, rowptr, col, val), row_begin=nu, block=True)
            for exchange in ("p2p", "collective"):
                prop = tdist.GridPropagator(gp, rank, ug, ig, L, dev, row_group=row_group, exchange=exchange)
                c0, c1 = gp.cols(gg)
                for _ in range(2):
                    g_u, g_i = prop.propagate(uw[u0:u1, c0:c1].contiguous(), iw[:, c0:c1].contiguous())
                torch.cuda.synchronize()
                h0, h1 = gp.final_users(rank)
                errs[(G_, R_, exchange)] = max(float((g_u - ref[h0:h1]).abs().max() / ref.abs().max()),
                                               float((g_i - ref[nu:]).abs().max() / ref.abs().max()))
                dist.barrier()
                prop.close()
        # feature-sliced, whole-graph handle, one call per rank
        fp = tdist.FeatureSlicePartition(nu, ni, d, world)
        sp = tdist.SlicedPropagator(fp, rank, whole, L, dev, exchange="p2p")
        s_u, s_i = sp.propagate(*fp.slice_tables(rank, uw, iw))
        torch.cuda.synchronize()
        f0, f1 = fp.users(rank)
        sliced_identical = bool(torch.equal(s_u, ref[f0:f1]) and torch.equal(s_i, ref[nu:]))
        # item-sharded eval + cross-GPU merge == single-GPU ranking
        users = torch.arange(1024, dtype=torch.int32, device=dev)
        a_ids, a_sc = ops.eval_topk(whole, ref[:nu], ref[nu:], k, users=users)
        b_ids, b_sc = tdist.sharded_eval_topk(whole, ref[:nu], ref[nu:], users, k, rank, world, gather=True)
        merged_identical = bool(torch.equal(a_ids, b_ids) and torch.equal(a_sc, b_sc))
        dist.barrier()
        sp.close()
        # the same two exchanges through the C-ABI communicator (tgcn_comm_*, tgcn_allreduce_sum_f32, tgcn_topk_exchange):
        # grid 1 x 2 with the per-hop all-reduce on the library's own ncclComm_t, single-layer output through the p2p
        # exchange, item-sharded eval with the partial tables exchanged by tgcn_topk_exchange
        comm = tdist.CabiComm(list(range(world)), rank, dev)
        gp = tdist.GridPartition(rowptr, nu, ni, d, 1, 2)
        u0, u1 = gp.rows.users(rank)
        ug = ops.Graph(nu, ni, *gp.rows.user_block(rank, rowptr, col, val), row_begin=u0, block=True)
        ig = ops.Graph(nu, ni, *gp.rows.item_block(rank, rowptr, col, val), row_begin=nu, block=True)
        prop = tdist.GridPropagator(gp, rank, ug, ig, L, dev, exchange="p2p", comm=comm)
        h0, h1 = gp.final_users(rank)
        ref_single = ops.propagate_fwd(whole, uw, iw, L, single=True)
        for single, want in ((False, ref), (True, ref_single), (False, ref)):
            g_u, g_i = prop.propagate(uw[u0:u1].contiguous(), iw, single=single)
            torch.cuda.synchronize()
            errs[("cabi", 2, "single" if single else "mean")] = max(float((g_u - want[h0:h1]).abs().max() / want.abs().max()),
                                                                   float((g_i - want[nu:]).abs().max() / want.abs().max()))
        c_ids, c_sc = tdist.sharded_eval_topk(whole, ref[:nu], ref[nu:], users, k, rank, world, gather=True, comm=comm)
        cabi_identical = bool(torch.equal(a_ids, c_ids) and torch.equal(a_sc, c_sc))
        send = torch.full((5,), float(rank + 1), device=dev)
        recv = torch.empty(5 * world, device=dev)
        ops.comm_allgather(comm.handle, send, recv)
        torch.cuda.synchronize()
        gathered_ok = recv.view(world, 5).eq(torch.arange(1, world + 1, device=dev, dtype=torch.float32)[:, None]).all().item()
        dist.barrier()
        prop.close()
        comm.close()
        ok = sliced_identical and merged_identical and cabi_identical and gathered_ok and all(
            e <= (0.0 if key[1] == 1 else 1e-6) for key, e in errs.items())
        ret[rank] = "ok" if ok else f"mismatch: {errs} sliced={sliced_identical} merged={merged_identical} cabi={cabi_identical}"
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_grid_and_sliced_propagation_over_nccl_and_peer_memory_two_gpus():
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
