"""CPU, world_size 2, gloo: the multi-GPU host logic (row partition, column relabelling, all-gather layout, item-sharded
eval exchange + merge).  The CUDA kernels are replaced by CPU stand-ins built on the oracle so only the plumbing is
under test here; the kernels themselves are covered by tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SENT = 2 ** 31 - 1


def _cpu_spmm(graph, x, y, addends, divisor):
    rp, col, val = graph
    n_local = rp.numel() - 1
    row = torch.repeat_interleave(torch.arange(n_local), (rp[1:] - rp[:-1]).long())
    a = torch.sparse_coo_tensor(torch.stack([row, col.long()]), val, (n_local, x.shape[0]))
    acc = torch.sparse.mm(a, x)
    s = acc
    if addends:
        s = addends[0].clone()
        for t in addends[1:]:
            s = s + t
        s = s + acc
    y.copy_(s / divisor if divisor != 1.0 else s)
    return y


def _cpu_rank(train_lists):
    from oracle import lightgcn_oracle as O

    def rank_fn(mask_graph, user_vecs, item_vecs, k, users, item_range):
        i0, i1 = item_range
        scores = (user_vecs[users.long()] @ item_vecs[i0:i1].T).numpy()
        for r, u in enumerate(users.tolist()):
            t = np.asarray(train_lists[u])
            t = t[(t >= i0) & (t < i1)] - i0
            scores[r, t] = -np.inf
        kk = min(k, i1 - i0)
        ids, sc = O.canonical_topk(scores, kk)
        out_i = np.full((len(users), k), SENT, dtype=np.int32)
        out_s = np.full((len(users), k), -np.inf, dtype=np.float32)
        out_i[:, :kk] = ids + i0
        out_s[:, :kk] = sc
        out_i[~np.isfinite(out_s)] = SENT  # unfinalised partial lists carry sentinels, not masked items
        return torch.from_numpy(out_i), torch.from_numpy(out_s)
    return rank_fn


def _cpu_merge(mask_graph, part_ids, part_scores, users):
    p, n, k = part_ids.shape
    ids = part_ids.permute(1, 0, 2).reshape(n, p * k).numpy().astype(np.int64)
    sc = part_scores.permute(1, 0, 2).reshape(n, p * k).numpy()
    out_i = np.empty((n, k), np.int32)
    out_s = np.empty((n, k), np.float32)
    for r in range(n):
        order = np.lexsort((ids[r], -sc[r].astype(np.float64)))[:k]
        out_i[r], out_s[r] = ids[r][order], sc[r][order]
    return torch.from_numpy(out_i), torch.from_numpy(out_s)


def _worker(rank, world, port, case, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import load_golden
        from helpers import golden_lists
        from oracle import lightgcn_oracle as O
        from textgcn_b200 import dist as tdist
        g = load_golden(case)
        nu, ni, L = int(g["n_users"]), int(g["n_items"]), int(g["n_layers"])
        n = nu + ni
        rowptr = torch.from_numpy(O.coo_to_csr(g["norm_row"], n)).to(torch.int32)
        col = torch.from_numpy(g["norm_col"]).to(torch.int32)
        val = torch.from_numpy(g["norm_val"])
        part = tdist.RowPartition(rowptr, world)
        assert part.starts[0] == 0 and part.starts[-1] == n
        nnz_blocks = [int(rowptr[part.starts[p + 1]] - rowptr[part.starts[p]]) for p in range(world)]
        assert max(nnz_blocks) - min(nnz_blocks) <= 2 * int((rowptr[1:] - rowptr[:-1]).max())  # balanced by nnz
        block = part.local_block(rank, rowptr, col, val)
        d = g["user_w"].shape[1]
        prop = tdist.DistPropagator(part, rank, block, d, L, "cpu", spmm_fn=_cpu_spmm)
        e0 = torch.cat([torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"])])
        s, e = part.rows(rank)
        out_local = prop.propagate(e0[s:e].contiguous(), single=bool(g["single"]))
        full = prop.gather_full(out_local)
        ref = np.concatenate([g["rep_user"], g["rep_item"]])
        err = np.abs(full.numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-6, err
        # bipartite scheme: users partitioned, item table all-reduced per hop
        bp = tdist.BipartitePartition(rowptr, nu, ni, world)
        assert bp.starts[0] == 0 and bp.starts[-1] == nu
        ug, ig = bp.user_block(rank, rowptr, col, val), bp.item_block(rank, rowptr, col, val)
        u0, u1 = bp.users(rank)
        assert int(ug[0][-1]) == int(ig[0][-1])                  # every owned edge appears once in each direction
        bprop = tdist.BipartitePropagator(bp, rank, ug, ig, d, L, "cpu", spmm_fn=_cpu_spmm,
                                          mean_fn=lambda adds, out, div: out.copy_(sum(adds[1:], adds[0]) / div))
        out_u = torch.empty((u1 - u0, d))
        out_i = torch.empty((ni, d))
        bprop.propagate(e0[u0:u1].contiguous(), e0[nu:].contiguous(), out_u, out_i, single=bool(g["single"]))
        assert np.abs(out_u.numpy() - g["rep_user"][u0:u1]).max() / np.abs(g["rep_user"]).max() < 1e-6
        assert np.abs(out_i.numpy() - g["rep_item"]).max() / np.abs(g["rep_item"]).max() < 1e-6
        # same with the item rows (and their all-reduce) split into 3 chunks
        irp, icol, ival = ig
        chunks = []
        for c in range(3):
            r0, r1 = ni * c // 3, ni * (c + 1) // 3
            lo, hi = int(irp[r0]), int(irp[r1])
            chunks.append((r0, r1, ((irp[r0:r1 + 1] - irp[r0]).contiguous(), icol[lo:hi].contiguous(), ival[lo:hi].contiguous())))
        cprop = tdist.BipartitePropagator(bp, rank, ug, ig, d, L, "cpu", spmm_fn=_cpu_spmm, item_chunks=chunks,
                                          mean_fn=lambda adds, out, div: out.copy_(sum(adds[1:], adds[0]) / div))
        out_u2, out_i2 = torch.empty_like(out_u), torch.empty_like(out_i)
        cprop.propagate(e0[u0:u1].contiguous(), e0[nu:].contiguous(), out_u2, out_i2, single=bool(g["single"]))
        assert torch.equal(out_u2, out_u) and torch.equal(out_i2, out_i)
        # feature-sliced scheme: all hops local on d/P columns, one exchange at the end (collective form under gloo)
        if d % (4 * world) == 0:
            fp = tdist.FeatureSlicePartition(nu, ni, d, world)
            assert fp.users(0)[0] == 0 and fp.users(world - 1)[1] == nu and fp.cols(world - 1)[1] == d

            def _cpu_local(graph, us, its, n_layers, single, out):
                rp_, col_, val_ = graph
                x = torch.cat([us, its])
                layers = [x]
                for _ in range(n_layers):
                    layers.append(_cpu_spmm((rp_, col_, val_), layers[-1], torch.empty_like(x), [], 1.0))
                out.copy_(layers[-1] if single else torch.mean(torch.stack(layers), dim=0))
                return out

            sp = tdist.SlicedPropagator(fp, rank, (rowptr, col, val), L, "cpu", exchange="collective", local_fn=_cpu_local)
            us, its = fp.slice_tables(rank, e0[:nu], e0[nu:])
            s_u, s_i = sp.propagate(us, its, single=bool(g["single"]))
            f0, f1 = fp.users(rank)
            assert s_u.shape == (f1 - f0, d) and s_i.shape == (ni, d)
            assert np.abs(s_u.numpy() - g["rep_user"][f0:f1]).max() / np.abs(g["rep_user"]).max() < 1e-6
            assert np.abs(s_i.numpy() - g["rep_item"]).max() / np.abs(g["rep_item"]).max() < 1e-6
            # grid scheme on 2 ranks, both degenerate shapes: 2 slices x 1 row part, 1 slice x 2 row parts
            mean_cpu = lambda adds, out, div: out.copy_(sum(adds[1:], adds[0]) / div)  # noqa: E731
            for G_, R_ in ((2, 1), (1, 2)):
                gp = tdist.GridPartition(rowptr, nu, ni, d, G_, R_)
                gg, rr = gp.coords(rank)
                row_group = None
                for g_id in range(G_):  # every rank creates every group, in the same order
                    grp = dist.new_group(gp.row_group_ranks(g_id))
                    if g_id == gg:
                        row_group = grp
                ugb = gp.rows.user_block(rr, rowptr, col, val)
                igb = gp.rows.item_block(rr, rowptr, col, val)
                gprop = tdist.GridPropagator(gp, rank, ugb, igb, L, "cpu", row_group=row_group, exchange="collective",
                                             spmm_fn=_cpu_spmm, mean_fn=mean_cpu)
                c0, c1 = gp.cols(gg)
                gu0, gu1 = gp.rows.users(rr)
                g_u, g_i = gprop.propagate(e0[gu0:gu1, c0:c1].contiguous(), e0[nu:, c0:c1].contiguous(), single=bool(g["single"]))
                h0, h1 = gp.final_users(rank)
                assert g_u.shape == (h1 - h0, d) and g_i.shape == (ni, d)
                assert np.abs(g_u.numpy() - g["rep_user"][h0:h1]).max() / np.abs(g["rep_user"]).max() < 1e-6, (G_, R_)
                assert np.abs(g_i.numpy() - g["rep_item"]).max() / np.abs(g["rep_item"]).max() < 1e-6, (G_, R_)
        # item-sharded eval with cross-rank merge
        k = int(max(g["ks"]))
        tl = golden_lists(g)
        users = torch.arange(nu - nu % world, dtype=torch.int32)
        ids, sc = tdist.sharded_eval_topk(None, full[:nu], full[nu:], users, k, rank, world, gather=True,
                                          rank_fn=_cpu_rank(tl), merge_fn=_cpu_merge)
        o_ids, o_sc = O.predict_topk(full[:nu], full[nu:], users.numpy(), tl, k, round_decimals=None)
        fin = np.isfinite(o_sc)
        assert np.array_equal(ids.numpy()[fin], o_ids[fin]) and np.array_equal(sc.numpy()[fin], o_sc[fin])
        ret[rank] = "ok"
    except Exception as exc:  # pragma: no cover
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["small_lgcn_d64", "small_lgcn_d32_single"])
def test_row_partitioned_propagation_and_item_sharded_eval_world2(case):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, case, ret), nprocs=world, join=True)
    assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_partition_relabel_roundtrip():
    from textgcn_b200 import dist as tdist
    rowptr = torch.tensor([0, 5, 5, 9, 30, 31, 40], dtype=torch.int32)
    part = tdist.RowPartition(rowptr, 3)
    assert part.starts[0] == 0 and part.starts[-1] == 6 and sorted(part.starts) == part.starts
    col = torch.arange(6, dtype=torch.int32)
    rel = part.relabel(col)
    own = part.owner(col.long())
    for c in range(6):
        p = int(own[c])
        assert part.starts[p] <= c < part.starts[p + 1]
        assert int(rel[c]) == p * part.max_rows + (c - part.starts[p])
    assert tdist.item_shard(10, 4, 3) == (9, 10) and tdist.item_shard(10, 4, 0) == (0, 3)


def test_grid_and_slice_partition_geometry():
    """Pure host logic: every user / column / item row has exactly one owner and the ranks of a row group are contiguous."""
    from textgcn_b200 import dist as tdist
    rowptr = torch.cumsum(torch.tensor([0] + [3, 1, 7, 2, 2, 9, 1, 4, 5, 1] + [4] * 7), 0).to(torch.int32)  # 10 users, 7 items
    nu, ni, d = 10, 7, 32
    for G, R in ((1, 4), (2, 2), (4, 1), (2, 4), (8, 1)):
        gp = tdist.GridPartition(rowptr, nu, ni, d, G, R)
        assert gp.world_size == G * R and gp.ds * G == d
        assert sorted(r for g in range(G) for r in gp.row_group_ranks(g)) == list(range(G * R))
        cover_users, cover_final = [], []
        for rank in range(G * R):
            g, r = gp.coords(rank)
            assert rank == g * R + r and gp.cols(g) == (g * gp.ds, (g + 1) * gp.ds)
            cover_final += list(range(*gp.final_users(rank)))
            if g == 0:
                cover_users += list(range(*gp.rows.users(r)))
        # every user has exactly one final owner, and that owner sits in the row group that computed the user (the rows never
        # leave their row partition: only the G feature-slice partners exchange them)
        assert cover_users == list(range(nu)) and sorted(cover_final) == list(range(nu))
        for rank in range(G * R):
            g, r = gp.coords(rank)
            f0, f1 = gp.final_users(rank)
            u0, u1 = gp.rows.users(r)
            assert u0 <= f0 <= f1 <= u1 and f1 - f0 <= gp.per
            assert gp.slice_partners(r)[g] == rank
        shares = [tdist.item_shard(ni, R, r) for r in range(R)]
        assert [i for a, b in shares for i in range(a, b)] == list(range(ni))
    fp = tdist.FeatureSlicePartition(nu, ni, d, 4)
    assert [u for q in range(4) for u in range(*fp.users(q))] == list(range(nu)) and fp.cols(3) == (24, 32)
    with pytest.raises(ValueError):
        tdist.FeatureSlicePartition(nu, ni, 20, 4)
    with pytest.raises(ValueError):
        tdist.GridPartition(rowptr, nu, ni, 20, 2, 2)
