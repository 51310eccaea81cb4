"""The screened eval path (precision="screen" / "auto"): one TF32 product per score finds the candidates, exact fp32
re-scoring orders them, a per-row certificate decides whether the row needs the 3xTF32 second pass (eval_tc.cu).

The property under test is that the RESULT is the exact top-k regardless of how rough the first pass is — on ordinary
embeddings (nothing or next to nothing queued), on near-duplicate items (whole clusters inside the error band, rows
queued for the second pass), with train-item masking, gathered / packed user rows, item sub-ranges and merges.
"""
import numpy as np
import pytest
import torch

from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from textgcn_b200 import ops as _ops
    return _ops


def _reference(ue, ie, k, mask=None):
    """fp64 scores on the device -> canonical (score desc, id asc) top-k."""
    sc = ue.double() @ ie.double().T
    if mask is not None:
        sc[mask] = -float("inf")
    return O.canonical_topk(sc.cpu().numpy(), k)


def _check(got_ids, got_sc, o_ids, o_sc, bound, max_inexact=5):
    assert np.abs(got_sc.cpu().numpy() - o_sc).max() <= bound
    st = O.topk_lists_equivalent(got_ids.cpu().numpy().astype(np.int64), got_sc.cpu().numpy(), o_ids, o_sc.astype(np.float32),
                                 rtol=0, atol=2 * bound)
    assert st["bad"] == 0 and st["exact"] >= st["rows"] - max_inexact, st


@pytest.mark.parametrize("n_rank,n_items,d,k", [(1500, 60000, 64, 20), (700, 300000, 128, 20), (129, 9000, 100, 10), (300, 2000, 32, 1)])
def test_screen_matches_fp64_on_ordinary_embeddings(ops, n_rank, n_items, d, k):
    gen = torch.Generator(device=DEV).manual_seed(n_items + d)
    ue = torch.randn(n_rank, d, generator=gen, device=DEV) * 0.3
    ie = torch.randn(n_items, d, generator=gen, device=DEV) * (0.1 + torch.rand(n_items, 1, generator=gen, device=DEV))  # ragged item norms
    o_ids, o_sc = _reference(ue, ie, k)
    bound = 1e-5 * float(ue.norm(dim=1).max() * ie.norm(dim=1).max())
    for precision in ("screen", "auto"):
        stats = {}
        ids, sc = ops.eval_topk(None, ue, ie, k, precision=precision, stats=stats)
        _check(ids, sc, o_ids, o_sc, bound)
        if precision == "screen" or n_items >= (65536 if d > 96 else 98304 if d > 64 else 131072):  # auto: long sweeps only
            assert stats["precision"] == "screen" and stats["second_pass_rows"] <= n_rank // 20, stats
        else:
            assert stats["precision"] == "3xtf32" and stats["second_pass_rows"] is None, stats


@pytest.mark.parametrize("spread", [0.0, 1e-6, 1e-4, 3e-3])
def test_screen_near_duplicate_items_go_through_the_second_pass(ops, spread):
    """Clusters of 64 items that differ by `spread` (relative): far more than 40 candidates sit inside the TF32 error band
    of the k-th best, so the first pass cannot certify those rows — the answer must still be exact."""
    gen = torch.Generator(device=DEV).manual_seed(7)
    d, k, n_rank = 64, 20, 400
    centres = torch.randn(500, d, generator=gen, device=DEV) * 0.3
    ie = centres.repeat_interleave(64, dim=0)
    ie = ie * (1 + spread * torch.randn(ie.shape, generator=gen, device=DEV))
    ue = torch.randn(n_rank, d, generator=gen, device=DEV) * 0.3
    o_ids, o_sc = _reference(ue, ie, k)
    bound = 1e-5 * float(ue.norm(dim=1).max() * ie.norm(dim=1).max())
    stats = {}
    ids, sc = ops.eval_topk(None, ue, ie, k, precision="screen", stats=stats)
    # ties (spread 0) and near-ties are resolved by fp32 rounding: every row may need the tie tolerance, none may be wrong
    _check(ids, sc, o_ids, o_sc, bound, max_inexact=n_rank)
    if spread <= 1e-4:
        assert stats["second_pass_rows"] > 0, stats
    ids3, sc3 = ops.eval_topk(None, ue, ie, k, precision="3xtf32")
    _check(ids3, sc3, o_ids, o_sc, bound, max_inexact=n_rank)
    if spread == 0.0:  # exact duplicates: canonical order (lowest ids of the plateau) — identical tables from every path
        ids32, _ = ops.eval_topk(None, ue, ie, k, precision="fp32")
        assert torch.equal(ids, ids32)
    # item-sharded + merged (rows take the second pass in some shards and not in others): same exact scores, same table —
    # except that items within the 3xTF32 error of each other (spread 1e-6) may swap at the k-th place
    cuts = [0, 9000, 20000, ie.shape[0]]
    parts = [ops.eval_topk(None, ue, ie, k, item_range=(a, b), finalize=False, precision="screen") for a, b in zip(cuts[:-1], cuts[1:])]
    m_ids, m_sc = ops.topk_merge(None, torch.stack([p[0] for p in parts]).contiguous(), torch.stack([p[1] for p in parts]).contiguous())
    if spread != 1e-6:
        assert torch.equal(m_ids, ids) and torch.equal(m_sc, sc)
    else:
        _check(m_ids, m_sc, o_ids, o_sc, bound, max_inexact=n_rank)


def test_screen_with_mask_gathered_users_ranges_and_merge(ops):
    rng = np.random.default_rng(3)
    nu, n_items, d, k, n_rank = 3000, 20000, 64, 20, 1111
    tu, ti = O.synthetic_interactions(nu, n_items, 90000, seed=5)
    row, col, val = O.norm_adj_coo(tu, ti, nu, n_items)
    gr = ops.Graph.from_norm_matrix(O.sparse_tensor(row, col, val, nu + n_items).to(DEV), nu, n_items)
    gen = torch.Generator(device=DEV).manual_seed(11)
    ue = torch.randn(nu, d, generator=gen, device=DEV) * 0.3
    ie = torch.randn(n_items, d, generator=gen, device=DEV) * 0.3
    users = rng.permutation(nu)[:n_rank]
    u_dev = ops.as_index(users, DEV)
    mask = torch.zeros(n_rank, n_items, dtype=torch.bool, device=DEV)
    pos = {int(u): j for j, u in enumerate(users)}
    sel = np.array([(pos[int(u)], int(i)) for u, i in zip(tu, ti) if int(u) in pos], dtype=np.int64)
    mask[torch.from_numpy(sel[:, 0]).to(DEV), torch.from_numpy(sel[:, 1]).to(DEV)] = True
    o_ids, o_sc = _reference(ue[torch.from_numpy(users).to(DEV)], ie, k, mask)
    bound = 1e-5 * float(ue.norm(dim=1).max() * ie.norm(dim=1).max())
    # rows gathered by user id
    ids, sc = ops.eval_topk(gr, ue, ie, k, users=u_dev, precision="screen")
    _check(ids, sc, o_ids, o_sc, bound)
    # rows already packed in list order
    packed = ue[torch.from_numpy(users).to(DEV)].contiguous()
    ids_p, sc_p = ops.eval_topk(gr, packed, ie, k, users=u_dev, by_position=True, precision="screen")
    assert torch.equal(ids_p, ids) and torch.equal(sc_p, sc)
    # item sub-ranges + merge (the multi-GPU eval path)
    cuts = [0, 4000, 4100, n_items]
    parts = [ops.eval_topk(gr, ue, ie, k, users=u_dev, item_range=(a, b), finalize=False, precision="screen")
             for a, b in zip(cuts[:-1], cuts[1:])]
    m_ids, m_sc = ops.topk_merge(gr, torch.stack([p[0] for p in parts]).contiguous(), torch.stack([p[1] for p in parts]).contiguous(),
                                 users=u_dev)
    assert torch.equal(m_ids, ids) and torch.equal(m_sc, sc)
    # few users: the item range is split over CTAs (partial lists + merge inside the call)
    few = u_dev[:40].contiguous()
    ids_f, sc_f = ops.eval_topk(gr, ue, ie, k, users=few, precision="screen")
    assert torch.equal(ids_f, ids[:40]) and torch.equal(sc_f, sc[:40])


def test_screen_few_rankable_items(ops):
    """Fewer rankable items than k: the list is completed with train items, lowest id first, score -inf (G9)."""
    rng = np.random.default_rng(9)
    nu, n_items, d, k = 200, 30, 64, 20
    tu, ti = O.synthetic_interactions(nu, n_items, 3000, seed=2)
    row, col, val = O.norm_adj_coo(tu, ti, nu, n_items)
    gr = ops.Graph.from_norm_matrix(O.sparse_tensor(row, col, val, nu + n_items).to(DEV), nu, n_items)
    ue = torch.from_numpy((rng.integers(-32, 33, size=(nu, d)) / 16).astype(np.float32)).to(DEV)
    ie = torch.from_numpy((rng.integers(-8, 9, size=(n_items, d)) / 16).astype(np.float32)).to(DEV)
    ids, sc = ops.eval_topk(gr, ue, ie, k, precision="screen")
    ids32, sc32 = ops.eval_topk(gr, ue, ie, k, precision="fp32")
    assert torch.equal(ids, ids32) and torch.equal(sc, sc32)
    assert bool((sc == -float("inf")).any())  # the fixture does contain short lists


def test_screen_is_run_to_run_deterministic(ops):
    """The candidates reach a row's inserter thread from two scanner threads in a timing-dependent order and uncertified rows are
    queued with atomics; neither may show in the result."""
    gen = torch.Generator(device=DEV).manual_seed(21)
    n_rank, n_items, d, k = 900, 150000, 128, 20
    ue = torch.randn(n_rank, d, generator=gen, device=DEV) * 0.3
    ie = (torch.randn(n_items // 50, d, generator=gen, device=DEV) * 0.3).repeat_interleave(50, dim=0)  # plateaus of exact ties
    ie = ie * (1 + 1e-3 * torch.randn(ie.shape, generator=gen, device=DEV) * (torch.arange(n_items, device=DEV) % 3 == 0)[:, None])  # + near-ties inside the band
    first = ops.eval_topk(None, ue, ie, k, precision="screen")
    for _ in range(3):
        again = ops.eval_topk(None, ue, ie, k, precision="screen")
        assert torch.equal(first[0], again[0]) and torch.equal(first[1], again[1])


@pytest.mark.parametrize("n_rank,n_items,d,k,bias", [(300, 20000, 1600, 20, True), (200, 30000, 260, 10, False), (257, 9000, 64, 20, True),
                                                      (140, 70000, 128, 24, True)])
def test_screen_streamed_form_with_bias_terms(ops, n_rank, n_items, d, k, bias):
    """Wide contractions (the LTR score width) and bias terms take the streamed screened kernel: raw operands travel through the
    ring chunk by chunk, the bias terms in one extra chunk whose share of the error bound is accounted separately, and the exact
    re-scoring reads the user rows from global memory."""
    gen = torch.Generator(device=DEV).manual_seed(d + n_items)
    ue = torch.randn(n_rank, d, generator=gen, device=DEV) * (0.3 / (d / 64) ** 0.5)
    ie = torch.randn(n_items, d, generator=gen, device=DEV) * (0.3 / (d / 64) ** 0.5)
    ub = torch.randn(n_rank, generator=gen, device=DEV) * 0.5
    ib = torch.randn(n_items, generator=gen, device=DEV) * 0.05
    sc64 = ue.double() @ ie.double().T
    if bias:
        sc64 = sc64 + ub.double()[:, None] + ib.double()[None, :]
    o_ids, o_sc = O.canonical_topk(sc64.cpu().numpy(), k)
    bound = 1e-5 * float(ue.norm(dim=1).max() * ie.norm(dim=1).max() + (ub.abs().max() + ib.abs().max() if bias else 0))
    kw = dict(user_bias=ub, item_bias=ib) if bias else {}
    stats = {}
    ids, sc = ops.eval_topk(None, ue, ie, k, precision="screen", stats=stats, **kw)
    _check(ids, sc, o_ids, o_sc, bound)
    assert stats["precision"] == "screen" and stats["second_pass_rows"] <= n_rank // 10, stats
    # gathered users + mask-free item sub-ranges + merge give the same table
    perm = torch.randperm(n_rank, generator=gen, device=DEV).to(torch.int32)
    kw_g = dict(user_bias=ub, item_bias=ib) if bias else {}
    ids_g, sc_g = ops.eval_topk(None, ue, ie, k, users=perm, precision="screen", **kw_g)
    assert torch.equal(ids_g, ids[perm.long()]) and torch.equal(sc_g, sc[perm.long()])
    cuts = [0, n_items // 4, n_items]
    parts = [ops.eval_topk(None, ue, ie, k, item_range=(a, b), finalize=False, precision="screen", **kw) for a, b in zip(cuts[:-1], cuts[1:])]
    m_ids, m_sc = ops.topk_merge(None, torch.stack([p[0] for p in parts]).contiguous(), torch.stack([p[1] for p in parts]).contiguous())
    assert torch.equal(m_ids, ids) and torch.equal(m_sc, sc)
