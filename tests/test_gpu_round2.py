"""GPU parity, second batch: the kernel-backed bodies of the §8(b) matrix-returning methods (csrc/dense.cu), the device
metrics kernel (csrc/metrics.cu), the sampler-sentinel guard of the fused BPR kernel, the --load_base order on the
standalone classes, the eval kernel at the headline (C5) shape and an SpMM whose tables exceed 2^32 bytes."""
import logging
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import TOL, StubDataset, golden_lists, load_weights, params_from_golden, rel_err
from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from textgcn_b200 import ops as _ops
    return _ops


def _randn(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(DEV)


# ------------------------------------------------------------------------------------------------ a11 / a18 / a20
@pytest.mark.parametrize("m,n,k", [(1, 5, 4), (130, 257, 64), (300, 1000, 1600), (64, 63001, 64), (2048, 999, 128)])
def test_score_batchwise_kernel_matches_fp64(ops, m, n, k):
    a, b = _randn(m, k, seed=m), _randn(n, k, seed=n)
    ref = a.double() @ b.double().T
    out = ops.score_batchwise(a, b)
    assert out.shape == (m, n)
    scale = float(a.norm(dim=1).max() * b.norm(dim=1).max())
    assert float((out.double() - ref).abs().max()) <= 1e-6 * scale           # fp32 FMA accumulation, norm-wise (H4)
    rb, cb = _randn(m, seed=7), _randn(n, seed=8)
    out_b = ops.score_batchwise(a, b, row_bias=rb, col_bias=cb)
    assert float((out_b.double() - (ref + rb.double()[:, None] + cb.double()[None, :])).abs().max()) <= 1e-6 * (scale + 8)
    # feature planes of (M, N, F) written in place, neighbouring planes untouched; operands as column slices of wider tables
    wide_a, wide_b = _randn(m, k + 8, seed=1), _randn(n, k + 12, seed=2)
    planes = torch.full((m, n, 3), 7.0, device=DEV)
    ops.score_batchwise(wide_a[:, 4:4 + k], wide_b[:, 8:8 + k], out=planes, plane=1)
    ref_p = wide_a[:, 4:4 + k].double() @ wide_b[:, 8:8 + k].double().T
    assert float((planes[:, :, 1].double() - ref_p).abs().max()) <= 1e-6 * scale * 2
    assert bool((planes[:, :, 0] == 7).all()) and bool((planes[:, :, 2] == 7).all())


def test_score_batchwise_is_differentiable_like_matmul(ops):
    from textgcn_b200.models import _ScoreBatchwiseFn
    a, b = _randn(37, 64, seed=1).requires_grad_(True), _randn(91, 64, seed=2).requires_grad_(True)
    w = _randn(37, 91, seed=3)
    (_ScoreBatchwiseFn.apply(a, b) * w).sum().backward()
    assert rel_err(a.grad.cpu(), (w @ b.detach()).cpu()) < TOL and rel_err(b.grad.cpu(), (w.T @ a.detach()).cpu()) < TOL


@pytest.mark.parametrize("b,c,d", [(1, 7, 16), (48, 1000, 64), (33, 90, 128), (5, 3, 320)])
def test_score_pairwise_adv_kernel(ops, b, c, d):
    u, it = _randn(b, d, seed=b), _randn(b, c, d, seed=c)
    out = ops.score_pairwise_adv(u, it)
    assert out.shape == (b, c)                                                   # (1, C) stays 2-D (G14)
    ref = torch.einsum("bd,bcd->bc", u.double(), it.double())
    assert float((out.double() - ref).abs().max()) <= 1e-6 * float(u.norm(dim=1).max() * it.norm(dim=2).max())


@pytest.mark.parametrize("b,d,D", [(1, 16, 8), (64, 64, 24), (257, 128, 768)])
def test_ltr_features_rows_kernel(ops, b, d, D):
    packed_u, packed_i = _randn(b, d + 2 * D, seed=1), _randn(b, d + 2 * D + 4, seed=2)   # views of packed rows: strided operands
    ue, ur, ud = packed_u[:, :d], packed_u[:, d:d + D], packed_u[:, d + D:]
    ie, ir, idesc = packed_i[:, :d], packed_i[:, d:d + D], packed_i[:, d + D:d + 2 * D]
    out = ops.ltr_features_rows(ue, ie, ur, ud, ir, idesc)
    ref = O.ltr_features_pairwise(ue.double().cpu(), ur.double().cpu(), ud.double().cpu(), ie.double().cpu(), ir.double().cpu(),
                                  idesc.double().cpu())
    assert out.shape == (b, 5)
    assert float((out.double().cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


# ------------------------------------------------------------------------------------------------ a13 / n4
@pytest.mark.parametrize("case", ["dummy_lgcn", "small_lgcn_d64", "small_lgcn_d128_l4", "small_ltr_pop"])
def test_metrics_kernel_matches_reference_golden(ops, case):
    from textgcn_b200 import metrics as M
    g = load_golden(case)
    test = golden_lists(g, "test")
    res = M.calculate_metrics(torch.from_numpy(g["pred_ids"]).to(DEV), [test[u].tolist() for u in g["test_users"]], g["ks"].tolist())
    for m in M.METRICS:
        assert np.allclose(res[m], g["metric_" + m], rtol=0, atol=1e-12), (case, m)


def test_metrics_kernel_duplicates_ragged_truth_and_scale(ops):
    from textgcn_b200 import metrics as M
    # duplicates in y_true count in the recall denominator, repeated predictions count once (np.intersect1d, utils.py:46)
    res = M.calculate_metrics(torch.tensor([[1, 1, 2]], device=DEV), [[1, 1, 5]], [3])
    ref = O.calculate_metrics([[1, 1, 2]], [[1, 1, 5]], [3])
    for m in M.METRICS:
        assert np.allclose(res[m], ref[m], rtol=0, atol=1e-12), m
    # ragged truth, -1 filler ids (G9 tail of a tiny catalogue), several k incl. kmax = 128, against the oracle
    rng = np.random.default_rng(0)
    n, kmax, n_items = 20000, 128, 500
    pred = np.stack([rng.permutation(n_items)[:kmax] for _ in range(n)]).astype(np.int32)
    pred[::7, 100:] = -1
    lens = rng.integers(1, 9, size=n)
    truth = [rng.integers(n_items, size=l).tolist() for l in lens]
    ks = [1, 5, 20, 40, 128]
    res = M.calculate_metrics(torch.from_numpy(pred).to(DEV), truth, ks)
    ref = O.calculate_metrics(pred.tolist(), truth, ks)
    for m in M.METRICS:
        assert np.allclose(res[m], ref[m], rtol=0, atol=1e-11), m
    # TruthCSR.from_pairs on the device == from_lists; and the pass is deterministic (fixed-order reduction)
    rows = torch.from_numpy(np.repeat(np.arange(n), lens)).to(DEV)
    items = torch.from_numpy(np.concatenate(truth)).to(DEV)
    csr = M.TruthCSR.from_pairs(rows, items, n)
    a = ops.topk_metrics(torch.from_numpy(pred).to(DEV), csr.ptr, csr.ids, ks)
    b = ops.topk_metrics(torch.from_numpy(pred).to(DEV), csr.ptr, csr.ids, ks)
    assert torch.equal(a, b)
    for mi, m in enumerate(M.METRICS):
        assert np.allclose(a[:, mi].cpu().numpy(), ref[m], rtol=0, atol=1e-11), m
    # 4M rows (no Python loop over users anywhere): the mean over a tiled table equals the mean over the tile
    big = torch.from_numpy(pred).to(DEV).repeat(200, 1)
    ptr = torch.cat([csr.ptr[:-1] + r * csr.ptr[-1] for r in range(200)] + [csr.ptr[-1:] * 200])
    big_res = ops.topk_metrics(big, ptr, csr.ids.repeat(200), ks)
    assert torch.allclose(big_res, a, rtol=0, atol=1e-10)


# ------------------------------------------------------------------------------------------------ ADVICE: sentinel rows
def test_fused_bpr_skips_sampler_sentinel_rows(ops):
    from textgcn_b200.models import _FusedBprFn
    g = load_golden("small_lgcn_d64")
    graph = ops.Graph.from_norm_matrix(O.sparse_tensor(g["norm_row"], g["norm_col"], g["norm_val"], int(g["n_users"] + g["n_items"])).to(DEV),
                                       int(g["n_users"]), int(g["n_items"]))
    batch = torch.from_numpy(g["batch"]).clone()
    b = batch.shape[0]

    def run(rows):
        uw = torch.from_numpy(g["user_w"]).to(DEV).requires_grad_(True)
        iw = torch.from_numpy(g["item_w"]).to(DEV).requires_grad_(True)
        losses = _FusedBprFn.apply(uw, iw, graph, 3, False, None, 0.0, ops.as_index(rows[:, 0], DEV), ops.as_index(rows[:, 1], DEV),
                                   ops.as_index(rows[:, 2:].t(), DEV), 1e-4)
        losses.sum().backward()
        return losses.detach().cpu().double(), uw.grad.cpu(), iw.grad.cpu()

    full_l, full_gu, full_gi = run(batch)
    bad = batch.clone()
    bad[::3, 1] = -1             # sampler sentinel: no positive (tgcn_sample_bpr_batch)
    bad[1::3, 2] = -1            # sentinel first negative only: that negative is skipped
    bad[5, 0] = 10 ** 6          # out-of-range user id
    l, gu, gi = run(bad)
    assert torch.isfinite(l).all() and torch.isfinite(gu).all() and torch.isfinite(gi).all()
    keep = torch.ones(b, dtype=torch.bool)
    keep[::3] = False
    keep[5] = False
    # oracle on the surviving work: per-row terms keep the divisors batch and batch * n_neg
    n_neg = batch.shape[1] - 2
    norm = O.sparse_tensor(g["norm_row"], g["norm_col"], g["norm_val"], int(g["n_users"] + g["n_items"]))
    ue, ie = O.propagate(norm, torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"]), 3)
    tot = 0.0
    for r in range(b):
        if not keep[r]:
            continue
        u, p = int(batch[r, 0]), int(batch[r, 1])
        for j in range(n_neg):
            if r % 3 == 1 and j == 0:
                continue
            x = float((ue[u].double() * (ie[int(batch[r, 2 + j])].double() - ie[p].double())).sum())
            tot += float(torch.nn.functional.selu(torch.tensor(x, dtype=torch.float64)))
    assert abs(float(l[0]) - tot / (b * n_neg)) <= 1e-5 * abs(tot / (b * n_neg))
    assert float(l[0]) != float(full_l[0])


# ------------------------------------------------------------------------------------------------ ADVICE: --load_base (G18)
@pytest.mark.parametrize("cls_name", ["LTRLinear", "LTRLinearWPop"])
def test_standalone_ltr_load_base_freeze(ops, cls_name, tmp_path):
    import textgcn_b200.models as MD
    g = load_golden("small_ltr_pop")
    ds = StubDataset(g, DEV)
    base = MD.BaseModel(params_from_golden(g, save=True, save_path=str(tmp_path / "base")), ds)
    load_weights(base, g)
    base_metrics = base.evaluate()
    base.checkpoint(1)
    assert os.path.exists(tmp_path / "base" / "best.pkl")
    seen = []

    class Rec(logging.Handler):
        def emit(self, record):
            seen.append(record.getMessage())

    log = logging.getLogger("test_load_base")
    log.setLevel(logging.INFO)
    log.addHandler(Rec())
    model = getattr(MD, cls_name)(params_from_golden(g, load_base=str(tmp_path / "base"), freeze=True, logger=log), ds)
    # loaded inside _add_vars, before the head existed, and evaluated with plain LightGCN scoring
    assert any("Performance of the loaded model" in m for m in seen)
    logged = [m for m in seen if m.startswith("recall")]
    assert logged and logged[0].split()[1:] == [f"{v:.4f}" for v in base_metrics["recall"]]
    assert torch.equal(model.embedding_user.weight, base.embedding_user.weight) and not model.embedding_user.weight.requires_grad
    assert model.metrics_logger["recall"].shape[0] == 0
    assert "score_batchwise" in model.__dict__ and model._ltr_active()
    ids, _ = model.predict_device(g["test_users"])                      # LTR scoring from here on
    assert ids.shape == (len(g["test_users"]), int(max(g["ks"])))
    model.training = True
    loss = model.get_loss(torch.from_numpy(g["batch"]))
    loss.backward()
    assert model.embedding_user.weight.grad is None and model.layers[0].weight.grad is not None


# ------------------------------------------------------------------------------------------------ headline shapes
def test_eval_kernel_at_c5_shape_against_fp64(ops):
    """C5's eval shape — 2M items, d = 128, k = 20, two item splits per user tile — on random tables: a 256-user sample against
    an fp64 ranking (tie-aware), structural properties on all 19 200 rows."""
    from textgcn_b200.graph import graph_from_interactions
    nu, ni, d, k = 19200, 2_000_000, 128, 20
    rng = np.random.default_rng(0)
    tu = np.concatenate([np.arange(nu), rng.integers(nu, size=400_000)])
    ti = np.concatenate([rng.integers(ni, size=nu), rng.integers(ni, size=400_000)])
    graph = graph_from_interactions(tu, ti, nu, ni, DEV)
    uv, iv = _randn(nu, d, seed=1) * 0.1, _randn(ni, d, seed=2) * 0.1
    ids, sc = ops.eval_topk(graph, uv, iv, k)
    assert bool((sc[:, 1:] <= sc[:, :-1]).all()) and bool(((ids >= 0) & (ids < ni)).all())
    sample = torch.from_numpy(rng.choice(nu, 256, replace=False)).to(DEV)
    dense = uv[sample].double() @ iv.double().T
    tl = O.train_lists_from_edges(tu, ti, nu)
    for r, u in enumerate(sample.tolist()):
        dense[r, torch.from_numpy(tl[u]).to(DEV)] = float("-inf")
    top = torch.topk(dense, k + 8, dim=1)
    # canonical order (score desc, id asc) of the fp64 reference, then the tie-aware comparison on fp32-rounded scores
    o_sc, o_ids = top.values.cpu().numpy(), top.indices.cpu().numpy()
    order = np.lexsort((o_ids, -o_sc), axis=1)
    o_sc, o_ids = np.take_along_axis(o_sc, order, 1)[:, :k], np.take_along_axis(o_ids, order, 1)[:, :k]
    st = O.topk_lists_equivalent(ids[sample].cpu().numpy().astype(np.int64), sc[sample].cpu().numpy(), o_ids, o_sc.astype(np.float32),
                                 rtol=1e-5, atol=1e-6)
    assert st["bad"] == 0 and st["exact"] >= 240, st
    assert float((sc[sample].double().cpu() - torch.from_numpy(o_sc)).abs().max()) <= 1e-5 * float(uv.norm(dim=1).max() * iv.norm(dim=1).max())


def test_spmm_with_tables_beyond_4_gib(ops):
    """Byte offsets past 2^32 in the gather source, the output and the CSR positions' address arithmetic: 9M nodes x d = 128
    (4.6 GB per table).  Sampled rows against an fp64 gather-and-sum, including rows whose table offset exceeds 4 GiB."""
    from textgcn_b200.graph import norm_adj_csr
    nu, ni, d = 8_500_000, 500_000, 128
    gen = torch.Generator(device=DEV).manual_seed(0)
    tu = torch.cat([torch.arange(nu, device=DEV), torch.randint(0, nu, (6_000_000,), generator=gen, device=DEV)])
    ti = torch.cat([torch.randint(0, ni, (nu,), generator=gen, device=DEV), torch.randint(0, ni, (6_000_000,), generator=gen, device=DEV)])
    rowptr, col, val = norm_adj_csr(tu, ti, nu, ni)
    del tu, ti
    graph = ops.Graph(nu, ni, rowptr.contiguous(), col.contiguous(), val.contiguous())
    x = torch.randn(nu + ni, d, generator=gen, device=DEV)
    assert x.numel() * 4 > 2 ** 32
    y = ops.spmm(graph, x)
    rows = torch.cat([torch.randint(0, nu + ni, (2000,), generator=gen, device=DEV),
                      torch.arange(nu + ni - 500, nu + ni, device=DEV),            # item rows: output offset > 4 GiB, long rows
                      torch.arange(nu - 500, nu, device=DEV)])                     # last user rows
    rp = rowptr.to(torch.int64)
    worst = 0.0
    for r in rows.tolist():
        lo, hi = int(rp[r]), int(rp[r + 1])
        ref = (val[lo:hi].double()[:, None] * x[col[lo:hi].long()].double()).sum(0)
        worst = max(worst, float((y[r].double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
    assert worst < TOL, worst
    # and through the fused propagate (layer mean epilogue reads E0 and the workspace layer beyond 4 GiB)
    out = ops.propagate_fwd(graph, x[:nu], x[nu:], 2)
    y2 = ops.spmm(graph, y)
    ref_rows = (x[rows] + y[rows] + y2[rows]) / 3
    assert rel_err(out[rows].cpu(), ref_rows.cpu()) < TOL


def test_synthetic_graph_is_bit_reproducible_on_the_device():
    """Every rank of a multi-GPU run builds the workload itself: two builds must agree bit for bit (round 1's float64 device
    cumsum did not, and ranks ended up with slightly different graphs)."""
    from textgcn_b200.graph import norm_adj_csr
    from textgcn_b200.synthetic import interactions
    builds = []
    for _ in range(3):
        tu, ti = interactions(2_000_000, 400_000, 30_000_000, torch.device(DEV), seed=0)
        builds.append((tu, ti) + tuple(norm_adj_csr(tu, ti, 2_000_000, 400_000)))
    for other in builds[1:]:
        assert all(torch.equal(a, b) for a, b in zip(builds[0], other))
    tu, ti = builds[0][:2]
    assert int(torch.unique(tu).numel()) == 2_000_000 and int(torch.unique(ti).numel()) == 400_000 and tu.numel() == 30_000_000


@pytest.mark.parametrize("case,single", [("small_lgcn_d64", False), ("small_lgcn_d128_l4", False), ("small_lgcn_d32_single", True)])
def test_propagate_host_pipelined_matches_device_path(ops, case, single):
    """tgcn_propagate_host (the e2e entry: uploads on a copy stream, layer 1's user pass overlapping the user-table upload, last layer
    in row chunks with the download behind it) returns exactly what tgcn_propagate_fwd computes from device-resident tables."""
    g = load_golden(case)
    nu, ni, L = int(g["n_users"]), int(g["n_items"]), int(g["n_layers"])
    graph = ops.Graph.from_norm_matrix(O.sparse_tensor(g["norm_row"], g["norm_col"], g["norm_val"], nu + ni).to(DEV), nu, ni)
    uw, iw = torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"])
    d = uw.shape[1]
    want = ops.propagate_fwd(graph, uw.to(DEV), iw.to(DEV), L, single=single)
    h_u, h_i = uw.clone().pin_memory(), iw.clone().pin_memory()
    h_o = torch.full((nu + ni, d), float("nan")).pin_memory()
    stage = torch.empty((2 * (nu + ni), d), dtype=torch.float32, device=DEV)
    for _ in range(3):   # back to back: the staging buffers and events are reused
        ops.propagate_host(graph, h_u, h_i, h_o, L, stage, single=single)
    torch.cuda.synchronize()
    assert torch.equal(h_o, want.cpu())
    assert rel_err(h_o[:nu].numpy(), g["rep_user"]) < TOL and rel_err(h_o[nu:].numpy(), g["rep_item"]) < TOL


def test_propagate_host_at_electronics_shape_with_long_rows(ops):
    from textgcn_b200.graph import graph_from_interactions
    nu, ni, ne, d, L = 190_000, 63_000, 1_700_000, 64, 3
    tu, ti = O.synthetic_interactions(nu, ni, ne, seed=0)
    graph = graph_from_interactions(tu, ti, nu, ni, DEV)
    assert graph.n_segments > 0                      # long item rows: the segment ranges of the chunked launches are exercised
    gen = torch.Generator().manual_seed(1)
    uw, iw = torch.randn(nu, d, generator=gen) * 0.1, torch.randn(ni, d, generator=gen) * 0.1
    want = ops.propagate_fwd(graph, uw.to(DEV), iw.to(DEV), L)
    h_u, h_i = uw.pin_memory(), iw.pin_memory()
    h_o = torch.empty((nu + ni, d)).pin_memory()
    stage = torch.empty((2 * (nu + ni), d), dtype=torch.float32, device=DEV)
    ops.propagate_host(graph, h_u, h_i, h_o, L, stage)
    torch.cuda.synchronize()
    assert torch.equal(h_o, want.cpu())
