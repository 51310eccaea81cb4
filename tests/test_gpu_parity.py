"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the reference's golden vectors.

Bars (north_star): propagated embeddings / scores / gradients within 1e-5 relative (norm-wise); top-k lists
bit-exact under the canonical order (score desc, item id asc) when scores are exact, tie-aware otherwise;
Recall / NDCG / Precision identical.
"""
import numpy as np
import pytest
import torch

from conftest import LGCN_CASES, load_golden
from helpers import TOL, StubDataset, golden_lists, golden_norm, load_weights, params_from_golden, rel_err
from oracle import lightgcn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from textgcn_b200 import ops as _ops
    return _ops


def _graph(ops, g):
    return ops.Graph.from_norm_matrix(golden_norm(g).to(DEV), int(g["n_users"]), int(g["n_items"]))


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


# ------------------------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("case", LGCN_CASES)
def test_norm_adj_device_build_is_bit_exact(ops, case):
    from textgcn_b200.graph import norm_adj_csr
    g = load_golden(case)
    rowptr, col, val = norm_adj_csr(_cuda(g["train_u"]), _cuda(g["train_i"]), int(g["n_users"]), int(g["n_items"]))
    assert np.array_equal(col.cpu().numpy(), g["norm_col"])
    assert np.array_equal(val.cpu().numpy().view(np.uint32), g["norm_val"].view(np.uint32))
    assert np.array_equal(rowptr.cpu().numpy(), O.coo_to_csr(g["norm_row"], int(g["n_users"] + g["n_items"])))


# ------------------------------------------------------------------------------------------------ a2-a6
@pytest.mark.parametrize("case", LGCN_CASES)
def test_propagate_matches_reference_golden(ops, case):
    g = load_golden(case)
    gr = _graph(ops, g)
    out = ops.propagate_fwd(gr, _cuda(g["user_w"]), _cuda(g["item_w"]), int(g["n_layers"]), bool(g["single"])).cpu().numpy()
    nu = int(g["n_users"])
    assert rel_err(out[:nu], g["rep_user"]) < TOL
    assert rel_err(out[nu:], g["rep_item"]) < TOL


@pytest.mark.parametrize("d", [16, 32, 48, 64, 128, 256, 320])
def test_spmm_all_widths_and_long_rows(ops, d):
    # hub items with > 512 interactions exercise the segment + fix-up path
    rng = np.random.default_rng(d)
    nu, ni = 3000, 40
    u = np.concatenate([np.arange(nu), np.arange(0, nu, 2), rng.integers(nu, size=4000)])
    i = np.concatenate([np.zeros(nu, np.int64), np.ones(nu // 2, np.int64), rng.integers(2, ni, size=4000)])
    i[-ni:] = np.arange(ni)
    row, col, val = O.norm_adj_coo(u, i, nu, ni)
    norm = O.sparse_tensor(row, col, val, nu + ni)
    gr = ops.Graph.from_norm_matrix(norm.to(DEV), nu, ni)
    assert gr.n_segments > 0
    x = torch.randn(nu + ni, d, generator=torch.Generator().manual_seed(1))
    y = ops.spmm(gr, x.to(DEV)).cpu()
    ref = torch.sparse.mm(norm.double(), x.double())
    assert rel_err(y.numpy(), ref.numpy()) < 2e-6
    # fused epilogue: y = (x + Â·x) / 2, then accumulate
    y2 = torch.empty_like(x, device=DEV)
    ops.spmm_ex(gr, x.to(DEV), y2, addends=[x.to(DEV)], divisor=2.0)
    assert rel_err(y2.cpu().numpy(), ((x.double() + ref) / 2).numpy()) < 2e-6
    ops.spmm_ex(gr, x.to(DEV), y2, accumulate=True)
    assert rel_err(y2.cpu().numpy(), ((x.double() + ref) / 2 + ref).numpy()) < 2e-6


def test_propagate_dropout_forward_and_adjoint(ops):
    g = load_golden("small_lgcn_d64")
    gr = _graph(ops, g)
    keep = torch.from_numpy(g["train_keep"])
    p = float(g["dropout"])
    uw, iw = torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"])
    ref_u, ref_i = O.propagate(golden_norm(g), uw, iw, 3, keep_mask=keep, dropout=p)
    out = ops.propagate_fwd(gr, uw.to(DEV), iw.to(DEV), 3, keep=keep.to(DEV), dropout=p).cpu()
    assert rel_err(out.numpy(), torch.cat([ref_u, ref_i]).numpy()) < TOL
    # <P x, y> == <x, Pᵀ y> with the same mask: checks the transpose permutation path
    gen = torch.Generator().manual_seed(3)
    x, y = torch.randn(190, 64, generator=gen), torch.randn(190, 64, generator=gen)
    px = ops.propagate_fwd(gr, x[:120].contiguous().to(DEV), x[120:].contiguous().to(DEV), 3, keep=keep.to(DEV), dropout=p)
    pty = ops.propagate_bwd(gr, y.to(DEV), 3, keep=keep.to(DEV), dropout=p)
    lhs = float((px.double().cpu() * y.double()).sum())
    rhs = float((x.double() * pty.double().cpu()).sum())
    assert abs(lhs - rhs) < 1e-5 * max(abs(lhs), 1.0)


# ------------------------------------------------------------------------------------------------ a7-a10
@pytest.mark.parametrize("case", LGCN_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_fused_bpr_loss_and_grads_match_reference(ops, case, mode):
    from textgcn_b200.models import _FusedBprFn
    g = load_golden(case)
    gr = _graph(ops, g)
    keep = _cuda(g["train_keep"]) if mode == "train" else None
    uw = _cuda(g["user_w"]).requires_grad_(True)
    iw = _cuda(g["item_w"]).requires_grad_(True)
    batch = torch.from_numpy(g["batch"])
    users, pos, negs = (ops.as_index(batch[:, 0], DEV), ops.as_index(batch[:, 1], DEV), ops.as_index(batch[:, 2:].t(), DEV))
    losses = _FusedBprFn.apply(uw, iw, gr, int(g["n_layers"]), bool(g["single"]), keep, float(g["dropout"]),
                               users, pos, negs, float(g["reg_lambda"]))
    (losses[0] + losses[1]).backward()
    bpr, reg = losses.tolist()
    # the loss is a mean of O(1e-2) SELU terms of both signs: the bound is norm-wise (1e-5 of the term scale)
    assert abs(bpr - float(g[f"{mode}_bpr"])) <= TOL * abs(float(g[f"{mode}_bpr"])) + 1e-7
    assert abs(reg - float(g[f"{mode}_reg"])) <= TOL * abs(float(g[f"{mode}_reg"])) + 1e-10
    assert rel_err(uw.grad.cpu().numpy(), g[f"{mode}_grad_embedding_user_weight"]) < TOL
    assert rel_err(iw.grad.cpu().numpy(), g[f"{mode}_grad_embedding_item_weight"]) < TOL


def test_propagate_autograd_function(ops):
    from textgcn_b200.models import _PropagateFn
    g = load_golden("small_lgcn_d32_single")
    gr = _graph(ops, g)
    uw = _cuda(g["user_w"]).requires_grad_(True)
    iw = _cuda(g["item_w"]).requires_grad_(True)
    out = _PropagateFn.apply(uw, iw, gr, 2, True, None, 0.0)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(0)).to(DEV)
    (out * w).sum().backward()
    uw2 = torch.from_numpy(g["user_w"]).double().requires_grad_(True)
    iw2 = torch.from_numpy(g["item_w"]).double().requires_grad_(True)
    ue, ie = O.propagate(golden_norm(g).double(), uw2, iw2, 2, single=True)
    (torch.cat([ue, ie]) * w.cpu().double()).sum().backward()
    assert rel_err(uw.grad.cpu().numpy(), uw2.grad.numpy()) < TOL
    assert rel_err(iw.grad.cpu().numpy(), iw2.grad.numpy()) < TOL


# ------------------------------------------------------------------------------------------------ a11-a13
@pytest.mark.parametrize("case", LGCN_CASES)
def test_predict_topk_matches_reference_golden(ops, case):
    from textgcn_b200 import metrics as M
    g = load_golden(case)
    gr = _graph(ops, g)
    nu = int(g["n_users"])
    emb = torch.cat([_cuda(g["rep_user"]), _cuda(g["rep_item"])]).contiguous()
    kmax = int(max(g["ks"]))
    users = ops.as_index(g["test_users"], DEV)
    ids, sc = ops.eval_topk(gr, emb[:nu], emb[nu:], kmax, users=users)
    ids, sc = ids.cpu().numpy().astype(np.int64), sc.round(decimals=4).cpu().numpy()
    tl = golden_lists(g)
    o_ids, o_sc = O.predict_topk(torch.from_numpy(g["rep_user"]), torch.from_numpy(g["rep_item"]), g["test_users"], tl, kmax)
    st = O.topk_lists_equivalent(ids, sc, o_ids, o_sc, rtol=1e-5, atol=1.01e-4)
    assert st["bad"] == 0, st
    assert st["exact"] >= st["rows"] - 2, st
    # against the reference's own lists: same ids wherever its scores are finite
    fin = np.isfinite(g["pred_scores"])
    assert np.array_equal(np.isfinite(sc), fin)
    assert (ids[fin] != g["pred_ids"][fin]).sum() <= 2
    assert np.abs(sc[fin] - g["pred_scores"][fin]).max() <= 1.01e-4
    # metrics identical to the reference's calculate_metrics output
    test = golden_lists(g, "test")
    res = M.calculate_metrics(torch.from_numpy(ids).to(DEV), [test[u].tolist() for u in g["test_users"]], g["ks"].tolist())
    if st["exact"] == st["rows"]:
        for m in M.METRICS:
            assert np.allclose(res[m], g["metric_" + m], rtol=0, atol=1e-12), m


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "screen"])
@pytest.mark.parametrize("n_rank,n_items,d,k", [(300, 5000, 64, 20), (40, 20000, 128, 40), (1000, 700, 32, 64), (5, 130, 64, 7),
                                                (130, 1000, 96, 20), (700, 40000, 128, 20), (257, 3000, 100, 24)])
def test_topk_bit_exact_on_exact_arithmetic(ops, n_rank, n_items, d, k, precision):
    """Dyadic-grid embeddings make every dot product exact in fp32 in any summation order, so the lists must be
    bit-identical to the canonical order, ties (plentiful here) included (SURVEY.md §8c iv).  For the screened path the
    ties are the hard case: whole plateaus sit inside the re-scoring band, and rows whose plateau outgrows the 40-entry
    list must come back exact from the second pass."""
    if precision == "screen" and k > 24:
        pytest.skip("screened path: k <= 24")
    rng = np.random.default_rng(n_items)
    nu = n_rank + 17
    ue = (rng.integers(-32, 33, size=(nu, d)) / 16).astype(np.float32)
    ie = (rng.integers(-8, 9, size=(n_items, d)) / 16).astype(np.float32)
    tu, ti = O.synthetic_interactions(nu, n_items, max(nu, n_items) * 3, seed=1)
    row, col, val = O.norm_adj_coo(tu, ti, nu, n_items)
    gr = ops.Graph.from_norm_matrix(O.sparse_tensor(row, col, val, nu + n_items).to(DEV), nu, n_items)
    users = rng.permutation(nu)[:n_rank]
    ids, sc = ops.eval_topk(gr, _cuda(ue), _cuda(ie), k, users=ops.as_index(users, DEV), precision=precision)
    tl = O.train_lists_from_edges(tu, ti, nu)
    o_ids, o_sc = O.predict_topk(torch.from_numpy(ue), torch.from_numpy(ie), users, tl, k, round_decimals=None)
    assert np.array_equal(ids.cpu().numpy(), o_ids)
    assert np.array_equal(sc.cpu().numpy(), o_sc)
    # item-sharded evaluation + merge gives the same table (multi-GPU eval path)
    cuts = [0, n_items // 3, n_items // 2, n_items]
    parts = [ops.eval_topk(gr, _cuda(ue), _cuda(ie), k, users=ops.as_index(users, DEV), item_range=(a, b), finalize=False,
                           precision=precision)
             for a, b in zip(cuts[:-1], cuts[1:])]
    m_ids, m_sc = ops.topk_merge(gr, torch.stack([p[0] for p in parts]).contiguous(),
                                 torch.stack([p[1] for p in parts]).contiguous(), users=ops.as_index(users, DEV))
    assert np.array_equal(m_ids.cpu().numpy(), o_ids) and np.array_equal(m_sc.cpu().numpy(), o_sc)


@pytest.mark.parametrize("n_rank,n_items,d,k", [(150, 1200, 48, 20), (64, 3000, 160, 10), (130, 700, 1600, 20), (40, 900, 260, 33)])
def test_3xtf32_wide_and_ragged_contractions_with_bias(ops, n_rank, n_items, d, k):
    """Zero-padded K (not a multiple of 32), the streamed-user-tile variant (K > 128, the LTR width 1600) and the bias
    chunk: bit-exact on the dyadic fixture (bias values dyadic too), against both the oracle and the fp32 kernel."""
    rng = np.random.default_rng(d)
    ue = (rng.integers(-16, 17, size=(n_rank, d)) / 16).astype(np.float32)
    ie = (rng.integers(-8, 9, size=(n_items, d)) / 16).astype(np.float32)
    ub = (rng.integers(-64, 65, size=n_rank) / 8).astype(np.float32)
    ib = (rng.integers(-64, 65, size=n_items) / 8).astype(np.float32)
    ref = ue.astype(np.float64) @ ie.astype(np.float64).T
    for bias in (False, True):
        full = ref + (ub[:, None].astype(np.float64) + ib[None, :] if bias else 0.0)
        o_ids, o_sc = O.canonical_topk(full, k)
        kw = dict(user_bias=_cuda(ub), item_bias=_cuda(ib)) if bias else {}
        for precision in ("3xtf32", "fp32") + (("screen",) if k <= 24 else ()):
            ids, sc = ops.eval_topk(None, _cuda(ue), _cuda(ie), k, precision=precision, **kw)
            assert np.array_equal(ids.cpu().numpy(), o_ids), (bias, precision)
            assert np.array_equal(sc.cpu().numpy(), o_sc.astype(np.float32)), (bias, precision)


@pytest.mark.parametrize("d,n_items,k", [(64, 9000, 20), (128, 3000, 40), (32, 1500, 20), (1600, 2500, 20)])
def test_3xtf32_scores_within_stated_tolerance(ops, d, n_items, k):
    """The tensor-core path (3xTF32) against fp64: |Δscore| <= 1e-5·‖u‖‖i‖, ids tie-aware."""
    gen = torch.Generator().manual_seed(d)
    nu = 700
    ue = torch.randn(nu, d, generator=gen) * 0.3
    ie = torch.randn(n_items, d, generator=gen) * 0.3
    ids, sc = ops.eval_topk(None, ue.to(DEV), ie.to(DEV), k, precision="3xtf32")
    ids32, sc32 = ops.eval_topk(None, ue.to(DEV), ie.to(DEV), k, precision="fp32")
    ref = (ue.double() @ ie.double().T).numpy()
    o_ids, o_sc = O.canonical_topk(ref, k)
    bound = 1e-5 * float(ue.norm(dim=1).max() * ie.norm(dim=1).max())
    assert np.abs(sc.cpu().numpy() - o_sc).max() <= bound
    assert np.abs(sc32.cpu().numpy() - o_sc).max() <= bound
    got = [(ids, sc), (ids32, sc32)]
    if d <= 128 and k <= 24:  # the screened path: 1xTF32 candidates, exact fp32 scores
        ids_s, sc_s = ops.eval_topk(None, ue.to(DEV), ie.to(DEV), k, precision="screen")
        assert np.abs(sc_s.cpu().numpy() - o_sc).max() <= bound
        got.append((ids_s, sc_s))
    for got_ids, got_sc in got:
        st = O.topk_lists_equivalent(got_ids.cpu().numpy().astype(np.int64), got_sc.cpu().numpy(), o_ids, o_sc.astype(np.float32),
                                     rtol=0, atol=2 * bound)
        assert st["bad"] == 0 and st["exact"] >= st["rows"] - 5, st


def test_topk_short_lists_are_completed_with_masked_items(ops):
    g = load_golden("dummy_lgcn")  # 4 items, user 0 has 3 of them in train: only 1 rankable item (G9)
    gr = _graph(ops, g)
    emb = torch.cat([_cuda(g["rep_user"]), _cuda(g["rep_item"])]).contiguous()
    ids, sc = ops.eval_topk(gr, emb[:5], emb[5:], 3)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    assert ids[0].tolist() == [3, 0, 1] and np.isinf(sc[0, 1:]).all() and sc[0, 0] == pytest.approx(0.0724, abs=1e-4)
    assert ids[1].tolist()[:2] == [2, 0] and ids[1, 2] == 1 and np.isneginf(sc[1, 2])


# ------------------------------------------------------------------------------------------------ a14-a16
def test_adv_select_matches_reference(ops):
    g = load_golden("small_adv")
    gr = _graph(ops, g)
    data = torch.from_numpy(g["data"])
    emb = ops.propagate_fwd(gr, _cuda(g["user_w"]), _cuda(g["item_w"]), int(g["n_layers"]))
    kmax = int(max(g["ks"]))
    negs, counts, scores = ops.adv_select(gr, emb, ops.as_index(data[:, 0], DEV), ops.as_index(data[:, 1:], DEV), kmax, want_scores=True)
    assert rel_err(scores.cpu().numpy(), g["rankings"]) < TOL
    ref = O.adv_select_negatives(torch.from_numpy(g["rankings"]), data[:, 1:], data[:, 0].numpy(), golden_lists(g), kmax)
    negs, counts = negs.cpu().numpy(), counts.cpu().numpy()
    for b, r in enumerate(ref):
        assert counts[b] == len(r)
        assert np.array_equal(negs[b, :len(r)], r) and (negs[b, len(r):] == -1).all()


def test_adv_model_builds_the_reference_triples_and_loss(ops):
    from textgcn_b200.models import AdvSamplModel
    g = load_golden("small_adv")
    model = AdvSamplModel(params_from_golden(g), StubDataset(g, DEV))
    load_weights(model, g)
    model.training = False
    triples = model.select_triples(torch.from_numpy(g["data"]), sampled_pos=torch.from_numpy(g["sampled_pos"]))
    assert np.array_equal(triples.cpu().numpy(), g["triples"])
    model.zero_grad()
    from textgcn_b200.models import B200HotPath
    loss = B200HotPath.get_loss(model, triples)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    assert rel_err(model.embedding_user.weight.grad.cpu().numpy(), g["grad_user"]) < TOL
    assert rel_err(model.embedding_item.weight.grad.cpu().numpy(), g["grad_item"]) < TOL


# ------------------------------------------------------------------------------------------------ a17-a21
@pytest.mark.parametrize("case,cls_name", [("small_ltr_linear", "LTRLinear"), ("small_ltr_pop", "LTRLinearWPop")])
def test_ltr_models_match_reference(ops, case, cls_name):
    import textgcn_b200.models as MD
    from textgcn_b200 import metrics as M
    g = load_golden(case)
    n_head = int(g["n_head_layers"])
    ltr_layers = [int(g[f"head_w{i}"].shape[0]) for i in range(n_head - 1)]
    model = getattr(MD, cls_name)(params_from_golden(g, ltr_layers=ltr_layers), StubDataset(g, DEV))
    load_weights(model, g)
    model.training = False
    nu = int(g["n_users"])
    with torch.no_grad():
        ue, ie = model.representation
        emb = torch.cat([ue, ie]).contiguous()
        # pairwise features + head vs the reference's score_pairwise_ltr
        f = model.get_features_pairwise_fused(emb, torch.from_numpy(g["pair_users"]), torch.from_numpy(g["pair_items"]))
        sp = model.layers(f).cpu().numpy()
        assert sp.shape == g["score_pairwise"].shape
        assert np.abs(sp - g["score_pairwise"]).max() <= TOL * np.abs(g["score_batchwise"]).max()
    # full ranking through the collapsed single contraction vs the reference's dense score matrix
    kmax = int(max(g["ks"]))
    ids, sc = model.predict_device(g["test_users"])
    tl = golden_lists(g)
    dense = g["score_batchwise"][g["test_users"]].copy()
    for r, u in enumerate(g["test_users"]):
        dense[r, tl[u]] = -np.inf
    o_ids, o_sc = O.canonical_topk(dense, kmax)
    scale = np.abs(g["score_batchwise"]).max()
    st = O.topk_lists_equivalent(ids.cpu().numpy().astype(np.int64), sc.cpu().numpy(), o_ids, o_sc, rtol=0, atol=1.01e-4 + 2 * TOL * scale)
    assert st["bad"] == 0, st
    if st["exact"] == st["rows"]:
        test = golden_lists(g, "test")
        res = M.calculate_metrics(ids, [test[u].tolist() for u in g["test_users"]], g["ks"].tolist())
        for m in M.METRICS:
            assert np.allclose(res[m], g["metric_" + m], rtol=0, atol=1e-12), m
    # training loss + gradients (head and embeddings) vs the reference
    for mode in ("eval", "train"):
        model.zero_grad()
        model.training = mode == "train"
        model._loss_values = {"bpr": 0.0, "reg": 0.0}
        if mode == "train":
            keep = _cuda(g["train_keep"])
            model._draw_keep_mask = lambda keep=keep: keep
        loss = model.get_loss(torch.from_numpy(g["batch"]))
        loss.backward()
        assert abs(float(loss) - float(g[f"{mode}_loss"])) <= 2 * TOL * abs(float(g[f"{mode}_loss"])), mode
        for name, p in model.named_parameters():
            ref = g[f"{mode}_grad_" + name.replace(".", "_")]
            assert rel_err(p.grad.cpu().numpy(), ref) < 5 * TOL, (mode, name)
        model.__dict__.pop("_draw_keep_mask", None)


# ------------------------------------------------------------------------------------------------ model level
@pytest.mark.parametrize("case", ["dummy_lgcn", "small_lgcn_d64"])
def test_base_model_predict_evaluate_and_loss(ops, case):
    from textgcn_b200.models import BaseModel
    g = load_golden(case)
    model = BaseModel(params_from_golden(g), StubDataset(g, DEV))
    load_weights(model, g)
    ue, ie = model.representation
    assert rel_err(ue.detach().cpu().numpy(), g["rep_user"]) < TOL
    preds, scores = model.predict(g["test_users"], with_scores=True)
    assert isinstance(preds, list) and isinstance(preds[0], list) and len(preds[0]) == max(model.k)
    fin = np.isfinite(g["pred_scores"])
    assert (np.asarray(preds)[fin] != g["pred_ids"][fin]).sum() <= 2
    res = model.evaluate()
    if np.array_equal(np.asarray(preds)[fin], g["pred_ids"][fin]):
        for m in res:
            assert np.allclose(res[m], g["metric_" + m], rtol=0, atol=1e-12), m
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    model.training = False
    loss = model.get_loss(torch.from_numpy(g["batch"]))
    assert abs(float(loss) - float(g["eval_loss"])) <= TOL * abs(float(g["eval_loss"])) + 1e-9
    sd = model.state_dict()
    assert set(sd) == {"embedding_user.weight", "embedding_item.weight"}  # checkpoints interchange with the reference


def test_fit_reduces_loss_and_fused_adam_matches_torch(ops):
    from textgcn_b200.models import BaseModel
    g = load_golden("small_lgcn_d64")
    torch.manual_seed(0)
    batches = [torch.from_numpy(g["batch"])] * 4
    losses = {}
    for fused in (False, True):
        torch.manual_seed(0)
        model = BaseModel(params_from_golden(g, epochs=3, evaluate_every=3, fused_adam=fused, lr=1e-2, dropout=0.0), StubDataset(g, DEV))
        load_weights(model, g)
        model._loss_values = {"bpr": 0.0, "reg": 0.0}
        before = float(model.get_loss(batches[0]))
        model.fit(batches)
        model.training = False
        model._loss_values = {"bpr": 0.0, "reg": 0.0}
        after = float(model.get_loss(batches[0]))
        assert after < before
        losses[fused] = (after, model.embedding_user.weight.detach().cpu().numpy().copy())
    assert abs(losses[True][0] - losses[False][0]) <= 1e-4 * abs(losses[False][0])
    assert rel_err(losses[True][1], losses[False][1]) < 1e-4


def test_no_cpu_fallback():
    from textgcn_b200 import TgcnError
    from textgcn_b200 import ops as _ops
    with pytest.raises(TgcnError):
        _ops.spmm(None, torch.zeros(4, 4))  # CPU tensor must be refused, not silently computed


# ------------------------------------------------------------------------------------------------ full size
@pytest.fixture(scope="module")
def electronics(ops):
    """BASELINE.json configs[1]: Electronics-shaped synthetic graph (~190k users, ~63k items, 1.7M edges)."""
    from textgcn_b200.graph import graph_from_interactions
    nu, ni, ne = 190_000, 63_000, 1_700_000
    tu, ti = O.synthetic_interactions(nu, ni, ne, seed=0)
    gr = graph_from_interactions(tu, ti, nu, ni, DEV)
    torch.manual_seed(0)
    uw = (torch.randn(nu, 64) * 0.1)
    iw = (torch.randn(ni, 64) * 0.1)
    return dict(nu=nu, ni=ni, tu=tu, ti=ti, graph=gr, uw=uw, iw=iw)


def test_full_size_propagation_against_oracle(ops, electronics):
    e = electronics
    row, col, val = O.norm_adj_coo(e["tu"], e["ti"], e["nu"], e["ni"])
    assert np.array_equal(e["graph"].val.cpu().numpy().view(np.uint32), val.view(np.uint32))  # a1 bit-exact at full size
    norm = O.sparse_tensor(row, col, val, e["nu"] + e["ni"])
    ref = torch.cat(O.propagate(norm, e["uw"], e["iw"], 3))
    out = ops.propagate_fwd(e["graph"], e["uw"].to(DEV), e["iw"].to(DEV), 3)
    assert rel_err(out.cpu().numpy(), ref.numpy()) < TOL
    # linearity: P(2x - 3y) == 2P(x) - 3P(y)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(e["nu"] + e["ni"], 64, generator=gen).to(DEV)
    px = ops.propagate_fwd(e["graph"], x[:e["nu"]].contiguous(), x[e["nu"]:].contiguous(), 3)
    comb = 2 * x - 3 * torch.cat([e["uw"], e["iw"]]).to(DEV)
    pc = ops.propagate_fwd(e["graph"], comb[:e["nu"]].contiguous(), comb[e["nu"]:].contiguous(), 3)
    assert rel_err(pc.cpu().numpy(), (2 * px - 3 * out).cpu().numpy()) < TOL
    # adjoint identity with a dropout mask at full size
    keep = (torch.rand(e["graph"].nnz, generator=gen) < 0.6).to(DEV)
    pk = ops.propagate_fwd(e["graph"], x[:e["nu"]].contiguous(), x[e["nu"]:].contiguous(), 3, keep=keep, dropout=0.4)
    pt = ops.propagate_bwd(e["graph"], out, 3, keep=keep, dropout=0.4)
    lhs, rhs = float((pk.double() * out.double()).sum()), float((x.double() * pt.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)


def test_full_size_eval_properties_and_sample_against_oracle(ops, electronics):
    e = electronics
    emb = ops.propagate_fwd(e["graph"], e["uw"].to(DEV), e["iw"].to(DEV), 3)
    nu = e["nu"]
    ids, sc = ops.eval_topk(e["graph"], emb[:nu], emb[nu:], 20)
    ids_c, sc_c = ids.cpu().numpy().astype(np.int64), sc.cpu().numpy()
    assert np.all(np.diff(sc_c, axis=1) <= 0)                              # sorted descending
    assert np.all((ids_c >= 0) & (ids_c < e["ni"]))
    assert np.all(np.sort(ids_c, axis=1)[:, 1:] != np.sort(ids_c, axis=1)[:, :-1])  # no duplicates
    keys = set((e["tu"] * e["ni"] + e["ti"]).tolist())
    sample = np.random.default_rng(0).choice(nu, 256, replace=False)
    for u in sample:
        assert not any((int(u) * e["ni"] + int(i)) in keys for i in ids_c[u])         # train items excluded
    tl = O.train_lists_from_edges(e["tu"], e["ti"], nu)
    emb_c = emb.cpu()
    o_ids, o_sc = O.predict_topk(emb_c[:nu], emb_c[nu:], sample, tl, 20, round_decimals=None)
    st = O.topk_lists_equivalent(ids_c[sample], sc_c[sample], o_ids, o_sc, rtol=1e-5, atol=1e-7)
    assert st["bad"] == 0 and st["exact"] >= 250, st
    # merging a finished table with an empty part (all sentinels) changes nothing
    empty_ids = torch.full_like(ids, 2**31 - 1)
    empty_sc = torch.full_like(sc, float("-inf"))
    m_ids, m_sc = ops.topk_merge(e["graph"], torch.stack([empty_ids, ids]).contiguous(), torch.stack([empty_sc, sc]).contiguous())
    assert torch.equal(m_ids, ids) and torch.equal(m_sc, sc)


def test_full_size_adv_select_config3(ops, electronics):
    """BASELINE.json configs[2]: hardest-of-1000 negatives at the Electronics shape (B = 2048, k = 20)."""
    e = electronics
    emb = ops.propagate_fwd(e["graph"], e["uw"].to(DEV), e["iw"].to(DEV), 3)
    rng = np.random.default_rng(1)
    users = rng.integers(e["nu"], size=2048)
    cands = np.stack([rng.choice(e["ni"], size=1000, replace=False) for _ in users])
    negs, counts, scores = ops.adv_select(e["graph"], emb, ops.as_index(users, DEV), ops.as_index(cands, DEV), 20, want_scores=True)
    negs, counts, scores = negs.cpu().numpy(), counts.cpu().numpy(), scores.cpu()
    assert (counts == 20).all()
    tl = O.train_lists_from_edges(e["tu"], e["ti"], e["nu"])
    emb_c = emb.cpu()
    ref_rank = O.adv_rank_candidates(emb_c[:e["nu"]], emb_c[e["nu"]:], torch.from_numpy(users), torch.from_numpy(cands))
    assert rel_err(scores.numpy(), ref_rank.numpy()) < TOL
    # selection logic exactly, given the kernel's own scores (removes fp summation-order noise from the comparison)
    ref = O.adv_select_negatives(scores, torch.from_numpy(cands), users, tl, 20)
    for b in range(len(users)):
        assert np.array_equal(negs[b], ref[b])
        assert not np.isin(negs[b], tl[users[b]]).any()


def test_full_size_ltr_ranking_config4(ops, electronics):
    """BASELINE.json configs[3]: LTR ranking with random 768-d text tables through the collapsed single contraction."""
    e = electronics
    from textgcn_b200.models import LTRLinearWPop, make_params
    import logging
    nu, ni, D = e["nu"], e["ni"], 768
    gen = torch.Generator().manual_seed(4)

    class DS:
        pass

    ds = DS()
    ds.n_users, ds.n_items, ds.graph = nu, ni, e["graph"]
    ds.norm_matrix = None
    ds.test_users = np.arange(512)
    ds.true_test_lil = [[0]] * 512
    ds.items_as_avg_reviews = torch.randn(ni, D, generator=gen).to(DEV)
    ds.items_as_desc = torch.randn(ni, D, generator=gen).to(DEV)
    ds.users_as_avg_reviews = torch.randn(nu, D, generator=gen).to(DEV)
    ds.users_as_avg_desc = torch.randn(nu, D, generator=gen).to(DEV)
    ds.popularity_users = torch.rand(nu, 1, generator=gen).to(DEV)
    ds.popularity_items = torch.rand(ni, 1, generator=gen).to(DEV)
    model = LTRLinearWPop(make_params(k=[20], logger=logging.getLogger("t")), ds)
    with torch.no_grad():
        model.embedding_user.weight.copy_(e["uw"])
        model.embedding_item.weight.copy_(e["iw"])
    users = np.random.default_rng(2).choice(nu, 192, replace=False)
    ids, sc = model.predict_device(users)
    # oracle: the reference's 5 GEMMs + cat + Linear on the CPU, then mask + canonical top-k
    emb = ops.propagate_fwd(e["graph"], e["uw"].to(DEV), e["iw"].to(DEV), 3).cpu()
    tabs = {"users_rev": ds.users_as_avg_reviews.cpu(), "users_desc": ds.users_as_avg_desc.cpu(),
            "items_rev": ds.items_as_avg_reviews.cpu(), "items_desc": ds.items_as_desc.cpu()}
    ws = [l.weight.detach().cpu() for l in model.layers]
    bs = [l.bias.detach().cpu() for l in model.layers]
    ut = torch.from_numpy(users)
    dense = O.ltr_score_batchwise(emb[:nu][ut], emb[nu:], ut, tabs, ws, bs,
                                  (ds.popularity_users.cpu(), ds.popularity_items.cpu())).numpy()
    tl = O.train_lists_from_edges(e["tu"], e["ti"], nu)
    for r, u in enumerate(users):
        dense[r, tl[u]] = -np.inf
    o_ids, o_sc = O.canonical_topk(dense, 20)
    scale = float(np.abs(dense[np.isfinite(dense)]).max())
    st = O.topk_lists_equivalent(ids.cpu().numpy().astype(np.int64), sc.cpu().numpy(), o_ids, o_sc, rtol=0, atol=1.01e-4 + 2 * TOL * scale)
    assert st["bad"] == 0 and st["exact"] >= st["rows"] - 8, st


def test_gpu_sampler_properties_and_fit(ops):
    """n1: the GPU BPR sampler yields valid rows (positive in the user's train list, negatives outside it and distinct)
    with roughly uniform negatives, terminates on a user who interacted with almost every item (G20), and drives fit()."""
    from textgcn_b200.models import BaseModel
    from textgcn_b200.sampler import BprEpochSampler
    g = load_golden("small_lgcn_d64")
    gr = _graph(ops, g)
    tl = golden_lists(g)
    smp = BprEpochSampler(gr, batch_size=64, neg_samples=3, seed=7)
    rows = torch.cat(list(smp)).cpu().numpy()
    assert rows.shape == (smp.rows, 5) and len(smp) == (smp.rows + 63) // 64
    assert np.array_equal(np.bincount(rows[:, 0], minlength=gr.n_users), np.full(gr.n_users, smp.bucket_len))
    for u, p, *negs in rows:
        assert p in tl[u] and len(set(negs)) == 3 and not np.isin(negs, tl[u]).any()
    counts = np.bincount(rows[:, 2:].ravel(), minlength=gr.n_items)
    assert counts.min() > 0 and counts.max() < 4 * counts.mean()          # every item gets sampled, no gross skew
    assert int(smp.fail_count) == 0
    rows2 = torch.cat(list(BprEpochSampler(gr, batch_size=64, neg_samples=3, seed=7))).cpu().numpy()
    assert np.array_equal(rows, rows2)                                      # reproducible from the seed
    # dummy user 0 has 3 of 4 items: one negative exists, three cannot -> bounded, reported, no hang
    gd = load_golden("dummy_lgcn")
    grd = _graph(ops, gd)
    s1 = BprEpochSampler(grd, batch_size=8, neg_samples=1, seed=0)
    r1 = torch.cat(list(s1)).cpu().numpy()
    assert (r1[r1[:, 0] == 0][:, 2] == 3).all() and int(s1.fail_count) == 0
    s3 = BprEpochSampler(grd, batch_size=8, neg_samples=3, seed=0)
    r3 = torch.cat(list(s3)).cpu().numpy()
    assert (r3[r3[:, 0] == 0][:, 3:] == -1).all() and int(s3.fail_count) > 0
    # end to end: fit() over the sampler reduces the training loss
    model = BaseModel(params_from_golden(g, epochs=3, evaluate_every=3, lr=5e-3, fused_adam=True, dropout_rng="device"), StubDataset(g, DEV))
    load_weights(model, g)
    batch = torch.from_numpy(g["batch"][:, :3])
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    before = float(model.get_loss(batch))
    model.fit(BprEpochSampler(gr, batch_size=256, neg_samples=1, seed=1))
    model.training = False
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    assert float(model.get_loss(batch)) < before


def test_adv_device_samplers_and_fit(ops):
    """a14 / n1: candidate rows are [user, distinct uniform items]; device positives are distinct members of the user's
    train list; AdvSamplModel.fit() runs end to end on device-generated batches."""
    from textgcn_b200.models import AdvSamplModel
    from textgcn_b200.sampler import AdvEpochSampler
    g = load_golden("small_adv")
    gr = _graph(ops, g)
    tl = golden_lists(g)
    smp = AdvEpochSampler(gr, batch_size=32, seed=3)
    rows = torch.cat(list(smp)).cpu().numpy()
    assert rows.shape == (smp.rows, 1 + gr.n_items)                         # 70 items < 1000: every item, permuted
    assert np.array_equal(np.sort(rows[:, 1:], axis=1), np.tile(np.arange(gr.n_items), (len(rows), 1)))
    assert len({tuple(r) for r in rows[:, 1:].tolist()}) > len(rows) // 2      # rows get different permutations
    # larger item set: 1000 distinct candidates, roughly uniform
    tu, ti = O.synthetic_interactions(500, 5000, 20000, seed=2)
    row, col, val = O.norm_adj_coo(tu, ti, 500, 5000)
    big = ops.Graph.from_norm_matrix(O.sparse_tensor(row, col, val, 5500).to(DEV), 500, 5000)
    cands = torch.cat(list(AdvEpochSampler(big, batch_size=256, seed=1))).cpu().numpy()[:, 1:]
    assert cands.shape[1] == 1000 and cands.min() >= 0 and cands.max() < 5000
    assert all(len(np.unique(r)) == 1000 for r in cands[:200])
    freq = np.bincount(cands.ravel(), minlength=5000) / cands.shape[0]
    assert abs(freq.mean() - 0.2) < 1e-9 and freq.min() > 0.1 and freq.max() < 0.3
    # device positives
    model = AdvSamplModel(params_from_golden(g, epochs=2, evaluate_every=2, positive_sampler="device", dropout_rng="device", lr=5e-3),
                          StubDataset(g, DEV))
    load_weights(model, g)
    users = ops.as_index(np.arange(gr.n_users), DEV)
    pos = model._sample_positives_device(users).cpu().numpy()
    for u, prow in enumerate(pos):
        got = prow[prow >= 0]
        assert len(got) == min(5, len(tl[u])) and len(set(got)) == len(got) and np.isin(got, tl[u]).all()
    model.fit(AdvEpochSampler(gr, batch_size=64, seed=5))
    model.training = False
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    assert np.isfinite(float(model.get_loss(torch.from_numpy(g["data"]))))


def test_device_dropout_mask_statistics(ops):
    """n2: the device-drawn keep mask is Bernoulli(1 - p), reproducible from its seed, and differs between seeds."""
    n = 1_000_003
    for p in (0.0, 0.4, 0.9):
        k1 = ops.dropout_mask(n, p, 123, DEV)
        k2 = ops.dropout_mask(n, p, 123, DEV)
        k3 = ops.dropout_mask(n, p, 124, DEV)
        assert k1.dtype == torch.uint8 and set(torch.unique(k1).tolist()) <= {0, 1}
        assert torch.equal(k1, k2)
        frac = float(k1.float().mean())
        assert abs(frac - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n) + 1e-9
        if 0 < p < 1:
            assert not torch.equal(k1, k3)
            agree = float((k1 == k3).float().mean())
            assert abs(agree - (p * p + (1 - p) ** 2)) < 5e-3              # independent draws
            lag = float((k1[1:] == k1[:-1]).float().mean())
            assert abs(lag - (p * p + (1 - p) ** 2)) < 5e-3                # no serial correlation


@pytest.mark.parametrize("n_rank,n_items,d,k", [(200, 900, 64, 100), (150, 1200, 48, 65)])
def test_eval_auto_precision_falls_back_to_exact_kernel(ops, n_rank, n_items, d, k):
    """k > 64 is outside the tensor-core kernel (register-resident lists): "auto" must route to the exact fp32 kernel
    (bit-exact on the dyadic fixture) and an explicit "3xtf32" request must fail loudly, not silently."""
    from textgcn_b200 import TgcnError
    rng = np.random.default_rng(k)
    ue = (rng.integers(-32, 33, size=(n_rank, d)) / 16).astype(np.float32)
    ie = (rng.integers(-8, 9, size=(n_items, d)) / 16).astype(np.float32)
    ids, sc = ops.eval_topk(None, _cuda(ue), _cuda(ie), k)
    o_ids, o_sc = O.canonical_topk(ue.astype(np.float64) @ ie.astype(np.float64).T, k)
    assert np.array_equal(ids.cpu().numpy(), o_ids) and np.array_equal(sc.cpu().numpy(), o_sc.astype(np.float32))
    with pytest.raises(TgcnError):
        ops.eval_topk(None, _cuda(ue), _cuda(ie), k, precision="3xtf32")
