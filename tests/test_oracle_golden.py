"""CPU: the oracle restatement against golden vectors recorded from the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import LGCN_CASES, load_golden
from oracle import lightgcn_oracle as O


def _norm(g):
    n = int(g["n_users"] + g["n_items"])
    return O.sparse_tensor(g["norm_row"], g["norm_col"], g["norm_val"], n)


def _train_lists(g):
    return O.train_lists_from_edges(g["train_u"], g["train_i"], int(g["n_users"]))


@pytest.mark.parametrize("case", LGCN_CASES + ["small_adv", "small_ltr_linear"])
def test_norm_adj_bit_exact(case):
    g = load_golden(case)
    row, col, val = O.norm_adj_coo(g["train_u"], g["train_i"], int(g["n_users"]), int(g["n_items"]))
    assert np.array_equal(row, g["norm_row"]) and np.array_equal(col, g["norm_col"])
    assert np.array_equal(val.view(np.uint32), g["norm_val"].view(np.uint32))  # bit exact (G1)


def test_dummy_known_answers():
    g = load_golden("dummy_lgcn")
    row, col, val = O.norm_adj_coo(g["train_u"], g["train_i"], 5, 4)
    assert len(val) == 26
    dense = np.zeros((9, 9), np.float32)
    dense[row, col] = val
    assert dense[0, 5] == np.float32(1 / 3) and abs(dense[0, 6] - 1 / np.sqrt(12)) < 1e-7
    assert abs(dense[0, 7] - 1 / np.sqrt(6)) < 1e-7
    assert np.array_equal(dense, dense.T)


@pytest.mark.parametrize("case", LGCN_CASES)
def test_propagate_matches_reference(case):
    g = load_golden(case)
    ue, ie = O.propagate(_norm(g), torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"]),
                         int(g["n_layers"]), single=bool(g["single"]))
    assert np.array_equal(ue.numpy(), g["rep_user"]) and np.array_equal(ie.numpy(), g["rep_item"])
    # fp64 restatement bounds the fp32 rounding error well inside the 1e-5 parity budget
    ue64, ie64 = O.propagate(_norm(g).double(), torch.from_numpy(g["user_w"]).double(),
                             torch.from_numpy(g["item_w"]).double(), int(g["n_layers"]), single=bool(g["single"]))
    err = np.abs(ue64.numpy() - g["rep_user"]).max() / np.abs(g["rep_user"]).max()
    assert err < 1e-6


@pytest.mark.parametrize("case", LGCN_CASES)
def test_predict_and_metrics_match_reference(case):
    g = load_golden(case)
    ue, ie = torch.from_numpy(g["rep_user"]), torch.from_numpy(g["rep_item"])
    kmax = int(max(g["ks"]))
    ids, sc = O.predict_topk(ue, ie, g["test_users"], _train_lists(g), kmax)
    _, raw = O.predict_topk(ue, ie, g["test_users"], _train_lists(g), kmax, round_decimals=None)
    # the reference list is in torch.topk order: descending by the UNROUNDED score, so with no exact
    # ties among finite scores it is already canonical; the stored scores are rounded to 4 d.p. (G9)
    ref_ids, ref_sc = g["pred_ids"], g["pred_scores"]
    finite = np.isfinite(ref_sc)
    assert np.array_equal(np.isfinite(sc), finite)
    assert np.array_equal(sc[finite], ref_sc[finite])
    for r in range(len(ids)):
        f = finite[r]
        assert len(np.unique(raw[r][f])) == f.sum(), "fixture has exact score ties"
        assert np.array_equal(ids[r][f], ref_ids[r][f])            # bit-exact where scores are finite
    # -inf tail (G9): the reference's choice among masked items is arbitrary; ours is lowest id first
    tl = _train_lists(g)
    for r, u in enumerate(g["test_users"]):
        tail = ids[r][~finite[r]]
        assert np.array_equal(tail, np.sort(np.asarray(tl[u]))[:len(tail)])
    # metrics: identical to the reference's calculate_metrics on its own lists
    test_lists = O.train_lists_from_edges(g["test_u"], g["test_i"], int(g["n_users"]))
    y_true = [test_lists[u] for u in g["test_users"]]
    res = O.calculate_metrics(ids.tolist(), y_true, g["ks"].tolist())
    for m in ["recall", "precision", "hit", "ndcg", "f1"]:
        assert np.allclose(res[m], g["metric_" + m], rtol=0, atol=1e-12), m


@pytest.mark.parametrize("case", LGCN_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_loss_and_grads_match_reference(case, mode):
    g = load_golden(case)
    keep = torch.from_numpy(g[f"{mode}_keep"]) if mode == "train" else None
    out = O.train_step_loss_and_grads(_norm(g), torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"]),
                                      int(g["n_layers"]), torch.from_numpy(g["batch"]), float(g["reg_lambda"]),
                                      keep_mask=keep, dropout=float(g["dropout"]), single=bool(g["single"]))
    assert np.array_equal(out["loss"].numpy(), g[f"{mode}_loss"])
    assert np.array_equal(out["bpr"].numpy(), g[f"{mode}_bpr"])
    assert np.array_equal(out["reg"].numpy(), g[f"{mode}_reg"])
    assert np.array_equal(out["grad_user"].numpy(), g[f"{mode}_grad_embedding_user_weight"])
    assert np.array_equal(out["grad_item"].numpy(), g[f"{mode}_grad_embedding_item_weight"])


def test_adv_selection_matches_reference():
    g = load_golden("small_adv")
    ue, ie = O.propagate(_norm(g), torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"]), int(g["n_layers"]))
    data = torch.from_numpy(g["data"])
    users, cands = data[:, 0], data[:, 1:]
    rankings = O.adv_rank_candidates(ue, ie, users, cands)
    assert np.allclose(rankings.numpy(), g["rankings"], rtol=1e-6, atol=1e-7)
    negs = O.adv_select_negatives(rankings, cands, users.numpy(), _train_lists(g), int(max(g["ks"])))
    sampled = [row[row >= 0] for row in g["sampled_pos"]]
    triples = O.adv_build_triples(users.numpy(), sampled, negs)
    assert np.array_equal(triples, g["triples"])
    out = O.train_step_loss_and_grads(_norm(g), torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"]),
                                      int(g["n_layers"]), torch.from_numpy(triples), float(g["reg_lambda"]))
    assert np.allclose(out["loss"].numpy(), g["loss"], rtol=1e-6)
    assert np.allclose(out["grad_user"].numpy(), g["grad_user"], rtol=1e-5, atol=1e-9)


def _ltr_parts(g):
    tabs = {k: torch.from_numpy(g[k]) for k in ["users_rev", "users_desc", "items_rev", "items_desc"]}
    ws = [torch.from_numpy(g[f"head_w{i}"]) for i in range(int(g["n_head_layers"]))]
    bs = [torch.from_numpy(g[f"head_b{i}"]) for i in range(int(g["n_head_layers"]))]
    return tabs, ws, bs


@pytest.mark.parametrize("case,with_pop", [("small_ltr_linear", False), ("small_ltr_pop", True)])
def test_ltr_scores_match_reference(case, with_pop):
    g = load_golden(case)
    tabs, ws, bs = _ltr_parts(g)
    pop = (torch.from_numpy(g["pop_users"]), torch.from_numpy(g["pop_items"])) if with_pop else None
    ue, ie = O.propagate(_norm(g), torch.from_numpy(g["user_w"]), torch.from_numpy(g["item_w"]), int(g["n_layers"]))
    users = torch.arange(int(g["n_users"]))
    sb = O.ltr_score_batchwise(ue[users], ie, users, tabs, ws, bs, pop)
    assert np.allclose(sb.numpy(), g["score_batchwise"], rtol=1e-6, atol=1e-6)
    pu, pi = torch.from_numpy(g["pair_users"]), torch.from_numpy(g["pair_items"])
    sp = O.ltr_score_pairwise(ue[pu], ie[pi], pu, pi, tabs, ws, bs, pop)
    assert sp.shape == (len(pu), 1)                                 # G15
    assert np.allclose(sp.numpy(), g["score_pairwise"], rtol=1e-6, atol=1e-6)
    # collapsed affine head (G16) reproduces the stacked head
    w, b = O.collapse_linear_stack(ws, bs)
    f = O.ltr_features_pairwise(ue[pu], tabs["users_rev"][pu], tabs["users_desc"][pu], ie[pi],
                                tabs["items_rev"][pi], tabs["items_desc"][pi]).double()
    if with_pop:
        f = torch.cat([f, pop[0][pu].double(), pop[1][pi].double()], dim=-1)
    assert np.allclose((f @ w + b).numpy(), g["score_pairwise"][:, 0], rtol=1e-5, atol=1e-5)


def test_canonical_topk_and_tie_compare():
    s = np.array([[1, -np.inf, -np.inf, -np.inf, 1, 1]], dtype=np.float32)
    ids, sc = O.canonical_topk(s, 5)
    assert ids.tolist() == [[0, 4, 5, 1, 2]]                        # G10 example, canonical order
    rng = np.random.default_rng(0)
    sc = rng.integers(-3, 4, size=(50, 400)).astype(np.float32)     # heavy ties
    ids, out = O.canonical_topk(sc, 20)
    for r in range(50):
        order = np.lexsort((np.arange(400), -sc[r]))[:20]
        assert np.array_equal(ids[r], order)
    st = O.topk_lists_equivalent(ids, out, ids, out)
    assert st["exact"] == 50 and st["bad"] == 0


def test_metrics_restatement_edge_cases():
    res = O.calculate_metrics([[1, 2, 3], [9, 8, 7]], [[3, 5], [1]], [1, 3])
    assert res["recall"] == [0.0, 0.25] and res["hit"] == [0.0, 0.5]
    assert res["f1"][0] == 0.0                                      # 0/0 -> 0 (utils.py:55-62)


def test_synthetic_graph_generator():
    u, i = O.synthetic_interactions(500, 200, 3000, seed=0)
    assert len(u) == 3000 and len(np.unique(u * 200 + i)) == 3000
    assert len(np.unique(u)) == 500 and len(np.unique(i)) == 200    # every user and item has an edge
    assert np.all(np.diff(u * 200 + i) > 0)
