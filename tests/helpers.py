"""Shared test helpers (CPU-safe imports only at module level)."""
from __future__ import annotations

import logging
from types import SimpleNamespace

import numpy as np
import torch

from oracle import lightgcn_oracle as O

TOL = 1e-5  # north_star: propagated embeddings and scores within 1e-5 relative (norm-wise)


def rel_err(a, b) -> float:
    """max |a - b| / max |b|  (norm-wise relative error)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / denom)


def golden_norm(g) -> torch.Tensor:
    n = int(g["n_users"] + g["n_items"])
    return O.sparse_tensor(g["norm_row"], g["norm_col"], g["norm_val"], n)


def golden_lists(g, which="train"):
    return O.train_lists_from_edges(g[f"{which}_u"], g[f"{which}_i"], int(g["n_users"]))


class StubDataset:
    """The attributes the model classes copy (base_model.py:54-62, advanced_sampling.py:32-35,
    ltr_models.py:49-55, :216-219), filled from a golden fixture."""

    pos_samples = 5

    def __init__(self, g, device):
        self.n_users, self.n_items = int(g["n_users"]), int(g["n_items"])
        self.norm_matrix = golden_norm(g).to(device)
        tl = golden_lists(g, "train")
        test = golden_lists(g, "test")
        self.test_users = np.asarray(g["test_users"]) if "test_users" in g else np.unique(g["test_u"])
        self.true_test_lil = [test[u].tolist() for u in self.test_users]
        self.positive_lists = [{"list": t.tolist(), "set": set(t.tolist())} for t in tl]
        for src, dst in (("items_rev", "items_as_avg_reviews"), ("users_rev", "users_as_avg_reviews"),
                         ("users_desc", "users_as_avg_desc"), ("items_desc", "items_as_desc"),
                         ("pop_users", "popularity_users"), ("pop_items", "popularity_items")):
            if src in g:
                setattr(self, dst, torch.from_numpy(g[src]).to(device))
        self.all_items = range(self.n_items)


def params_from_golden(g, **kw):
    from textgcn_b200.models import make_params
    base = dict(k=[int(x) for x in g["ks"]], emb_size=int(g["user_w"].shape[1]), n_layers=int(g["n_layers"]),
                dropout=float(g["dropout"]), reg_lambda=float(g["reg_lambda"]), single=bool(g["single"]) if "single" in g else False,
                logger=logging.getLogger("test"))
    base.update(kw)
    return make_params(**base)


def load_weights(model, g):
    with torch.no_grad():
        model.embedding_user.weight.copy_(torch.from_numpy(g["user_w"]))
        model.embedding_item.weight.copy_(torch.from_numpy(g["item_w"]))
        if hasattr(model, "layers"):
            for li, layer in enumerate(model.layers):
                layer.weight.copy_(torch.from_numpy(g[f"head_w{li}"]))
                layer.bias.copy_(torch.from_numpy(g[f"head_b{li}"]))
