"""TEST / BASELINE INFRASTRUCTURE (never imported by the product): the reference's OWN ``BaseModel`` — unmodified, imported
from /root/reference or from the copy staged in baseline/_ref — constructed on ``device='cpu'`` from a stub dataset that
exposes exactly the attributes ``BaseModel._copy_dataset_params`` reads (base_model.py:54-62), for ``bench.py``'s
``cpu_baseline`` leg and ``--impl reference`` arm (BASELINE.md "CPU-baseline plan", steps 1-5).

What is timed is the reference's stock code path: ``model.representation`` (base_model.py:93-106: cat, L x torch.sparse.mm,
stack, mean, split) and ``model.predict`` (:235-276: one representation, then per batch of 2048 users matmul, the pandas
``explode`` mask, ``topk``, ``round``, ``.tolist()``).
"""
from __future__ import annotations

import logging
import os
import time
from types import SimpleNamespace

import numpy as np
import torch


def _shims():
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    import ref_shims
    return ref_shims


def reference_root():
    return _shims().reference_root()


def build_model(n_users: int, n_items: int, rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, user_w: torch.Tensor,
                item_w: torch.Tensor, n_layers: int, ks, batch_size: int = 2048):
    """The reference ``BaseModel`` on the CPU over the CSR (rowptr int64 (N+1), col (nnz) global column ids, val fp32) of a
    normalised adjacency Â with N = n_users + n_items rows; E0 = (user_w, item_w)."""
    import pandas as pd
    shims = _shims()
    shims.install()
    from TextGCN.base_model import BaseModel
    n = n_users + n_items
    counts = np.diff(rowptr)
    row = np.repeat(np.arange(n, dtype=np.int64), counts)
    norm = torch.sparse_coo_tensor(torch.from_numpy(np.stack([row, col.astype(np.int64)])), torch.from_numpy(val.astype(np.float32)),
                                   (n, n)).coalesce()           # dataset.py:151-157 leaves a coalesced COO with int64 indices
    ucol = col[:rowptr[n_users]].astype(np.int64) - n_users
    lists = np.split(ucol, rowptr[1:n_users])
    ds = SimpleNamespace(
        n_users=n_users, n_items=n_items, norm_matrix=norm,
        true_test_lil=[[0]], test_df=pd.DataFrame({"user_id": [0]}),
        train_user_dict=pd.Series([x.tolist() for x in lists], index=pd.RangeIndex(n_users, name="user_id")),   # dataset.py:108
        user_mapping=pd.DataFrame({"remap_id": [0], "org_id": ["u0"]}),
        item_mapping=pd.DataFrame({"remap_id": [0], "org_id": ["i0"]}),
    )
    params = SimpleNamespace(k=sorted(ks), lr=1e-3, uid="cpu_baseline", save=False, quiet=True, epochs=1, logger=logging.getLogger("reference_cpu"),
                             device=torch.device("cpu"), dropout=0.4, emb_size=int(user_w.shape[1]), n_layers=n_layers,
                             save_path="runs/cpu_baseline", batch_size=batch_size, reg_lambda=1e-4, evaluate_every=1, neg_samples=1,
                             slurm=True, single=False, load=None)
    model = BaseModel(params, ds)
    with torch.no_grad():
        model.embedding_user.weight.copy_(user_w)
        model.embedding_item.weight.copy_(item_w)
    model.training = False
    return model


def time_representation(model, steps: int, warmup: int, budget_s: float = 150.0):
    """Per-call seconds of ``model.representation`` (no dropout: ``training`` is False, as inside ``predict``)."""
    with torch.no_grad():
        for _ in range(warmup):
            model.representation
        times = []
        stop = time.perf_counter() + budget_s
        for _ in range(steps):
            t = time.perf_counter()
            model.representation
            times.append(time.perf_counter() - t)
            if time.perf_counter() > stop:
                break
    return times


def time_predict(model, users: np.ndarray):
    """Seconds of ``model.predict(users)`` split as (total, the representation call inside it)."""
    t = time.perf_counter()
    with torch.no_grad():
        model.representation
    t_rep = time.perf_counter() - t
    t = time.perf_counter()
    preds = model.predict(users)
    return time.perf_counter() - t, t_rep, preds
