"""Construction of the normalised adjacency Â (reference: dataset.py:122-157) without dgl / scipy.

``norm_adj_csr`` reproduces the reference values bit for bit: the product (d[r]·a)·d[c] is evaluated in
float64 and cast to float32, with d = rowsum^-0.5 taken from ``np.power`` on the host for each distinct
degree (so the float64 rounding is numpy's, exactly as in the reference) and 0 for empty rows.  Everything
else is torch ops on whatever device the interaction arrays live on, so the 200M-edge graph is built on the
GPU in seconds.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .ops import Graph


def norm_adj_csr(train_u: torch.Tensor, train_i: torch.Tensor, n_users: int, n_items: int
                 ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(rowptr int32 (N+1), col int32 (nnz), val fp32 (nnz)) of Â = D^-1/2 (A + Aᵀ) D^-1/2, rows sorted by
    (row, col) — the order of the reference's coalesced COO (dataset.py:138)."""
    u = train_u.to(torch.int64)
    i = train_i.to(torch.int64)
    n = n_users + n_items
    r = torch.cat([u, i + n_users])
    c = torch.cat([i + n_users, u])
    key = r * n + c
    ukey, counts = torch.unique(key, sorted=True, return_counts=True)   # coalesce: duplicates add up
    row = torch.div(ukey, n, rounding_mode="floor")
    col = ukey - row * n
    a = counts.to(torch.float64)
    rowsum = torch.zeros(n, dtype=torch.float64, device=key.device).index_add_(0, row, a)
    # d = rowsum ** -0.5 through numpy on the distinct degree values (bit-exact with the reference's np.power)
    degs = torch.unique(rowsum)
    degs_np = degs.cpu().numpy()
    with np.errstate(divide="ignore"):
        dinv_np = np.power(degs_np, -0.5)
    dinv_np[np.isinf(dinv_np)] = 0.0
    table = torch.from_numpy(dinv_np).to(key.device)
    d_inv = table[torch.searchsorted(degs, rowsum)]
    val = ((d_inv[row] * a) * d_inv[col]).to(torch.float32)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rowptr.to(torch.int32), col.to(torch.int32), val


def graph_from_interactions(train_u, train_i, n_users: int, n_items: int, device) -> Graph:
    device = torch.device(device)
    tu = torch.as_tensor(np.asarray(train_u) if not isinstance(train_u, torch.Tensor) else train_u).to(device)
    ti = torch.as_tensor(np.asarray(train_i) if not isinstance(train_i, torch.Tensor) else train_i).to(device)
    rowptr, col, val = norm_adj_csr(tu, ti, n_users, n_items)
    return Graph(n_users, n_items, rowptr.contiguous(), col.contiguous(), val.contiguous())


def csr_to_norm_matrix(rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor) -> torch.Tensor:
    """The reference's ``dataset.norm_matrix`` view of the same data: coalesced COO, int64 indices."""
    n = rowptr.numel() - 1
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    row = torch.repeat_interleave(torch.arange(n, device=rowptr.device), counts)
    idx = torch.stack([row, col.to(torch.int64)])
    return torch.sparse_coo_tensor(idx, val, (n, n), is_coalesced=True)
