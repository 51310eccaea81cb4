"""Synthetic Amazon-shaped interaction graphs built on the device (BASELINE.json configs).

Log-normal user / item popularity (σ_u = 1.0, σ_i = 1.3), (u, i) pairs drawn from the product distribution,
deduplicated, at least one edge per user and per item, exactly ``n_edges`` rows sorted by (u, i).  All torch ops
on the target device with a seeded generator, so every rank of a multi-GPU run builds the identical graph and the
200M-edge configuration takes seconds.
"""
from __future__ import annotations

from typing import Tuple

import torch

WORKLOADS = {
    # name: (n_users, n_items, n_edges, emb, layers)   — BASELINE.json configs[1] and configs[4]
    "c2": (190_000, 63_000, 1_700_000, 64, 3),
    "c5": (10_000_000, 2_000_000, 200_000_000, 128, 4),
    "tiny": (3_000, 1_000, 30_000, 64, 3),
}


def interactions(n_users: int, n_items: int, n_edges: int, device, seed: int = 0,
                 sigma_u: float = 1.0, sigma_i: float = 1.3) -> Tuple[torch.Tensor, torch.Tensor]:
    assert max(n_users, n_items) <= n_edges <= n_users * n_items
    gen = torch.Generator(device=device).manual_seed(seed)
    wu = torch.exp(torch.randn(n_users, generator=gen, device=device, dtype=torch.float64) * sigma_u)
    wi = torch.exp(torch.randn(n_items, generator=gen, device=device, dtype=torch.float64) * sigma_i)
    cu = torch.cumsum(wu / wu.sum(), 0)
    ci = torch.cumsum(wi / wi.sum(), 0)

    def draw(cdf, m, hi):
        r = torch.rand(m, generator=gen, device=device, dtype=torch.float64)
        return torch.searchsorted(cdf, r).clamp_(max=hi - 1)

    fu = torch.arange(n_users, device=device)
    gi = torch.arange(n_items, device=device)
    forced = torch.unique(torch.cat([fu * n_items + draw(ci, n_users, n_items), draw(cu, n_items, n_users) * n_items + gi]))
    keys = forced
    while keys.numel() < n_edges:
        m = int((n_edges - keys.numel()) * 1.25) + 1024
        chunk = 50_000_000
        parts = [keys]
        for s in range(0, m, chunk):
            c = min(chunk, m - s)
            parts.append(draw(cu, c, n_users) * n_items + draw(ci, c, n_items))
        keys = torch.unique(torch.cat(parts))
    if keys.numel() > n_edges:
        is_forced = torch.isin(keys, forced, assume_unique=True)
        extra = keys[~is_forced]
        n_keep = n_edges - forced.numel()
        perm = torch.randperm(extra.numel(), generator=gen, device=device)[:n_keep]
        keys = torch.sort(torch.cat([forced, extra[perm]])).values
    u = torch.div(keys, n_items, rounding_mode="floor")
    return u, keys - u * n_items
