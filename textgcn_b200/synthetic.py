"""Synthetic Amazon-shaped interaction graphs built on the device (BASELINE.json configs).

Log-normal user / item popularity (σ_u = 1.0, σ_i = 1.3), (u, i) pairs drawn from the product distribution,
deduplicated, at least one edge per user and per item, exactly ``n_edges`` rows sorted by (u, i).  All torch ops
on the target device with a seeded generator, so the 200M-edge configuration takes seconds.

Every rank of a multi-GPU run builds the graph itself and all ranks must end up with the IDENTICAL graph, so every step is
bit-reproducible: in particular the popularity CDFs are INTEGER prefix sums (fixed-point weights).  Round 1 drew from a
float64 ``torch.cumsum`` on the device, which is not run-to-run deterministic in its last bits; a handful of boundary
draws then landed on a neighbouring user / item, ranks disagreed on a few edges of Â, and the self-verifying bench of
round 2 caught it as a parity failure of the multi-GPU result (profiles/r02/README.md).
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

    def int_cdf(n, sigma):
        """Log-normal weights in 32-bit fixed point and their exact (int64, order-independent) prefix sums."""
        w = torch.exp(torch.randn(n, generator=gen, device=device, dtype=torch.float64) * sigma)
        q = (w / w.max() * float(1 << 32)).to(torch.int64).clamp_(min=1)
        return torch.cumsum(q, 0)

    cu = int_cdf(n_users, sigma_u)
    ci = int_cdf(n_items, sigma_i)

    def draw(cdf, m, hi):
        r = torch.randint(0, int(cdf[-1]), (m,), generator=gen, device=device, dtype=torch.int64)
        return torch.searchsorted(cdf, r, right=True).clamp_(max=hi - 1)

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
        # a random subset by seeded integer priorities + a stable sort (bit-reproducible, unlike relying on randperm's kernel)
        prio = torch.randint(0, 1 << 62, (extra.numel(),), generator=gen, device=device, dtype=torch.int64)
        perm = torch.argsort(prio, stable=True)[:n_keep]
        keys = torch.sort(torch.cat([forced, extra[perm]])).values
    u = torch.div(keys, n_items, rounding_mode="floor")
    return u, keys - u * n_items
