"""Fused dense Adam over the embedding tables (reference: base_model.py:111, :126).

Same update rule and defaults as ``torch.optim.Adam`` (no amsgrad / weight decay); each table is updated by
one float4-vectorised kernel (7·N·d·4 bytes of HBM traffic, the minimum).  Parameters whose size is not a
multiple of 4 (the 5- or 7-weight LTR head) are tiny and are updated with the same formulas in torch.
"""
from __future__ import annotations

import math

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad.contiguous()
                if p.is_cuda and p.dtype == torch.float32 and p.numel() % 4 == 0 and p.is_contiguous():
                    ops.adam_step(p, g, st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"], st["step"])
                else:
                    st["exp_avg"].lerp_(g, 1 - b1)
                    st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
                    bc1 = 1 - b1 ** st["step"]
                    bc2 = 1 - b2 ** st["step"]
                    denom = (st["exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(group["eps"])
                    p.addcdiv_(st["exp_avg"], denom, value=-group["lr"] / bc1)
        return loss
