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
    """``capturable=True`` keeps the step count and the bias corrections in device memory (one tiny kernel per step
    and parameter group), so ``step()`` issues no per-step host scalar and can be replayed from a CUDA graph; it then
    only accepts parameters the fused kernel handles (fp32, contiguous, size % 4 == 0)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = capturable

    def _step_capturable(self):
        for group in self.param_groups:
            b1, b2 = group["betas"]
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            if "dev_step" not in group:
                group["dev_step"] = torch.zeros(1, dtype=torch.int64, device=dev)
                group["dev_bc"] = torch.zeros(2, dtype=torch.float32, device=dev)
            ops.adam_prepare(group["dev_step"], group["dev_bc"], b1, b2)
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.numel() % 4 == 0 and p.is_contiguous()):
                    raise ValueError("FusedAdam(capturable=True) needs fp32 contiguous CUDA parameters with numel % 4 == 0")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ops.adam_step_dev(p, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], group["lr"], b1, b2, group["eps"],
                                  group["dev_bc"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self.capturable:
            self._step_capturable()
            return loss
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
