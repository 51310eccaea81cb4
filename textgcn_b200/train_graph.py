"""One CUDA graph per training step: zero_grad -> device dropout draw -> propagate -> fused BPR(SELU)+L2 ->
Horner backward -> fused Adam, captured once and replayed.  The reference's step (base_model.py:118-126) launches the same
work from Python every iteration; at the Electronics-shaped config the kernels take ~0.8 ms and the host ~0.4 ms.

What makes the step capturable: the dropout draw counter and Adam's step count / bias corrections live in device
memory (``tgcn_dropout_mask_dev``, ``tgcn_adam_prepare`` + ``tgcn_adam_step_dev``), the batch is copied into a static
buffer, and the loss accumulators are updated in place.  Batches of another shape run eagerly through the same
optimizer state, so a ragged last batch is fine.

The step calls the kernels directly instead of going through autograd: loss = bpr + reg, so both upstream gradients are
exactly 1 and ``backward`` is ``tgcn_propagate_bwd`` accumulating into the regulariser's gradient buffer, whose two halves
ARE the gradients of the two embedding tables.  Compared with the autograd route (``_FusedBprFn``) that removes two
(N, d) scaling passes, one zero fill and the accumulate-into-``.grad`` pass per step (≈ 0.1 ms of 1.0 ms at c2) and all
per-step allocations.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import TgcnError
from .optim import FusedAdam


class GraphedTrainStep:
    def __init__(self, model, optimizer: FusedAdam, warmup: int = 3):
        if not isinstance(optimizer, FusedAdam) or not optimizer.capturable:
            raise TgcnError("GraphedTrainStep needs FusedAdam(capturable=True)")
        if model.dropout > 0 and model.dropout_rng != "device":
            raise TgcnError('GraphedTrainStep needs dropout_rng="device" (the host generator cannot be captured)')
        self.model, self.opt, self.warmup = model, optimizer, warmup
        self.device = model.graph.device
        if model.dropout > 0 and model.__dict__.get("_b200_dev_draws") is None:
            # continue the model's draw sequence from wherever the eager steps left it
            model.__dict__["_b200_dev_draws"] = torch.full((1,), int(model.__dict__.get("_b200_mask_draws", 0)),
                                                           dtype=torch.int64, device=self.device)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.shape = None
        self.static_batch: Optional[torch.Tensor] = None
        self.static_loss: Optional[torch.Tensor] = None
        self.loss_sums = torch.zeros(2, dtype=torch.float32, device=self.device)  # [bpr, reg] summed over the steps taken
        self.calls = 0
        params = [p for p in model.parameters() if p.requires_grad]
        uw, iw = model.embedding_user.weight, model.embedding_item.weight
        if len(params) != 2 or not (params[0] is uw and params[1] is iw):
            raise TgcnError("GraphedTrainStep trains exactly the two embedding tables (plain LightGCN BPR steps)")
        n, d = model.n_users + model.n_items, uw.shape[1]
        self.emb = torch.empty((n, d), dtype=torch.float32, device=self.device)
        self.grad_emb = torch.empty_like(self.emb)          # dL/d(representation), scattered by the fused BPR kernel
        self.grad_w0 = torch.empty_like(self.emb)           # dL/dE0: regulariser part + Horner backward accumulated into it
        uw.grad, iw.grad = self.grad_w0[:model.n_users], self.grad_w0[model.n_users:]   # views: the optimizer reads them in place

    def _step(self, data: torch.Tensor) -> torch.Tensor:
        m = self.model
        g, L, single, p = m.graph, m.n_layers, m._single(), float(m.dropout)
        users, pos, negs = m._split_batch(data)
        uw, iw = m.embedding_user.weight.detach(), m.embedding_item.weight.detach()
        keep = m._draw_keep_mask()
        ops.propagate_fwd(g, uw, iw, L, single, keep, p, out=self.emb)
        self.grad_emb.zero_()
        self.grad_w0.zero_()
        losses = ops.bpr_fwd_bwd(g.n_users, g.n_items, self.emb, uw, iw, users, pos, negs, float(m.reg_lambda), self.grad_emb, self.grad_w0)
        ops.propagate_bwd(g, self.grad_emb, L, single, keep, p, grad_in=self.grad_w0, accumulate=True)
        if m.embedding_user.weight.grad is None or m.embedding_user.weight.grad.data_ptr() != self.grad_w0.data_ptr():
            m.embedding_user.weight.grad, m.embedding_item.weight.grad = self.grad_w0[:g.n_users], self.grad_w0[g.n_users:]
        self.opt.step()
        self.loss_sums.add_(losses)
        return losses[0] + losses[1]

    def __call__(self, data: torch.Tensor) -> torch.Tensor:
        """One optimisation step on ``data`` ((B, 2 + n_neg) int64).  Returns the step's loss (a device scalar that the next
        call overwrites)."""
        data = data.to(self.device)
        self.calls += 1
        if self.graph is not None and tuple(data.shape) == self.shape:
            self.static_batch.copy_(data)
            self.graph.replay()
            return self.static_loss
        if self.graph is None and self.calls > self.warmup and (self.shape is None or tuple(data.shape) == self.shape):
            # capture on the side stream torch asks for; nothing runs during capture, so replay once for this batch
            self.shape = tuple(data.shape)
            self.static_batch = data.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.device(self.device), torch.cuda.graph(g):
                self.static_loss = self._step(self.static_batch)
            self.graph = g
            self.model.graph.pin_workspace()  # the captured launches hold its address
            g.replay()
            return self.static_loss
        # warm-up steps (real steps: parameter gradients and optimizer state must exist before capture) and odd shapes
        if self.shape is None and self.calls >= self.warmup:
            self.shape = tuple(data.shape)
        return self._step(data)
