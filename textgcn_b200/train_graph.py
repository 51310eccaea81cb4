"""One CUDA graph per training step: zero_grad -> device dropout draw -> propagate -> fused BPR(SELU)+L2 ->
Horner backward -> fused Adam, captured once and replayed.  The reference's step (base_model.py:118-126) launches the same
work from Python every iteration; at the Electronics-shaped config the kernels take ~0.8 ms and the host ~0.4 ms.

What makes the step capturable: the dropout draw counter and Adam's step count / bias corrections live in device
memory (``tgcn_dropout_mask_dev``, ``tgcn_adam_prepare`` + ``tgcn_adam_step_dev``), the batch is copied into a static
buffer, and the loss accumulators are updated in place.  Batches of another shape run eagerly through the same
optimizer state, so a ragged last batch is fine.
"""
from __future__ import annotations

from typing import Optional

import torch

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

    def _step(self, data: torch.Tensor) -> torch.Tensor:
        m = self.model
        self.opt.zero_grad(set_to_none=False)
        users, pos, negs = m._split_batch(data)
        losses = m._fused_losses(users, pos, negs, m.reg_lambda)
        self.loss_sums.add_(losses.detach())
        loss = losses[0] + losses[1]
        loss.backward()
        self.opt.step()
        return loss.detach()

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
