"""Host-side mirror of the reference's model API over the sm_100a kernels.

The reference's plugin surface is Python method override on ``BaseModel``.  ``B200HotPath``
overrides exactly the hot-path methods — ``representation``, ``layer_aggregation``, ``get_loss`` / ``bpr_loss``,
``predict``, ``evaluate`` — keeping names, argument meaning and return types, so it can be mixed in front of the
reference's own classes or used with the standalone classes below, which restate the thin
non-hot-path shell (constructor contract, ``fit`` loop, checkpointing) so the package runs where the reference
is not installed.  State-dict keys are the reference's (``embedding_user.weight``, ``embedding_item.weight``,
``layers.N.weight/bias``), so checkpoints interchange.

All compute goes through ``textgcn_b200.ops`` -> libtgcn_b200.so.  There is no CPU path: constructing a model on
a non-CUDA device raises.
"""
from __future__ import annotations

import os
import random
import shutil
from collections import defaultdict
from types import SimpleNamespace
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import metrics as M
from . import ops
from ._lib import TgcnError


def make_params(**kw) -> SimpleNamespace:
    """Namespace with the reference's defaults (parser.py:11-161) for programmatic construction."""
    import logging
    p = dict(k=[20, 40], lr=1e-3, uid="b200", save=False, quiet=True, epochs=1000, logger=logging.getLogger("textgcn_b200"),
             device=torch.device("cuda"), dropout=0.4, emb_size=64, n_layers=3, save_path="runs/b200", batch_size=2048,
             reg_lambda=1e-4, evaluate_every=25, neg_samples=1, slurm=True, single=False, load=None, load_base=None,
             freeze=False, ltr_layers=[], seed=0)
    p.update(kw)
    p["k"] = sorted(p["k"])
    return SimpleNamespace(**p)


# ------------------------------------------------------------------------------------------------
# autograd glue
# ------------------------------------------------------------------------------------------------
class _PropagateFn(torch.autograd.Function):
    """representation as one differentiable op: forward = K fused SpMM passes, backward = K transposed passes
    in Horner form (no saved activations: the op is linear in E0)."""

    @staticmethod
    def forward(ctx, user_w, item_w, graph, n_layers, single, keep, dropout):
        ctx.graph, ctx.n_layers, ctx.single, ctx.keep, ctx.dropout = graph, n_layers, single, keep, dropout
        return ops.propagate_fwd(graph, user_w.detach().contiguous(), item_w.detach().contiguous(), n_layers, single, keep, dropout)

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.graph
        grad_in = ops.propagate_bwd(g, grad_out.contiguous(), ctx.n_layers, ctx.single, ctx.keep, ctx.dropout)
        return grad_in[:g.n_users], grad_in[g.n_users:], None, None, None, None, None


class _SpmmFn(torch.autograd.Function):
    """layer_aggregation (operator-level plugin point): Y = Â·X; Â is symmetric so backward is the same op."""

    @staticmethod
    def forward(ctx, x, graph):
        ctx.graph = graph
        return ops.spmm(graph, x.detach().contiguous())

    @staticmethod
    def backward(ctx, grad_out):
        return ops.spmm(ctx.graph, grad_out.contiguous()), None


class _FusedBprFn(torch.autograd.Function):
    """get_loss as one op: propagate -> fused gather/score/SELU/L2 + gradient scatter; backward = Horner passes.

    Returns losses = [bpr, reg].  The gradient w.r.t. the propagated embeddings is produced by the same kernel
    that computes the loss, so backward only has to push it through Âᵀ.
    """

    @staticmethod
    def forward(ctx, user_w, item_w, graph, n_layers, single, keep, dropout, users, pos, negs, reg_lambda):
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        uw, iw = user_w.detach().contiguous(), item_w.detach().contiguous()
        emb = ops.propagate_fwd(graph, uw, iw, n_layers, single, keep, dropout)
        grad_emb = torch.zeros_like(emb) if need_grad else None
        grad_w0 = torch.zeros_like(emb) if need_grad else None
        losses = ops.bpr_fwd_bwd(graph.n_users, graph.n_items, emb, uw, iw, users, pos, negs, reg_lambda, grad_emb, grad_w0)
        ctx.graph, ctx.n_layers, ctx.single, ctx.keep, ctx.dropout = graph, n_layers, single, keep, dropout
        ctx.grad_emb, ctx.grad_w0 = grad_emb, grad_w0
        return losses

    @staticmethod
    def backward(ctx, g_losses):
        g = ctx.graph
        grad_emb, grad_w0 = ctx.grad_emb, ctx.grad_w0
        ctx.grad_emb = ctx.grad_w0 = None
        # scale by the upstream gradients ON THE DEVICE (0-dim operands, no host sync): two O(N·d) elementwise passes are
        # far cheaper than draining the stream to read two floats
        grad_emb.mul_(g_losses[0])
        grad_w0.mul_(g_losses[1])
        ops.propagate_bwd(g, grad_emb, ctx.n_layers, ctx.single, ctx.keep, ctx.dropout, grad_in=grad_w0, accumulate=True)
        return (grad_w0[:g.n_users], grad_w0[g.n_users:]) + (None,) * 9


class _LtrPairFeaturesFn(torch.autograd.Function):
    """(B, 5|7) LTR pairwise features; only feature 0 (emb·emb) depends on trainable embeddings."""

    @staticmethod
    def forward(ctx, emb, n_users, users, items, tabs, pop):
        ctx.n_users, ctx.users, ctx.items = n_users, users, items
        ctx.save_for_backward(emb)
        return ops.ltr_pairwise_features(n_users, emb.detach().contiguous(), users, items, tabs["users_rev"], tabs["users_desc"],
                                         tabs["items_rev"], tabs["items_desc"], *(pop if pop is not None else (None, None)))

    @staticmethod
    def backward(ctx, grad_f):
        (emb,) = ctx.saved_tensors
        grad_emb = None
        if ctx.needs_input_grad[0]:
            grad_emb = torch.zeros_like(emb)
            ops.ltr_pairwise_emb_bwd(ctx.n_users, emb.contiguous(), ctx.users, ctx.items, grad_f[:, 0].contiguous(), grad_emb)
        return grad_emb, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# the hot-path overrides
# ------------------------------------------------------------------------------------------------
class B200HotPath:
    """Mixin holding the kernel-backed overrides of BaseModel's hot-path methods."""

    dropout_rng = "host"  # "host": torch.rand on the CPU generator like base_model.py:82; "device": CUDA generator
    eval_precision = "auto"  # "fp32" = exact FMA kernel; "3xtf32" = tcgen05 tensor cores; "auto" = 3xTF32 when eligible

    # -- graph ---------------------------------------------------------------------------------
    @property
    def graph(self) -> ops.Graph:
        g = self.__dict__.get("_b200_graph")
        if g is None:
            dev = torch.device(self.device)
            if dev.type != "cuda":
                raise TgcnError("textgcn_b200 models need a CUDA device: there is no CPU fallback")
            g = ops.Graph.from_norm_matrix(self.norm_matrix, self.n_users, self.n_items, device=dev)
            self.__dict__["_b200_graph"] = g
        return g

    def _draw_keep_mask(self) -> Optional[torch.Tensor]:
        """Bernoulli(1-p) keep mask over nnz(Â) (base_model.py:82); None in eval mode."""
        if not self.training or self.dropout <= 0:
            return None
        nnz = self.graph.nnz
        if self.dropout_rng == "device":  # one kernel, counter-based hash keyed by torch's seed and a per-model draw counter
            base = (torch.initial_seed() * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
            draws = self.__dict__.get("_b200_dev_draws")
            if draws is not None:  # CUDA-graph mode (train_graph.GraphedTrainStep): the draw counter lives on the device
                return ops.dropout_mask_dev(nnz, float(self.dropout), base, draws)
            n = self.__dict__["_b200_mask_draws"] = self.__dict__.get("_b200_mask_draws", 0) + 1
            seed = (base + n * 0xD6E8FEB86659FD93) & (2 ** 64 - 1)
            return ops.dropout_mask(nnz, float(self.dropout), seed, self.graph.device)
        return (torch.rand(nnz) < (1 - self.dropout)).to(self.graph.device)

    def _single(self) -> bool:
        return getattr(self.layer_combination, "__name__", "") == "layer_combination_single"

    # -- a2..a6 --------------------------------------------------------------------------------
    @property
    def representation(self):
        """base_model.py:93-106 as one fused call: dropout mask, K SpMM passes, layer mean, split."""
        out = _PropagateFn.apply(self.embedding_user.weight, self.embedding_item.weight, self.graph, self.n_layers,
                                 self._single(), self._draw_keep_mask(), float(self.dropout))
        return torch.split(out, [self.n_users, self.n_items])

    def layer_aggregation(self, norm_matrix, emb_matrix):
        """base_model.py:141-148.  ``norm_matrix`` must be the model's own Â (the kernel reads the CSR handle)."""
        if norm_matrix is not self.norm_matrix:
            raise TgcnError("layer_aggregation only accepts the model's norm_matrix; edge dropout is applied as a "
                            "keep-mask inside `representation`, not by rebuilding the matrix")
        return _SpmmFn.apply(emb_matrix, self.graph)

    # -- a7..a10 -------------------------------------------------------------------------------
    def score_pairwise(self, users_emb, items_emb, users, items):
        """base_model.py:166-171 (tiny (B, d) elementwise work: left to torch)."""
        return torch.sum(users_emb * items_emb, dim=1)

    def score_batchwise(self, users_emb, items_emb, users):
        raise TgcnError("score_batchwise is fused into predict(): the dense (B, n_items) score matrix is never materialised")

    def _split_batch(self, data):
        dev = self.graph.device
        data = data.to(dev)
        users, pos = data[:, 0], data[:, 1]
        negs = data[:, 2:].t()
        return ops.as_index(users, dev), ops.as_index(pos, dev), ops.as_index(negs, dev)

    def _fused_losses(self, users, pos, negs, reg_lambda):
        return _FusedBprFn.apply(self.embedding_user.weight, self.embedding_item.weight, self.graph, self.n_layers,
                                 self._single(), self._draw_keep_mask(), float(self.dropout), users, pos, negs, float(reg_lambda))

    def get_loss(self, data):
        """base_model.py:181-184: (B, 2 + n_neg) int64 batch -> bpr + reg, both from one fused kernel."""
        users, pos, negs = self._split_batch(data)
        losses = self._fused_losses(users, pos, negs, self.reg_lambda)
        self._loss_values["bpr"] += losses[0].detach()
        self._loss_values["reg"] += losses[1].detach()
        return losses[0] + losses[1]

    def bpr_loss(self, users, pos, negs):
        """base_model.py:186-198 (SELU, mean over negatives)."""
        dev = self.graph.device
        negs_t = torch.stack([torch.as_tensor(n) for n in negs]) if not isinstance(negs, torch.Tensor) else negs
        losses = self._fused_losses(ops.as_index(users, dev), ops.as_index(pos, dev), ops.as_index(negs_t, dev), 0.0)
        self._loss_values["bpr"] += losses[0].detach()
        return losses[0]

    def reg_loss(self, users, pos, negs):
        """base_model.py:200-210 on layer-0 rows (B·(2 + n_neg)·d floats: left to torch)."""
        negs_t = torch.stack(list(negs)) if not isinstance(negs, torch.Tensor) else negs
        loss = (self.embedding_user(users).norm(2).pow(2) + self.embedding_item(pos).norm(2).pow(2)
                + self.embedding_item(negs_t).norm(2).pow(2).mean())
        res = self.reg_lambda * loss / len(users) / 2
        self._loss_values["reg"] += res.detach()
        return res

    # -- a11..a13 ------------------------------------------------------------------------------
    def _rank(self, emb: torch.Tensor, users: torch.Tensor, k: int):
        """Fused score + mask + top-k for LightGCN scoring: user/item vectors are rows of the (N, d) table."""
        nu = self.n_users
        return ops.eval_topk(self.graph, emb[:nu], emb[nu:], k, users=users, precision=self.eval_precision)

    @torch.no_grad()
    def predict_device(self, users, k: Optional[int] = None):
        """Device-resident form of predict: (ids (n, k) int32, scores (n, k) fp32, rounded to 4 d.p.)."""
        self.training = False
        k = max(self.k) if k is None else k
        emb = ops.propagate_fwd(self.graph, self.embedding_user.weight.detach().contiguous(),
                                self.embedding_item.weight.detach().contiguous(), self.n_layers, self._single())
        users_t = ops.as_index(np.asarray(users) if not isinstance(users, torch.Tensor) else users, self.graph.device)
        ids, scores = self._rank(emb, users_t, k)
        return ids, scores.round(decimals=4)

    @torch.no_grad()
    def predict(self, users, save: bool = False, with_scores: bool = False):
        """base_model.py:235-276: list of lists of top max(k) item ids per user (canonical order: score desc,
        item id asc; train items excluded; short lists completed with train items at -inf)."""
        ids, scores = self.predict_device(users)
        predictions = ids.tolist()
        scores = scores.tolist()
        if save:
            import pandas as pd
            predictions_unmapped = [[self.item_mapping_dict[i] for i in row] for row in predictions]
            users_unmapped = [self.user_mapping_dict[u] for u in users]
            pred_df = pd.DataFrame({"user_id": users_unmapped, "y_pred": predictions_unmapped, "scores": scores})
            pred_df.to_csv(os.path.join(self.save_path, "predictions.tsv"), sep="\t", index=False)
            self.logger.info(f"Predictions are saved in `{os.path.join(self.save_path, 'predictions.tsv')}`")
        if with_scores:
            return predictions, scores
        return predictions

    @torch.no_grad()
    def evaluate(self, epoch=None):
        """base_model.py:212-233 with the metrics computed from the device table (utils.py:36-63 semantics)."""
        self.eval()
        self.training = False
        ids, _ = self.predict_device(self.test_users)
        results = M.calculate_metrics(ids, self.true_test_lil, self.k)
        self.logger.info(" " * 11 + "".join([f"@{i:<6}" for i in self.k]))
        for i in results:
            self.metrics_logger[i] = np.append(self.metrics_logger[i], [results[i]], axis=0)
            self.logger.info(f"{i:11}" + " ".join([f"{j:.4f}" for j in results[i]]))
        return results


# ------------------------------------------------------------------------------------------------
# standalone shell (constructor contract, fit loop, checkpointing: base_model.py:23-75, :108-139, :278-299)
# ------------------------------------------------------------------------------------------------
def early_stop(res) -> bool:
    """utils.py:79-90."""
    if len(res["recall"]) < 3:
        return False
    declining = all(np.less(m[-1], m[-2]).all() and np.less(m[-2], m[-3]).all() for m in res.values())
    converged = all(np.allclose(m[-1], m[-2], atol=1e-4) for m in res.values()) and \
        all(np.allclose(m[-1], m[-3], atol=1e-4) for m in res.values())
    return converged or declining


class BaseModel(B200HotPath, nn.Module):
    """LightGCN with BPR; same constructor contract as the reference: ``Model(params, dataset)``."""

    def __init__(self, params, dataset):
        super().__init__()
        self._copy_params(params)
        self._copy_dataset_params(dataset)
        self._init_embeddings(params.emb_size)
        self._add_vars(params)
        self.load_model(getattr(params, "load", None))
        self.to(params.device)

    def _copy_params(self, params):
        for name in ["k", "lr", "uid", "save", "quiet", "epochs", "logger", "device", "dropout", "emb_size", "n_layers",
                     "save_path", "batch_size", "reg_lambda", "evaluate_every", "neg_samples"]:
            setattr(self, name, getattr(params, name))
        self.device = torch.device(self.device)
        if self.device.type != "cuda":
            raise TgcnError("textgcn_b200 models need a CUDA device: there is no CPU fallback")
        self.slurm = params.slurm or params.quiet
        if getattr(params, "single", False):
            self.layer_combination = self.layer_combination_single
        self.fused_adam = getattr(params, "fused_adam", False)
        self.dropout_rng = getattr(params, "dropout_rng", "host")
        self.eval_precision = getattr(params, "eval_precision", "auto")
        self.nan_check = getattr(params, "nan_check", "step")
        self.cuda_graph = getattr(params, "cuda_graph", False)  # replay the training step from one CUDA graph (train_graph.py)

    def _copy_dataset_params(self, dataset):
        self.n_users = dataset.n_users
        self.n_items = dataset.n_items
        self.norm_matrix = dataset.norm_matrix
        self.true_test_lil = dataset.true_test_lil
        self.train_user_dict = getattr(dataset, "train_user_dict", None)
        if hasattr(dataset, "test_users"):
            self.test_users = np.asarray(dataset.test_users)
        else:
            self.test_users = np.sort(dataset.test_df.user_id.unique())
        if hasattr(dataset, "user_mapping"):
            self.user_mapping_dict = dict(dataset.user_mapping[["remap_id", "org_id"]].values)
            self.item_mapping_dict = dict(dataset.item_mapping[["remap_id", "org_id"]].values)
        if getattr(dataset, "graph", None) is not None:
            self.__dict__["_b200_graph"] = dataset.graph

    def _init_embeddings(self, emb_size):
        self.embedding_user = nn.Embedding(num_embeddings=self.n_users, embedding_dim=emb_size).to(self.device)
        self.embedding_item = nn.Embedding(num_embeddings=self.n_items, embedding_dim=emb_size).to(self.device)
        nn.init.normal_(self.embedding_user.weight, std=0.1)
        nn.init.normal_(self.embedding_item.weight, std=0.1)

    def _add_vars(self, params):
        self.metrics = list(M.METRICS)
        self.metrics_logger = {i: np.zeros((0, len(self.k))) for i in self.metrics}
        self.training = False
        self._loss_values = defaultdict(float)

    def layer_combination(self, vectors):
        """base_model.py:150-157 (kept for API parity; `representation` fuses it into the last SpMM pass)."""
        return torch.mean(torch.stack(vectors), axis=0)

    def layer_combination_single(self, vectors):
        return vectors[-1]

    @property
    def embedding_matrix(self):
        return torch.cat([self.embedding_user.weight, self.embedding_item.weight])

    def fit(self, batches):
        """base_model.py:108-139."""
        graphed = None
        if self.cuda_graph and type(self).get_loss is B200HotPath.get_loss:  # fixed-shape BPR steps only (not AdvSampl / LTR)
            from .optim import FusedAdam
            from .train_graph import GraphedTrainStep
            self.optimizer = FusedAdam(self.parameters(), lr=self.lr, capturable=True)
            graphed = GraphedTrainStep(self, self.optimizer)
        elif self.fused_adam:
            from .optim import FusedAdam
            self.optimizer = FusedAdam(self.parameters(), lr=self.lr)
        else:
            self.optimizer = torch.optim.Adam(self.parameters(), lr=self.lr)
        for epoch in range(1, self.epochs + 1):
            self.train()
            self.training = True
            self._loss_values = defaultdict(float)
            epoch_loss = 0
            nan_seen = torch.zeros((), dtype=torch.bool, device=self.device)
            for data in batches:
                if graphed is not None:
                    batch_loss = graphed(data)
                    nan_seen |= batch_loss.isnan()
                    epoch_loss += batch_loss
                    continue
                self.optimizer.zero_grad()
                batch_loss = self.get_loss(data)
                if self.nan_check == "step":  # the reference's per-step device sync (base_model.py:123)
                    assert not batch_loss.isnan(), f"loss is NA at epoch {epoch}"
                else:                          # same check without draining the stream every step
                    nan_seen |= batch_loss.detach().isnan()
                epoch_loss += batch_loss.detach()
                batch_loss.backward()
                self.optimizer.step()
            assert not bool(nan_seen), f"loss is NA at epoch {epoch}"
            if graphed is not None:  # the graph accumulates [bpr, reg] in place; hand the epoch's sums to the logger
                sums = graphed.loss_sums.clone()
                graphed.loss_sums.zero_()
                self._loss_values["bpr"], self._loss_values["reg"] = sums[0], sums[1]
            if epoch % self.evaluate_every:
                continue
            self.logger.info(f"Epoch {epoch}: {' '.join([f'{k} = {float(v):.4f}' for k, v in self._loss_values.items()])}")
            self.evaluate(epoch)
            self.checkpoint(epoch)
            if early_stop(self.metrics_logger):
                self.logger.warning(f"Early stopping triggerred at epoch {epoch}")
                break
        else:
            self.checkpoint(self.epochs)

    def load_model(self, load_path):
        """base_model.py:278-289."""
        if load_path is None:
            return
        if os.path.isdir(load_path):
            load_path = os.path.join(load_path, "best.pkl")
        self.logger.info(f"Loading model {load_path}")
        self.load_state_dict(torch.load(load_path, map_location=self.device))
        self.logger.info("Performance of the loaded model:")
        self.evaluate()
        self.metrics_logger = {i: np.zeros((0, len(self.k))) for i in self.metrics}

    def checkpoint(self, epoch):
        """base_model.py:291-299."""
        if not self.save:
            return
        os.makedirs(self.save_path, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(self.save_path, "latest_checkpoint.pkl"))
        if self.metrics_logger[self.metrics[0]][:, 0].max() == self.metrics_logger[self.metrics[0]][-1][0]:
            self.logger.info(f"Updating best model at epoch {epoch}")
            shutil.copyfile(os.path.join(self.save_path, "latest_checkpoint.pkl"), os.path.join(self.save_path, "best.pkl"))


# ------------------------------------------------------------------------------------------------
# (d) dynamic negative sampling  (advanced_sampling.py:25-69)
# ------------------------------------------------------------------------------------------------
class B200AdvSampl:
    """Overrides of AdvSamplModel: ranking + positive removal + top max(k) as one kernel per batch."""

    positive_sampler = "host"  # "host": Python random.sample like the reference; "device": keyed permutation kernel

    def score_pairwise_adv(self, users_emb, items_emb):
        raise TgcnError("score_pairwise_adv is fused into get_loss(): the (B, 1000, d) gather is never materialised")

    def _sample_positives_device(self, users: torch.Tensor) -> torch.Tensor:
        """Device counterpart of ``_sample_positives`` (advanced_sampling.py:63-64), (B, pos_samples) int64, -1 padded."""
        g = self.graph
        out = torch.empty((users.numel(), self.pos_samples), dtype=torch.int64, device=g.device)
        self.__dict__["_b200_pos_calls"] = self.__dict__.get("_b200_pos_calls", 0) + 1
        seed = (torch.initial_seed() * 0x9E3779B1 + self.__dict__["_b200_pos_calls"] * 0x85EBCA77) & (2 ** 63 - 1)
        with torch.cuda.device(g.device):
            ops.check(g.lib.tgcn_sample_positives(g.handle, users.numel(), self.pos_samples, users.data_ptr(), seed,
                                                  out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out

    def _sample_positives(self, users: Sequence[int]) -> torch.Tensor:
        """advanced_sampling.py:63-64: <= pos_samples random positives per user from Python's RNG, -1 padded."""
        out = np.full((len(users), self.pos_samples), -1, dtype=np.int64)
        for b, u in enumerate(users):
            positives = self.positive_lists[u]["list"]
            s = random.sample(positives, min(self.pos_samples, len(positives)))
            out[b, :len(s)] = s
        return torch.from_numpy(out)

    def select_triples(self, data, sampled_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B, 1 + C) rows [user, candidates...] -> (T, 3) int64 triples [user, pos, hardest negative], ordered
        like the reference's cat of cartesian_prod(positives, negatives) (positives outer)."""
        dev = self.graph.device
        data = data.to(dev)
        users64 = data[:, 0]
        users = ops.as_index(users64, dev)
        cands = ops.as_index(data[:, 1:], dev)
        with torch.no_grad():
            out = _PropagateFn.apply(self.embedding_user.weight.detach(), self.embedding_item.weight.detach(), self.graph,
                                     self.n_layers, self._single(), self._draw_keep_mask(), float(self.dropout))
            negs, counts, _ = ops.adv_select(self.graph, out, users, cands, max(self.k))
        if sampled_pos is None:
            sampled_pos = (self._sample_positives_device(users) if self.positive_sampler == "device"
                           else self._sample_positives(users64.tolist()))
        sampled_pos = sampled_pos.to(dev)
        valid = (sampled_pos >= 0)[:, :, None] & (negs >= 0)[:, None, :]
        b, p, n = torch.nonzero(valid, as_tuple=True)  # row-major: batch, then positive, then negative
        return torch.stack([users64[b], sampled_pos[b, p], negs[b, n].to(torch.int64)], dim=1)

    def get_loss(self, data):
        """advanced_sampling.py:46-69 (two propagations per step, each with its own dropout draw: G12)."""
        return super().get_loss(self.select_triples(data))


class AdvSamplModel(B200AdvSampl, BaseModel):
    def _copy_params(self, params):
        super()._copy_params(params)
        self.positive_sampler = getattr(params, "positive_sampler", "host")

    def _copy_dataset_params(self, dataset):
        super()._copy_dataset_params(dataset)
        self.positive_lists = getattr(dataset, "positive_lists", None)
        self.pos_samples = getattr(dataset, "pos_samples", 5)


# ------------------------------------------------------------------------------------------------
# (e) learning to rank on top of LightGCN  (ltr_models.py:38-241)
# ------------------------------------------------------------------------------------------------
class B200LTR:
    """Overrides of LTRBase / LTRLinear / LTRLinearWPop."""

    with_pop = False

    def _tabs(self):
        t = self.__dict__.get("_b200_tabs")
        if t is None:
            dev = self.graph.device
            t = {"users_rev": self.users_as_avg_reviews, "users_desc": self.users_as_avg_desc,
                 "items_rev": self.items_as_avg_reviews, "items_desc": self.items_as_desc}
            t = {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in t.items()}
            self.__dict__["_b200_tabs"] = t
        return t

    def _pop(self):
        if not self.with_pop:
            return None
        p = self.__dict__.get("_b200_pop")
        if p is None:
            dev = self.graph.device
            p = (self.popularity_users.to(device=dev, dtype=torch.float32).contiguous(),
                 self.popularity_items.to(device=dev, dtype=torch.float32).contiguous())
            self.__dict__["_b200_pop"] = p
        return p

    def collapsed_head(self):
        """The activation-free Linear stack (ltr_models.py:186-190) as one affine map: (w (F,), b) in float64."""
        layers = list(self.layers)
        w = layers[0].weight.detach().double()
        b = layers[0].bias.detach().double()
        for layer in layers[1:]:
            b = layer.weight.detach().double() @ b + layer.bias.detach().double()
            w = layer.weight.detach().double() @ w
        return w.reshape(-1), b.reshape(())

    def get_features_pairwise_fused(self, emb, users, items):
        """ltr_models.py:148-166 (+ popularity columns :234-241) from one gather kernel."""
        dev = self.graph.device
        return _LtrPairFeaturesFn.apply(emb, self.n_users, ops.as_index(users, dev), ops.as_index(items, dev), self._tabs(), self._pop())

    def score_pairwise_ltr(self, users_emb, items_emb, users, items):
        """ltr_models.py:206-210 / :234-241 -> (B, 1).  ``users_emb``/``items_emb`` are the gathered rows."""
        tabs, pop = self._tabs(), self._pop()
        def sm(x, y):
            return (x * y).sum(dim=1).unsqueeze(1)
        ur, ud, ir, idesc = tabs["users_rev"][users], tabs["users_desc"][users], tabs["items_rev"][items], tabs["items_desc"][items]
        f = torch.cat([sm(users_emb, items_emb), sm(ur, ir), sm(ud, idesc), sm(ur, idesc), sm(ud, ir)], dim=1)
        if pop is not None:
            f = torch.cat([f, pop[0][users], pop[1][items]], dim=-1)
        return self.layers(f)

    def bpr_loss(self, users, pos, negs):
        """base_model.py:186-198 with the LTR pairwise score: fused feature gathers + the tiny head in torch."""
        out = _PropagateFn.apply(self.embedding_user.weight, self.embedding_item.weight, self.graph, self.n_layers,
                                 self._single(), self._draw_keep_mask(), float(self.dropout))
        pos_scores = self.layers(self.get_features_pairwise_fused(out, users, pos))
        loss = 0
        for neg in negs:
            neg_scores = self.layers(self.get_features_pairwise_fused(out, users, neg))
            loss = loss + torch.mean(torch.nn.functional.selu(neg_scores - pos_scores))
        loss = loss / len(negs)
        self._loss_values["bpr"] += loss.detach()
        return loss

    def get_loss(self, data):
        users, pos, *negs = data.to(self.graph.device).t()
        return self.bpr_loss(users, pos, negs) + self.reg_loss(users, pos, negs)

    def _rank(self, emb, users, k):
        """score_batchwise_ltr (ltr_models.py:200-204, :227-232) + mask + top-k as ONE contraction of width
        d + 2·D: the head is affine, so its weights are folded into the packed item operand."""
        tabs, pop = self._tabs(), self._pop()
        w, b = self.collapsed_head()
        nu = self.n_users
        items_p = ops.ltr_pack_items(emb[nu:], tabs["items_rev"], tabs["items_desc"], w[:5].tolist())
        users_p = ops.ltr_pack_users(users, emb[:nu], tabs["users_rev"], tabs["users_desc"])
        user_bias = torch.full((users.numel(),), float(b), dtype=torch.float32, device=emb.device)
        item_bias = None
        if pop is not None:
            user_bias = (user_bias.double() + w[5] * pop[0][users.long(), 0].double()).float().contiguous()
            item_bias = (w[6] * pop[1][:, 0].double()).float().contiguous()
        return ops.eval_topk(self.graph, users_p, items_p, k, users=users, user_bias=user_bias, item_bias=item_bias,
                             by_position=True, precision=self.eval_precision)


class LTRLinear(B200LTR, BaseModel):
    def _copy_params(self, params):
        super()._copy_params(params)
        self.load_base = getattr(params, "load_base", None)
        self.freeze = getattr(params, "freeze", False)

    def _copy_dataset_params(self, dataset):
        super()._copy_dataset_params(dataset)
        self.items_as_avg_reviews = dataset.items_as_avg_reviews
        self.users_as_avg_reviews = dataset.users_as_avg_reviews
        self.users_as_avg_desc = dataset.users_as_avg_desc
        self.items_as_desc = dataset.items_as_desc
        self.all_items = getattr(dataset, "all_items", range(dataset.n_items))

    def _init_embeddings(self, emb_size):
        super()._init_embeddings(emb_size)
        if self.freeze:
            self.embedding_user.requires_grad_(False)
            self.embedding_item.requires_grad_(False)

    def _add_vars(self, params):
        super()._add_vars(params)
        self._ltr_ready = False
        if self.load_base:  # base model is loaded and evaluated with plain LightGCN scoring first (G18)
            self.load_model(self.load_base)
        self.feature_names = ["lightgcn score", "reviews", "desc", "reviews-description", "description-reviews"]
        self._setup_layers(params)
        self._ltr_ready = True

    def _setup_layers(self, params):
        layer_sizes = [len(self.feature_names)] + list(getattr(params, "ltr_layers", [])) + [1]
        self.layers = nn.Sequential(*[nn.Linear(i, j) for i, j in zip(layer_sizes, layer_sizes[1:])]).to(self.device)

    def _rank(self, emb, users, k):
        if not self._ltr_ready:
            return B200HotPath._rank(self, emb, users, k)
        return B200LTR._rank(self, emb, users, k)

    def get_loss(self, data):
        return B200LTR.get_loss(self, data)

    def evaluate(self, *args, **kwargs):
        """ltr_models.py:192-198."""
        if len(self.layers) == 1:
            self.logger.info("Feature weights from the top layer:")
            for f, w in zip(self.feature_names, self.layers[0].weight.tolist()[0]):
                self.logger.info(f"{f:<20} {w:.4}")
        return super().evaluate(*args, **kwargs)


class LTRLinearWPop(LTRLinear):
    with_pop = True

    def _copy_dataset_params(self, dataset):
        super()._copy_dataset_params(dataset)
        self.popularity_users = dataset.popularity_users
        self.popularity_items = dataset.popularity_items

    def _setup_layers(self, params):
        self.feature_names += ["user popularity", "item popularity"]
        super()._setup_layers(params)
