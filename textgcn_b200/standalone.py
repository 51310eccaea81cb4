"""Standalone model classes: the kernel-backed mixins of ``models.py`` on a small shell of our own, for use where the
reference package is not importable (and for the extras the reference has no counterpart for: the fused Adam, the
CUDA-graph training step, the device samplers, the sync-free NaN check).

Where the reference IS importable, the drop-in is ``textgcn_b200.dropin`` — the same mixins in front of the reference's
own classes, whose shell (fit loop, checkpointing, logging) then runs unchanged.  What this shell keeps compatible with
the reference is its CONTRACT, not its code: ``Model(params, dataset)`` (base_model.py:23-31) with the subclass hooks
``_copy_params`` / ``_copy_dataset_params`` / ``_init_embeddings`` / ``_add_vars``, ``fit(batches)``, ``predict`` /
``evaluate``, the checkpoint files ``latest_checkpoint.pkl`` / ``best.pkl`` holding a state_dict with the keys
``embedding_user.weight``, ``embedding_item.weight``, ``layers.N.*`` (base_model.py:278-299), the stopping rule of
utils.py:79-90, and the late re-binding of the LTR scoring methods (ltr_models.py:172-179, G18).
"""
from __future__ import annotations

import os
import shutil
from collections import defaultdict

import numpy as np
import torch
from torch import nn

from . import metrics as M
from ._lib import TgcnError
from .models import B200AdvSampl, B200HotPath, B200LTR

_PARAM_FIELDS = ("k", "lr", "uid", "save", "quiet", "epochs", "logger", "device", "dropout", "emb_size", "n_layers", "save_path",
                 "batch_size", "reg_lambda", "evaluate_every", "neg_samples")
_EXTRA_FIELDS = dict(fused_adam=False, dropout_rng="host", eval_precision="auto", nan_check="step", cuda_graph=False,
                     strict_fused=False)


def early_stop(history) -> bool:
    """Stopping rule of utils.py:79-90 on the metric history {name: (n_evals, len(k)) array}: stop once the last three
    evaluations are either strictly declining for every metric and k, or within 1e-4 of each other."""
    if len(history["recall"]) < 3:
        return False
    last = [(np.asarray(m[-3]), np.asarray(m[-2]), np.asarray(m[-1])) for m in history.values()]
    declining = all((c < b).all() and (b < a).all() for a, b, c in last)
    flat = all(np.allclose(c, b, atol=1e-4) and np.allclose(c, a, atol=1e-4) for a, b, c in last)
    return declining or flat


class BaseModel(B200HotPath, nn.Module):
    """LightGCN with BPR on the sm_100a kernels; constructor contract ``Model(params, dataset)``."""

    def __init__(self, params, dataset):
        super().__init__()
        self._copy_params(params)
        self._copy_dataset_params(dataset)
        self._init_embeddings(params.emb_size)
        self._add_vars(params)
        self.load_model(getattr(params, "load", None))
        self.to(params.device)

    # -- construction hooks (same names as the reference's, so subclasses extend them the same way) ----
    def _copy_params(self, params):
        for name in _PARAM_FIELDS:
            setattr(self, name, getattr(params, name))
        for name, default in _EXTRA_FIELDS.items():
            setattr(self, name, getattr(params, name, default))
        self.device = torch.device(self.device)
        if self.device.type != "cuda":
            raise TgcnError("textgcn_b200 models need a CUDA device: there is no CPU fallback")
        self.slurm = params.slurm or params.quiet
        if getattr(params, "single", False):
            self.layer_combination = self.layer_combination_single

    def _copy_dataset_params(self, dataset):
        self.n_users, self.n_items = dataset.n_users, dataset.n_items
        self.norm_matrix = dataset.norm_matrix
        self.true_test_lil = dataset.true_test_lil
        self.train_user_dict = getattr(dataset, "train_user_dict", None)
        self.test_users = (np.asarray(dataset.test_users) if hasattr(dataset, "test_users")
                           else np.sort(dataset.test_df.user_id.unique()))
        if hasattr(dataset, "user_mapping"):
            self.user_mapping_dict = dict(dataset.user_mapping[["remap_id", "org_id"]].values)
            self.item_mapping_dict = dict(dataset.item_mapping[["remap_id", "org_id"]].values)
        if getattr(dataset, "graph", None) is not None:  # a prebuilt CSR handle (synthetic workloads skip the COO)
            self.__dict__["_b200_graph"] = dataset.graph

    def _init_embeddings(self, emb_size):
        self.embedding_user = nn.Embedding(self.n_users, emb_size, device=self.device)
        self.embedding_item = nn.Embedding(self.n_items, emb_size, device=self.device)
        for table in (self.embedding_user, self.embedding_item):
            nn.init.normal_(table.weight, std=0.1)  # base_model.py:63-68

    def _add_vars(self, params):
        self.metrics = list(M.METRICS)
        self._reset_history()
        self.training = False
        self._loss_values = defaultdict(float)

    def _reset_history(self):
        self.metrics_logger = {name: np.zeros((0, len(self.k))) for name in self.metrics}

    # -- API kept for parity; `representation` fuses both into the last SpMM pass ------------------------
    def layer_combination(self, vectors):
        return torch.mean(torch.stack(vectors), axis=0)

    def layer_combination_single(self, vectors):
        return vectors[-1]

    @property
    def embedding_matrix(self):
        return torch.cat([self.embedding_user.weight, self.embedding_item.weight])

    # -- training loop --------------------------------------------------------------------------------
    def _make_optimizer(self):
        """(optimizer, graphed step | None).  ``cuda_graph``: the whole step — mask draw, propagate, fused BPR, Horner
        backward, fused Adam — replayed from one CUDA graph (fixed-shape BPR batches only, not AdvSampl / LTR)."""
        if self.cuda_graph and type(self).get_loss is B200HotPath.get_loss:
            from .optim import FusedAdam
            from .train_graph import GraphedTrainStep
            opt = FusedAdam(self.parameters(), lr=self.lr, capturable=True)
            return opt, GraphedTrainStep(self, opt)
        if self.fused_adam:
            from .optim import FusedAdam
            return FusedAdam(self.parameters(), lr=self.lr), None
        return torch.optim.Adam(self.parameters(), lr=self.lr), None

    def _train_epoch(self, batches, graphed, epoch):
        nan_seen = torch.zeros((), dtype=torch.bool, device=self.device)
        for data in batches:
            if graphed is not None:
                nan_seen |= graphed(data).isnan()
                continue
            self.optimizer.zero_grad()
            loss = self.get_loss(data)
            if self.nan_check == "step":   # the reference's per-step device sync (base_model.py:123, G7)
                assert not loss.isnan(), f"loss is NA at epoch {epoch}"
            else:                          # the same check without draining the stream every step
                nan_seen |= loss.detach().isnan()
            loss.backward()
            self.optimizer.step()
        assert not bool(nan_seen), f"loss is NA at epoch {epoch}"
        if graphed is not None:  # the graph accumulates [bpr, reg] in place: hand the epoch's sums to the logger
            sums = graphed.loss_sums.clone()
            graphed.loss_sums.zero_()
            self._loss_values["bpr"], self._loss_values["reg"] = sums[0], sums[1]
        fails = getattr(batches, "fail_count", None)
        if fails is not None and int(fails) > 0:
            raise TgcnError(f"the device sampler could not find a negative for {int(fails)} rows this epoch (users whose "
                            "train items cover the catalogue); the fused BPR kernel skipped them")

    def fit(self, batches):
        """``batches``: any iterable of (B, 2 + n_neg) int64 batches — the reference's DataLoader (main.py:35) or the
        device samplers of ``textgcn_b200.sampler``.  Evaluates, checkpoints and applies the stopping rule every
        ``evaluate_every`` epochs, and checkpoints once more at the end of an uninterrupted run (base_model.py:108-139)."""
        self.optimizer, graphed = self._make_optimizer()
        stopped = False
        for epoch in range(1, self.epochs + 1):
            self.train()
            self.training = True
            self._loss_values = defaultdict(float)
            self._train_epoch(batches, graphed, epoch)
            if epoch % self.evaluate_every == 0:
                self.logger.info(f"Epoch {epoch}: " + " ".join(f"{k} = {float(v):.4f}" for k, v in self._loss_values.items()))
                self.evaluate(epoch)
                self.checkpoint(epoch)
                if early_stop(self.metrics_logger):
                    self.logger.warning(f"Early stopping triggerred at epoch {epoch}")
                    stopped = True
                    break
        if not stopped:
            self.checkpoint(self.epochs)

    # -- checkpoint files (formats of base_model.py:278-299: interchangeable with the reference's) ---------
    def load_model(self, load_path):
        if load_path is None:
            return
        path = os.path.join(load_path, "best.pkl") if os.path.isdir(load_path) else load_path
        self.logger.info(f"Loading model {path}")
        self.load_state_dict(torch.load(path, map_location=self.device))
        self.logger.info("Performance of the loaded model:")
        self.evaluate()
        self._reset_history()

    def checkpoint(self, epoch):
        if not self.save:
            return
        os.makedirs(self.save_path, exist_ok=True)
        latest = os.path.join(self.save_path, "latest_checkpoint.pkl")
        torch.save(self.state_dict(), latest)
        first_metric = self.metrics_logger[self.metrics[0]]
        if len(first_metric) and first_metric[:, 0].max() == first_metric[-1][0]:
            self.logger.info(f"Updating best model at epoch {epoch}")
            shutil.copyfile(latest, os.path.join(self.save_path, "best.pkl"))


class AdvSamplModel(B200AdvSampl, BaseModel):
    def _copy_params(self, params):
        super()._copy_params(params)
        self.positive_sampler = getattr(params, "positive_sampler", "host")

    def _copy_dataset_params(self, dataset):
        super()._copy_dataset_params(dataset)
        self.positive_lists = getattr(dataset, "positive_lists", None)
        self.pos_samples = getattr(dataset, "pos_samples", 5)


class LTRLinear(B200LTR, BaseModel):
    FEATURES = ("lightgcn score", "reviews", "desc", "reviews-description", "description-reviews")  # ltr_models.py:71-77

    def __init__(self, params, dataset):
        super().__init__(params, dataset)
        # the scoring methods switch to their LTR versions only now: a base model loaded inside _add_vars was evaluated
        # with plain LightGCN scoring (ltr_models.py:172-179, G18)
        self.evaluate = self.evaluate_ltr
        self.score_pairwise = self.score_pairwise_ltr
        self.score_batchwise = self.score_batchwise_ltr

    def _copy_params(self, params):
        super()._copy_params(params)
        self.load_base = getattr(params, "load_base", None)
        self.freeze = getattr(params, "freeze", False)

    def _copy_dataset_params(self, dataset):
        super()._copy_dataset_params(dataset)
        for name in ("items_as_avg_reviews", "users_as_avg_reviews", "users_as_avg_desc", "items_as_desc"):
            setattr(self, name, getattr(dataset, name))
        self.all_items = getattr(dataset, "all_items", range(dataset.n_items))

    def _init_embeddings(self, emb_size):
        super()._init_embeddings(emb_size)
        if self.freeze:
            self.embedding_user.requires_grad_(False)
            self.embedding_item.requires_grad_(False)

    def _add_vars(self, params):
        super()._add_vars(params)
        if self.load_base:
            self.load_model(self.load_base)
        self.feature_names = list(self.FEATURES)
        self._setup_layers(params)

    def _setup_layers(self, params):
        sizes = [len(self.feature_names), *getattr(params, "ltr_layers", []), 1]
        self.layers = nn.Sequential(*(nn.Linear(i, o) for i, o in zip(sizes, sizes[1:]))).to(self.device)


class LTRLinearWPop(LTRLinear):
    with_pop = True

    def _copy_dataset_params(self, dataset):
        super()._copy_dataset_params(dataset)
        self.popularity_users = dataset.popularity_users
        self.popularity_items = dataset.popularity_items

    def _setup_layers(self, params):
        self.feature_names += ["user popularity", "item popularity"]
        super()._setup_layers(params)
