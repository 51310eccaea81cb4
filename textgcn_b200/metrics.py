"""recall / precision / hit / ndcg / f1 @k from the (n, k) id table on the device (reference: utils.py:11-63).

SURVEY.md §8f n4: the reference's pandas row-apply takes 65 % of ``evaluate`` and ingests a Python list per user.  Here
the true test items are a CSR (``TruthCSR``: ptr int64 (n+1), ids int32) resident on the device and ONE kernel
(``tgcn_topk_metrics``, csrc/metrics.cu) turns the (n, kmax) id table of the fused eval kernel into the five metrics
for every k — no Python loop over users, 10M users in a few milliseconds.  Semantics are the reference's (distinct
intersection, recall over len(y_true) with duplicates, ndcg relevance on every occurrence, f1 = 0 where the
denominator is 0, float64 means).  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Union

import numpy as np
import torch

from . import ops
from ._lib import TgcnError

METRICS = ["recall", "precision", "hit", "ndcg", "f1"]


class TruthCSR:
    """``true_test_lil`` (base_model.py:57, dataset.py:118-120) as a device CSR: row r lists the test items of the r-th
    ranked user."""

    def __init__(self, ptr: torch.Tensor, ids: torch.Tensor):
        self.ptr, self.ids = ptr, ids

    @property
    def n_rows(self) -> int:
        return self.ptr.numel() - 1

    @classmethod
    def from_lists(cls, lists: Sequence[Sequence[int]], device) -> "TruthCSR":
        """From the reference's list of lists (one flat concatenation on the host, done once per model)."""
        lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists))
        ptr = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum(lens, out=ptr[1:])
        flat = np.fromiter((i for x in lists for i in x), dtype=np.int32, count=int(ptr[-1]))
        return cls(torch.from_numpy(ptr).to(device), torch.from_numpy(flat).to(device))

    @classmethod
    def from_pairs(cls, rows: torch.Tensor, items: torch.Tensor, n_rows: int) -> "TruthCSR":
        """From (row, item) pairs already on the device (any order; the order inside a row is kept): torch ops only."""
        rows = rows.to(torch.int64)
        order = torch.argsort(rows, stable=True)
        ptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=rows.device)
        ptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n_rows), 0)
        return cls(ptr, items[order].to(torch.int32).contiguous())


def calculate_metrics(pred_ids: torch.Tensor, y_true: Union[TruthCSR, Sequence[Sequence[int]]], ks: Sequence[int]
                      ) -> Dict[str, List[float]]:
    """{metric: [value at each k of sorted(ks)]} like utils.calculate_metrics (utils.py:36-63)."""
    if not isinstance(pred_ids, torch.Tensor) or not pred_ids.is_cuda:
        raise TgcnError("calculate_metrics needs the CUDA id table of predict_device (textgcn_b200 has no CPU path)")
    truth = y_true if isinstance(y_true, TruthCSR) else TruthCSR.from_lists(y_true, pred_ids.device)
    ks = sorted(int(k) for k in ks)
    vals = ops.topk_metrics(pred_ids.to(torch.int32).contiguous(), truth.ptr, truth.ids, ks).cpu().numpy()
    return {m: [float(vals[ki, mi]) for ki in range(len(ks))] for mi, m in enumerate(METRICS)}
