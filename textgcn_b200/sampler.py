"""GPU BPR sampler: an iterable of (B, 2 + n_neg) int64 CUDA batches that replaces the reference's
``BaseDataset.__getitem__`` + ``DataLoader(shuffle=True)`` (dataset.py:167-193, main.py:35) in ``model.fit(batches)``.

Epoch structure follows the reference: every user appears ``bucket_len = n_train // n_users`` times per epoch in a
shuffled order; each row draws a uniform positive (``random.choices``) and ``neg_samples`` uniform non-positive items.
The reference additionally keeps a user's negatives distinct across its whole epoch bucket; here they are distinct
within a row.  Parity is statistical only (host ``random`` cannot be reproduced), as SURVEY.md H6 notes.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import check


class BprEpochSampler:
    def __init__(self, graph: ops.Graph, batch_size: int = 2048, neg_samples: int = 1, seed: int = 0, max_tries: int = 32):
        self.graph, self.batch_size, self.neg_samples, self.seed, self.max_tries = graph, batch_size, neg_samples, seed, max_tries
        n_train = int(graph.rowptr[graph.n_users])
        self.bucket_len = max(1, n_train // graph.n_users)
        self.rows = self.bucket_len * graph.n_users
        self.epoch = 0
        self.fail_count = torch.zeros(1, dtype=torch.int32, device=graph.device)

    def __len__(self) -> int:
        return (self.rows + self.batch_size - 1) // self.batch_size

    def sample(self, users: torch.Tensor, seed: int) -> torch.Tensor:
        """(B, 2 + n_neg) int64 rows for the given int32 user ids."""
        g = self.graph
        users = ops._chk(users, torch.int32, "users", 1, align=4)
        out = torch.empty((users.numel(), 2 + self.neg_samples), dtype=torch.int64, device=g.device)
        with torch.cuda.device(g.device):
            check(g.lib.tgcn_sample_bpr_batch(g.handle, users.numel(), self.neg_samples, users.data_ptr(), seed & (2 ** 64 - 1),
                                              self.max_tries, out.data_ptr(), self.fail_count.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        return out

    def __iter__(self):
        g = self.graph
        gen = torch.Generator(device=g.device).manual_seed(self.seed * 1_000_003 + self.epoch)
        order = torch.randperm(self.rows, generator=gen, device=g.device)
        users = (order // self.bucket_len).to(torch.int32)  # row r belongs to user r // bucket_len (dataset.py:190)
        base = (self.seed * 0x9E3779B1 + self.epoch * 0x85EBCA77) & (2 ** 63 - 1)
        self.epoch += 1
        for i, s in enumerate(range(0, self.rows, self.batch_size)):
            yield self.sample(users[s:s + self.batch_size].contiguous(), base + i * 0xC2B2AE3D)


class AdvEpochSampler(BprEpochSampler):
    """Batches for AdvSamplModel: rows [user, 1000 distinct random candidate items] (advanced_sampling.py:10-22),
    generated on the device — the reference spends 7.3 s of an 8.7 s step in ``random.sample`` here."""

    max_neg_samples = 1000

    def __init__(self, graph: ops.Graph, batch_size: int = 2048, seed: int = 0):
        super().__init__(graph, batch_size, 1, seed)
        self.n_cand = min(graph.n_items, self.max_neg_samples)

    def sample(self, users: torch.Tensor, seed: int) -> torch.Tensor:
        g = self.graph
        users = ops._chk(users, torch.int32, "users", 1, align=4)
        out = torch.empty((users.numel(), 1 + self.n_cand), dtype=torch.int64, device=g.device)
        with torch.cuda.device(g.device):
            check(g.lib.tgcn_sample_candidates(g.n_items, users.numel(), self.n_cand, users.data_ptr(), seed & (2 ** 64 - 1),
                                               out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out
