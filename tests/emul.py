"""TEST INFRASTRUCTURE: a CPU emulation of the ``textgcn_b200.ops`` entry points (plain torch / numpy restatements of what
each kernel computes), installed with ``install(monkeypatch)``.  It exists so that the HOST logic of the mixins — method
resolution over the reference's classes, the G18 load order, autograd wiring, triple construction, shapes — is exercised
by the ``-m "not gpu"`` suite here (no GPU in the authoring container) with the same scenarios the GPU suite runs on the
real kernels (tests/dropin_scenarios.py).  Never imported by the product."""
from __future__ import annotations

import numpy as np
import torch

from oracle import lightgcn_oracle as O


class EmulGraph:
    def __init__(self, norm, n_users, n_items):
        self.norm = norm.coalesce().cpu()
        self.n_users, self.n_items = int(n_users), int(n_items)
        self.device = torch.device("cpu")
        self.nnz = self.norm._nnz()
        idx = self.norm.indices()
        self.rowptr = torch.from_numpy(O.coo_to_csr(idx[0].numpy(), n_users + n_items))
        self.col = idx[1]
        self.lists = [self.col[self.rowptr[u]:self.rowptr[u + 1]].numpy() - n_users for u in range(n_users)]

    @classmethod
    def from_norm_matrix(cls, norm, n_users, n_items, device=None):
        return cls(norm, n_users, n_items)

    @property
    def n_nodes(self):
        return self.n_users + self.n_items

    def matrix(self, keep, dropout, transposed=False):
        m = self.norm
        if keep is not None:
            m = O.dropout_matrix(m, keep.bool(), dropout)
        return m.t().coalesce() if transposed else m


def propagate_fwd(g, user_w, item_w, n_layers, single=False, keep=None, dropout=0.0, out=None):
    m = g.matrix(keep, dropout)
    cur = torch.cat([user_w, item_w])
    layers = [cur]
    for _ in range(n_layers):
        cur = torch.sparse.mm(m, cur)
        layers.append(cur)
    res = layers[-1] if single else torch.mean(torch.stack(layers), 0)
    if out is not None:
        out.copy_(res)
        return out
    return res


def propagate_bwd(g, grad_out, n_layers, single=False, keep=None, dropout=0.0, grad_in=None, accumulate=False):
    mt = g.matrix(keep, dropout, transposed=True)
    h = grad_out
    for _ in range(n_layers):
        h = torch.sparse.mm(mt, h) + (0 if single else grad_out)
    res = h if single else h / (n_layers + 1)
    if grad_in is None:
        return res
    if accumulate:
        grad_in.add_(res)
    else:
        grad_in.copy_(res)
    return grad_in


def spmm(g, x, out=None):
    return torch.sparse.mm(g.norm, x)


@torch.enable_grad()  # called from inside autograd.Function.forward
def bpr_fwd_bwd(n_users, n_items, emb, user_w, item_w, users, pos, negs, reg_lambda, grad_emb, grad_w0):
    e = emb.detach().clone().requires_grad_(True)
    w = torch.cat([user_w, item_w]).detach().clone().requires_grad_(True)
    u, p, n = users.long(), pos.long(), negs.long().reshape(-1, users.numel())
    ue = e[u]
    ps = (ue * e[n_users + p]).sum(1)
    bpr = sum(torch.nn.functional.selu((ue * e[n_users + nj]).sum(1) - ps).mean() for nj in n) / n.shape[0]
    reg = (w[u].square().sum() + w[n_users + p].square().sum() + w[n_users + n].square().sum()) * (reg_lambda / (2 * u.numel()))
    if grad_emb is not None:
        grad_emb.add_(torch.autograd.grad(bpr, e)[0])
    if grad_w0 is not None and reg_lambda != 0:
        grad_w0.add_(torch.autograd.grad(reg, w)[0])
    return torch.stack([bpr.detach(), reg.detach()])


def eval_topk(mask_graph, user_vecs, item_vecs, k, users=None, n_rank=None, item_range=None, user_bias=None, item_bias=None,
              finalize=True, by_position=False, precision="auto"):
    ids_u = users.long() if users is not None else torch.arange(n_rank or user_vecs.shape[0])
    uv = user_vecs if (by_position or users is None) else user_vecs[ids_u]
    sc = (uv.double() @ item_vecs.double().T).float()
    if user_bias is not None:
        sc = sc + (user_bias if by_position else user_bias[ids_u])[:, None]
    if item_bias is not None:
        sc = sc + item_bias[None, :]
    sc = sc.numpy().copy()
    if mask_graph is not None:
        for r, u in enumerate(ids_u.tolist()):
            sc[r, mask_graph.lists[u]] = -np.inf
    ids, scores = O.canonical_topk(sc, k)
    return torch.from_numpy(ids.astype(np.int32)), torch.from_numpy(scores.astype(np.float32))


def adv_select(g, emb, users, cands, kmax, want_scores=False):
    nu = g.n_users
    b, c = cands.shape
    sc = torch.einsum("bd,bcd->bc", emb[users.long()], emb[nu + cands.long()])
    negs = torch.full((b, kmax), -1, dtype=torch.int32)
    counts = torch.zeros(b, dtype=torch.int32)
    for r in range(b):
        order = np.lexsort((np.arange(c), -sc[r].numpy()))  # score desc, candidate position asc
        pos = set(g.lists[int(users[r])].tolist())
        keep = [int(cands[r, j]) for j in order if int(cands[r, j]) not in pos][:kmax]
        negs[r, :len(keep)] = torch.tensor(keep, dtype=torch.int32)
        counts[r] = len(keep)
    return negs, counts, (sc if want_scores else None)


def ltr_pairwise_features(n_users, emb, users, items, users_rev, users_desc, items_rev, items_desc, pop_users=None, pop_items=None):
    u, i = users.long(), items.long()
    f = O.ltr_features_pairwise(emb[u], users_rev[u], users_desc[u], emb[n_users + i], items_rev[i], items_desc[i])
    if pop_users is not None:
        f = torch.cat([f, pop_users[u].reshape(-1, 1), pop_items[i].reshape(-1, 1)], 1)
    return f


def ltr_pairwise_emb_bwd(n_users, emb, users, items, gf0, grad_emb):
    u, i = users.long(), n_users + items.long()
    grad_emb.index_add_(0, u, gf0[:, None] * emb[i])
    grad_emb.index_add_(0, i, gf0[:, None] * emb[u])


def ltr_pack_items(items_emb, items_rev, items_desc, w5):
    w = [float(x) for x in w5]
    return torch.cat([w[0] * items_emb, w[1] * items_rev + w[3] * items_desc, w[2] * items_desc + w[4] * items_rev], 1)


def ltr_pack_users(users, users_emb, users_rev, users_desc):
    u = users.long() if users is not None else torch.arange(users_emb.shape[0])
    return torch.cat([users_emb[u], users_rev[u], users_desc[u]], 1)


def score_batchwise(a, b, row_bias=None, col_bias=None, out=None, plane=None):
    s = a @ b.T
    if row_bias is not None:
        s = s + row_bias[:, None]
    if col_bias is not None:
        s = s + col_bias[None, :]
    if out is None:
        return s
    (out if plane is None else out[:, :, plane]).copy_(s)
    return out


def score_pairwise_adv(users_emb, items_emb):
    return torch.einsum("bd,bcd->bc", users_emb, items_emb)


def ltr_features_rows(ue, ie, ur, ud, ir, idesc):
    return O.ltr_features_pairwise(ue, ur, ud, ie, ir, idesc)


def topk_metrics(pred_ids, true_ptr, true_ids, ks):
    ptr = true_ptr.numpy()
    truth = [true_ids[ptr[r]:ptr[r + 1]].tolist() for r in range(len(ptr) - 1)]
    res = O.calculate_metrics(pred_ids.tolist(), truth, ks)
    return torch.tensor([[res[m][ki] for m in ("recall", "precision", "hit", "ndcg", "f1")] for ki in range(len(ks))], dtype=torch.float64)


_FUNCS = ["propagate_fwd", "propagate_bwd", "spmm", "bpr_fwd_bwd", "eval_topk", "adv_select", "ltr_pairwise_features",
          "ltr_pairwise_emb_bwd", "ltr_pack_items", "ltr_pack_users", "score_batchwise", "score_pairwise_adv", "ltr_features_rows",
          "topk_metrics"]


def install(monkeypatch):
    """Route ``textgcn_b200.ops`` through the restatements above and let the mixins build their graph on the CPU."""
    from textgcn_b200 import metrics, models, ops
    for name in _FUNCS:
        monkeypatch.setattr(ops, name, globals()[name])

    def graph(self):
        g = self.__dict__.get("_b200_graph")
        if g is None:
            g = self.__dict__["_b200_graph"] = EmulGraph(self.norm_matrix, self.n_users, self.n_items)
        return g

    monkeypatch.setattr(models.B200HotPath, "graph", property(graph))
    real = metrics.calculate_metrics

    def calc(pred_ids, y_true, ks):   # the product refuses CPU tensors; the emulation feeds it one
        truth = y_true if isinstance(y_true, metrics.TruthCSR) else metrics.TruthCSR.from_lists(y_true, "cpu")
        ks = sorted(int(k) for k in ks)
        vals = topk_metrics(pred_ids.to(torch.int32), truth.ptr, truth.ids, ks).numpy()
        return {m: [float(vals[ki, mi]) for ki in range(len(ks))] for mi, m in enumerate(metrics.METRICS)}

    monkeypatch.setattr(metrics, "calculate_metrics", calc)
    return real
