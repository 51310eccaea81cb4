"""Generate tests/golden/*.npz by running the UNMODIFIED reference (authoring container only).

TEST INFRASTRUCTURE ONLY.  Usage (from the repo root, needs /root/reference):

    python oracle/make_golden.py

The reference has no tests or golden vectors of its own, so the oracle and
the CUDA path are pinned against outputs of the reference itself: this script imports
``/root/reference`` read-only (through oracle/ref_shims.py), drives its public classes
(BaseDataset/BaseModel, AdvSamplModel, LTRLinear, LTRLinearWPop) on ``data/dummy`` and on
small seeded synthetic TSV datasets written to a temp dir, and records inputs + outputs.
Nothing from the reference is copied into the repo; only numeric vectors are stored.
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
from TextGCN import AdvSamplModel, BaseDataset, BaseModel, LTRLinear, LTRLinearWPop  # noqa: E402
from TextGCN.advanced_sampling import AdvSamplDataset  # noqa: E402
from TextGCN.parser import parse_args  # noqa: E402
from TextGCN.utils import calculate_metrics  # noqa: E402
import pandas as pd  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def write_dataset(folder, n_users, n_items, n_train, seed, n_test_max=3):
    """Seeded random TSV dataset; string ids chosen so the reference's string sort matters."""
    rng = np.random.default_rng(seed)
    pairs = set()
    for u in range(n_users):
        pairs.add((u, int(rng.integers(n_items))))
    for i in range(n_items):
        pairs.add((int(rng.integers(n_users)), i))
    wi = rng.lognormal(0, 1.0, n_items)
    wi /= wi.sum()
    while len(pairs) < n_train:
        pairs.add((int(rng.integers(n_users)), int(rng.choice(n_items, p=wi))))
    pairs = sorted(pairs)
    test = []
    for u in range(n_users):
        have = {i for (uu, i) in pairs if uu == u}
        free = [i for i in range(n_items) if i not in have]
        for i in rng.choice(free, size=int(rng.integers(1, n_test_max + 1)), replace=False):
            test.append((u, int(i)))
    os.makedirs(folder, exist_ok=True)
    pd.DataFrame([(f"user_{u}", f"asin_{i}") for u, i in pairs], columns=["user_id", "asin"]).to_csv(
        os.path.join(folder, "train.tsv"), sep="\t", index=False)
    pd.DataFrame([(f"user_{u}", f"asin_{i}") for u, i in test], columns=["user_id", "asin"]).to_csv(
        os.path.join(folder, "test.tsv"), sep="\t", index=False)


def args_for(model, data, ks, extra=()):
    return parse_args(["--model", model, "-d", data, "-k", *map(str, ks), "--gpu", "", "--quiet", "--slurm",
                       "--uid", "golden", *extra])


def dataset_arrays(ds):
    coo = ds.norm_matrix
    out = dict(
        n_users=np.int64(ds.n_users), n_items=np.int64(ds.n_items),
        train_u=ds.train_df.user_id.values.astype(np.int64), train_i=ds.train_df.asin.values.astype(np.int64),
        test_u=ds.test_df.user_id.values.astype(np.int64), test_i=ds.test_df.asin.values.astype(np.int64),
        norm_row=coo._indices()[0].numpy(), norm_col=coo._indices()[1].numpy(), norm_val=coo._values().numpy(),
    )
    return out


def make_batch(ds, rng, batch, n_neg):
    rows = []
    for _ in range(batch):
        u = int(rng.integers(ds.n_users))
        pos = ds.positive_lists[u]["list"]
        p = int(pos[int(rng.integers(len(pos)))])
        negs = []
        while len(negs) < n_neg:
            c = int(rng.integers(ds.n_items))
            if c not in ds.positive_lists[u]["set"]:
                negs.append(c)
        rows.append([u, p] + negs)
    return torch.tensor(rows, dtype=torch.int64)


def loss_and_grads(model, batch, seed, training):
    """get_loss + backward exactly as fit() does (base_model.py:121-125); returns the dropout
    keep-mask the reference drew (replayed from the same CPU generator seed, :82)."""
    model.zero_grad()
    model.training = training
    nnz = model.norm_matrix._values().shape[0]
    torch.manual_seed(seed)
    keep = (torch.rand(nnz) < (1 - model.dropout)) if training else torch.ones(nnz, dtype=torch.bool)
    torch.manual_seed(seed)
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    loss = model.get_loss(batch)
    loss.backward()
    out = dict(keep=keep.numpy(), loss=loss.detach().numpy(),
               bpr=torch.as_tensor(model._loss_values["bpr"]).detach().numpy(),
               reg=torch.as_tensor(model._loss_values["reg"]).detach().numpy())
    for name, p in model.named_parameters():
        out["grad_" + name.replace(".", "_")] = (p.grad.detach().numpy().copy() if p.grad is not None
                                                 else np.zeros(0, np.float32))
    model.training = False
    return out


def lgcn_case(name, data, ks, extra, batch_rows, n_neg, seed):
    torch.manual_seed(seed)
    args = args_for("lgcn", data, ks, ["--neg_samples", str(n_neg), *extra])
    ds = BaseDataset(args)
    model = BaseModel(args, ds)
    g = dataset_arrays(ds)
    g.update(user_w=model.embedding_user.weight.detach().numpy().copy(),
             item_w=model.embedding_item.weight.detach().numpy().copy(),
             n_layers=np.int64(args.n_layers), dropout=np.float64(args.dropout), reg_lambda=np.float64(args.reg_lambda),
             single=np.bool_(args.single), ks=np.asarray(args.k, dtype=np.int64))
    model.training = False
    with torch.no_grad():
        ue, ie = model.representation
    g.update(rep_user=ue.numpy().copy(), rep_item=ie.numpy().copy())
    preds, scores = model.predict(model.test_users, with_scores=True)
    g.update(test_users=np.asarray(model.test_users, dtype=np.int64), pred_ids=np.asarray(preds, dtype=np.int64),
             pred_scores=np.asarray(scores, dtype=np.float32))
    all_preds, all_scores = model.predict(range(ds.n_users), with_scores=True)
    g.update(pred_all_ids=np.asarray(all_preds, dtype=np.int64), pred_all_scores=np.asarray(all_scores, dtype=np.float32))
    res = model.evaluate()
    for m, v in res.items():
        g["metric_" + m] = np.asarray(v, dtype=np.float64)
    rng = np.random.default_rng(seed + 1)
    batch = make_batch(ds, rng, batch_rows, n_neg)
    g["batch"] = batch.numpy()
    for tag, training in (("eval", False), ("train", True)):
        for k, v in loss_and_grads(model, batch, seed + 7, training).items():
            g[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, {k: getattr(v, "shape", None) for k, v in g.items() if k.startswith(("pred_ids", "norm_val", "rep_user"))})
    return ds, model, args


def adv_case(name, data, ks, seed, batch_rows):
    torch.manual_seed(seed)
    args = args_for("adv_sampling", data, ks, [])
    ds = AdvSamplDataset(args)
    model = AdvSamplModel(args, ds)
    g = dataset_arrays(ds)
    g.update(user_w=model.embedding_user.weight.detach().numpy().copy(),
             item_w=model.embedding_item.weight.detach().numpy().copy(),
             n_layers=np.int64(args.n_layers), dropout=np.float64(args.dropout), reg_lambda=np.float64(args.reg_lambda),
             ks=np.asarray(args.k, dtype=np.int64), pos_samples=np.int64(ds.pos_samples))
    random.seed(seed)
    data_rows = torch.stack([ds[int(i)] for i in np.random.default_rng(seed).integers(len(ds), size=batch_rows)])
    g["data"] = data_rows.numpy()
    # replay python's RNG to record the positives the reference will sample (advanced_sampling.py:64)
    random.seed(seed + 3)
    sampled = []
    for u in data_rows[:, 0].tolist():
        pos = ds.positive_lists[u]["list"]
        sampled.append(random.sample(pos, min(ds.pos_samples, len(pos))))
    width = ds.pos_samples
    sp = np.full((batch_rows, width), -1, dtype=np.int64)
    for b, s in enumerate(sampled):
        sp[b, :len(s)] = s
    g["sampled_pos"] = sp
    captured = {}
    orig = BaseModel.get_loss

    def spy(self, data):
        captured["triples"] = data.detach().clone()
        return orig(self, data)

    BaseModel.get_loss = spy
    try:
        # eval-mode (no dropout) so the two propagations are deterministic
        model.training = False
        random.seed(seed + 3)
        model._loss_values = {"bpr": 0.0, "reg": 0.0}
        model.zero_grad()
        loss = model.get_loss(data_rows)
        loss.backward()
    finally:
        BaseModel.get_loss = orig
    g.update(triples=captured["triples"].numpy(), loss=loss.detach().numpy(),
             grad_user=model.embedding_user.weight.grad.numpy().copy(),
             grad_item=model.embedding_item.weight.grad.numpy().copy())
    with torch.no_grad():
        ue, ie = model.representation
        rankings = model.score_pairwise_adv(ue[data_rows[:, 0]], ie[data_rows[:, 1:]])
    g["rankings"] = rankings.numpy().reshape(batch_rows, -1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, "triples", g["triples"].shape)


def ltr_case(name, data, ks, seed, text_dim, with_pop, ltr_layers, batch_rows):
    torch.manual_seed(seed)
    model_name = "ltr_pop" if with_pop else "ltr_linear"
    extra = ["--ltr_layers", *map(str, ltr_layers)] if ltr_layers else []
    args = args_for(model_name, data, ks, extra)
    # LTRDataset needs review text + SBERT caches (out of scope): feed the model the
    # attributes it copies (ltr_models.py:49-55, :216-219) on top of a real BaseDataset.
    ds = BaseDataset(args)
    gen = torch.Generator().manual_seed(seed + 11)
    ds.items_as_avg_reviews = torch.randn(ds.n_items, text_dim, generator=gen)
    ds.users_as_avg_reviews = torch.randn(ds.n_users, text_dim, generator=gen)
    ds.users_as_avg_desc = torch.randn(ds.n_users, text_dim, generator=gen)
    ds.items_as_desc = torch.randn(ds.n_items, text_dim, generator=gen)
    ds.popularity_users = torch.rand(ds.n_users, 1, generator=gen)
    ds.popularity_items = torch.rand(ds.n_items, 1, generator=gen)
    cls = LTRLinearWPop if with_pop else LTRLinear
    model = cls(args, ds)
    g = dataset_arrays(ds)
    g.update(user_w=model.embedding_user.weight.detach().numpy().copy(),
             item_w=model.embedding_item.weight.detach().numpy().copy(),
             items_rev=ds.items_as_avg_reviews.numpy(), users_rev=ds.users_as_avg_reviews.numpy(),
             users_desc=ds.users_as_avg_desc.numpy(), items_desc=ds.items_as_desc.numpy(),
             pop_users=ds.popularity_users.numpy(), pop_items=ds.popularity_items.numpy(),
             n_layers=np.int64(args.n_layers), dropout=np.float64(args.dropout), reg_lambda=np.float64(args.reg_lambda),
             ks=np.asarray(args.k, dtype=np.int64), n_head_layers=np.int64(len(model.layers)))
    for li, layer in enumerate(model.layers):
        g[f"head_w{li}"] = layer.weight.detach().numpy().copy()
        g[f"head_b{li}"] = layer.bias.detach().numpy().copy()
    model.training = False
    with torch.no_grad():
        ue, ie = model.representation
        users = torch.arange(ds.n_users)
        g["score_batchwise"] = model.score_batchwise(ue[users], ie, users).numpy().copy()
        rng = np.random.default_rng(seed + 5)
        pu = torch.from_numpy(rng.integers(ds.n_users, size=batch_rows))
        pi = torch.from_numpy(rng.integers(ds.n_items, size=batch_rows))
        g["pair_users"], g["pair_items"] = pu.numpy(), pi.numpy()
        g["score_pairwise"] = model.score_pairwise(ue[pu], ie[pi], pu, pi).numpy().copy()
    preds, scores = model.predict(model.test_users, with_scores=True)
    g.update(test_users=np.asarray(model.test_users, dtype=np.int64), pred_ids=np.asarray(preds, dtype=np.int64),
             pred_scores=np.asarray(scores, dtype=np.float32))
    df = pd.DataFrame.from_dict({"user_id": model.test_users, "y_true": model.true_test_lil, "y_pred": preds,
                                 "scores": scores})
    for m, v in calculate_metrics(df, model.metrics, model.k).items():
        g["metric_" + m] = np.asarray(v, dtype=np.float64)
    batch = make_batch(ds, np.random.default_rng(seed + 1), batch_rows, 1)
    g["batch"] = batch.numpy()
    for tag, training in (("eval", False), ("train", True)):
        for k, v in loss_and_grads(model, batch, seed + 7, training).items():
            g[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, "batchwise", g["score_batchwise"].shape)


def main():
    os.makedirs(OUT, exist_ok=True)
    work = tempfile.mkdtemp(prefix="tgcn_golden_")
    os.chdir(work)  # parse_args creates runs/<data>/<uid>/ relative to cwd
    dummy = os.path.join(ref_shims.REFERENCE_ROOT, "data", "dummy")
    lgcn_case("dummy_lgcn", dummy, [2, 3], [], batch_rows=8, n_neg=1, seed=0)

    small = os.path.join(work, "small")
    write_dataset(small, n_users=120, n_items=70, n_train=900, seed=5)
    lgcn_case("small_lgcn_d64", small, [5, 10], ["--emb_size", "64"], batch_rows=96, n_neg=2, seed=1)
    lgcn_case("small_lgcn_d32_single", small, [20], ["--emb_size", "32", "--single", "--n_layers", "2"],
              batch_rows=64, n_neg=1, seed=2)
    lgcn_case("small_lgcn_d128_l4", small, [20, 40], ["--emb_size", "128", "--n_layers", "4"],
              batch_rows=64, n_neg=3, seed=3)
    adv_case("small_adv", small, [5, 10], seed=4, batch_rows=48)
    ltr_case("small_ltr_linear", small, [5, 10], seed=6, text_dim=24, with_pop=False, ltr_layers=[], batch_rows=64)
    ltr_case("small_ltr_pop", small, [5, 10], seed=7, text_dim=24, with_pop=True, ltr_layers=[4], batch_rows=64)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
