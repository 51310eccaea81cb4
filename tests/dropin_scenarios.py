"""Scenarios shared by tests/test_gpu_dropin.py (real kernels on cuda:0) and tests/test_dropin_emulated.py (the same
host logic on the CPU with the kernels emulated, tests/emul.py): the ``textgcn_b200.dropin`` mixins in front of the
UNMODIFIED reference classes against the same reference classes on the CPU."""
import logging
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import TOL, rel_err
from oracle import lightgcn_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shims  # noqa: E402

needs_reference = pytest.mark.skipif(ref_shims.reference_root() is None, reason="reference neither mounted nor staged in baseline/_ref")


def _parse(parse_args, argv):
    """parse_args overwrites CUDA_VISIBLE_DEVICES (parser.py:173); restore it so later (multi-process) tests see the GPUs."""
    saved = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        return parse_args(argv)
    finally:
        if saved is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved


def make_env(tmp_path_factory, dev):
    """Generator behind the module fixtures: yields the shared environment for device ``dev`` ("cuda" | "cpu" emulated)."""
    if dev == "cuda":
        torch.cuda.init()  # before parse_args can hide the devices
    ref_shims.install()
    import TextGCN
    from TextGCN.parser import parse_args
    import make_golden as MG  # the seeded TSV writer + batch maker used for the golden fixtures
    from textgcn_b200.dropin import make_dropin_classes
    work = tmp_path_factory.mktemp("dropin")
    cwd = os.getcwd()
    os.chdir(work)  # parse_args creates runs/<data>/<uid>/ relative to the cwd
    data = os.path.join(work, "small")
    MG.write_dataset(data, n_users=150, n_items=90, n_train=1300, seed=11)

    def args(model, gpu, uid, extra=()):
        a = _parse(parse_args, ["--model", model, "-d", data, "-k", "5", "10", "--gpu", "0" if (gpu and dev == "cuda") else "", "--quiet", "--slurm",
                                "--uid", uid, "--epochs", "2", "--evaluate_every", "1", "--batch_size", "64", *extra])
        assert a.device.type == (dev if gpu else "cpu")
        return a

    yield dict(T=TextGCN, MG=MG, cls=make_dropin_classes(TextGCN), args=args, work=str(work), dev=dev)
    os.chdir(cwd)
    logging.getLogger().handlers.clear()


def _pair(env, model_name, ds_cls, extra=(), uid="m", prep=None):
    """(reference model on the CPU, drop-in model on the GPU) with identical parameters."""
    T = env["T"]
    a_cpu, a_gpu = env["args"](model_name, False, uid + "_cpu", extra), env["args"](model_name, True, uid + "_gpu", extra)
    ds_cpu, ds_gpu = ds_cls(a_cpu), ds_cls(a_gpu)
    if prep:
        prep(ds_cpu, "cpu")
        prep(ds_gpu, env["dev"])
    ref_cls = {"lgcn": T.BaseModel, "adv_sampling": T.AdvSamplModel, "ltr_linear": T.LTRLinear, "ltr_pop": T.LTRLinearWPop}[model_name]
    torch.manual_seed(3)
    ref = ref_cls(a_cpu, ds_cpu)
    ours = env["cls"][model_name](a_gpu, ds_gpu)
    ours.load_state_dict(ref.state_dict())
    return ref, ours, ds_cpu, ds_gpu, a_cpu, a_gpu


def _same_lists(ids_a, sc_a, ids_b, sc_b):
    st = O.topk_lists_equivalent(np.asarray(ids_a, dtype=np.int64), np.asarray(sc_a, dtype=np.float64),
                                 np.asarray(ids_b, dtype=np.int64), np.asarray(sc_b, dtype=np.float64), atol=2e-4)  # scores are rounded to 4 d.p.
    assert st["bad"] == 0, st
    return st


def _step(model, batch, seed, training=True):
    model.zero_grad()
    model.training = training
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    torch.manual_seed(seed)
    random.seed(seed)
    loss = model.get_loss(batch)
    loss.backward()
    grads = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
    model.training = False
    return float(loss), grads


def scenario_lgcn(env):
    T, MG = env["T"], env["MG"]
    ref, ours, ds_cpu, ds_gpu, a_cpu, a_gpu = _pair(env, "lgcn", T.BaseDataset, uid="lgcn")
    from textgcn_b200.models import B200HotPath
    assert type(ours).fit is T.BaseModel.fit and type(ours).checkpoint is T.BaseModel.checkpoint   # the shell is the reference's
    assert type(ours).representation is B200HotPath.representation
    # representation / layer_aggregation / score_batchwise / score_pairwise by their reference signatures
    with torch.no_grad():
        ue_r, ie_r = ref.representation
        ue_o, ie_o = ours.representation
        assert rel_err(ue_o.cpu(), ue_r) < TOL and rel_err(ie_o.cpu(), ie_r) < TOL
        agg = ours.layer_aggregation(ours.norm_matrix, ours.embedding_matrix)
        assert rel_err(agg.cpu(), ref.layer_aggregation(ref.norm_matrix, ref.embedding_matrix)) < TOL
        users = np.arange(0, ds_cpu.n_users, 3)
        sb = ours.score_batchwise(ue_o[users], ie_o, users)
        assert sb.shape == (len(users), ds_cpu.n_items)
        assert rel_err(sb.cpu(), ref.score_batchwise(ue_r[users], ie_r, users)) < TOL
        sp = ours.score_pairwise(ue_o[:40], ie_o[:40], None, None)
        assert rel_err(sp.cpu(), ref.score_pairwise(ue_r[:40], ie_r[:40], None, None)) < TOL
    # one training step with the same dropout draw (torch.rand on the CPU generator, base_model.py:82)
    batch = MG.make_batch(ds_cpu, np.random.default_rng(5), 64, 1)
    for training in (False, True):
        l_r, g_r = _step(ref, batch, 17, training)
        l_o, g_o = _step(ours, batch, 17, training)
        assert abs(l_o - l_r) <= TOL * abs(l_r)
        for name in g_r:
            assert rel_err(g_o[name], g_r[name]) < TOL, (training, name)
    # fit: two epochs through the REFERENCE's loop (torch Adam, per-step isnan sync, evaluate + checkpoint every epoch)
    batches = [MG.make_batch(ds_cpu, np.random.default_rng(100 + i), 64, 1) for i in range(6)]
    for m in (ref, ours):
        torch.manual_seed(7)
        m.fit(batches)
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    for name in sd_r:
        assert torch.allclose(sd_o[name].cpu(), sd_r[name], rtol=0, atol=2e-6), name
    assert ours.metrics_logger["recall"].shape == (2, 2)
    for name in ref.metrics:
        assert np.allclose(ours.metrics_logger[name], ref.metrics_logger[name], rtol=0, atol=1e-9), name
    # predict (lists of lists) and evaluate
    p_r, s_r = ref.predict(ref.test_users, with_scores=True)
    p_o, s_o = ours.predict(ours.test_users, with_scores=True)
    _same_lists(p_o, s_o, p_r, s_r)
    assert isinstance(p_o, list) and isinstance(p_o[0], list) and len(p_o[0]) == 10
    res_r, res_o = ref.evaluate(), ours.evaluate()
    for name in res_r:
        assert np.allclose(res_o[name], res_r[name], rtol=0, atol=1e-9), name
    # predictions.tsv through predict(save=True) (base_model.py:268-273)
    ours.predict(range(ds_gpu.n_users), save=True)
    assert os.path.exists(os.path.join(ours.save_path, "predictions.tsv"))
    # checkpoints interchange in both directions (state-dict keys and file names are the reference's)
    for src, dst_name, gpu in ((ours, "lgcn_load_cpu", False), (ref, "lgcn_load_gpu", True)):
        assert os.path.exists(os.path.join(src.save_path, "best.pkl"))
        a = env["args"]("lgcn", gpu, dst_name, ["--load", src.save_path])
        cls = env["cls"]["lgcn"] if gpu else T.BaseModel
        loaded = cls(a, ds_gpu if gpu else ds_cpu)   # load_model evaluates immediately (base_model.py:278-289)
        best = torch.load(os.path.join(src.save_path, "best.pkl"), map_location="cpu")
        for name, v in loaded.state_dict().items():
            assert torch.equal(v.cpu(), best[name]), name
        assert loaded.metrics_logger["recall"].shape == (0, 2)


def _ltr_prep(ds, device, dim=24):
    gen = torch.Generator().manual_seed(23)
    for name, n in (("items_as_avg_reviews", ds.n_items), ("users_as_avg_reviews", ds.n_users), ("users_as_avg_desc", ds.n_users),
                    ("items_as_desc", ds.n_items)):
        setattr(ds, name, torch.randn(n, dim, generator=gen).to(device))
    ds.popularity_users = torch.rand(ds.n_users, 1, generator=gen).to(device)
    ds.popularity_items = torch.rand(ds.n_items, 1, generator=gen).to(device)


def scenario_ltr(env, model_name, monkeypatch):
    T, MG = env["T"], env["MG"]
    from textgcn_b200.models import B200HotPath, B200LTR
    # a trained base LightGCN to load (the reference's documented LTR workflow: --load_base <run> --freeze)
    base_ref, base_ours, *_ = _pair(env, "lgcn", T.BaseDataset, uid="base_" + model_name)
    base_metrics = base_ours.evaluate()
    base_ours.checkpoint(1)   # the reference's checkpoint(): latest_checkpoint.pkl + best.pkl
    seen = []
    orig = B200HotPath.evaluate
    monkeypatch.setattr(B200HotPath, "evaluate", lambda self, *a, **k: seen.append(orig(self, *a, **k)) or seen[-1])
    extra = ["--load_base", base_ours.save_path, "--freeze"] + (["--ltr_layers", "4"] if model_name == "ltr_pop" else [])
    ref, ours, ds_cpu, ds_gpu, a_cpu, a_gpu = _pair(env, model_name, T.BaseDataset, extra=extra, uid=model_name, prep=_ltr_prep)
    # G18: the base model was loaded and evaluated inside _add_vars, BEFORE the head existed, with plain LightGCN scoring
    assert len(seen) == 1
    for name in base_metrics:
        assert np.allclose(seen[0][name], base_metrics[name], rtol=0, atol=1e-12), name
    assert ours.__dict__["score_batchwise"].__func__ is B200LTR.score_batchwise_ltr       # the re-binding landed on the mixin
    assert ours.__dict__["evaluate"].__func__ is B200LTR.evaluate_ltr
    assert not ours.embedding_user.weight.requires_grad
    with torch.no_grad():
        ue_r, ie_r = ref.representation
        ue_o, ie_o = ours.representation
        users = np.arange(0, ds_cpu.n_users, 2)
        # a17-a21 by their reference signatures
        uv, iv = ours.get_user_vectors(ue_o[users], users), ours.get_item_vectors(ie_o, ours.all_items)
        assert set(uv) == {"emb", "desc", "reviews"} and iv["desc"].shape == (ds_cpu.n_items, 24)
        fb = ours.get_features_batchwise(uv, iv)
        fb_r = ref.get_features_batchwise(ref.get_user_vectors(ue_r[users], users), ref.get_item_vectors(ie_r, ref.all_items))
        assert fb.shape == fb_r.shape == (len(users), ds_cpu.n_items, 5)
        for f in range(5):
            assert rel_err(fb[..., f].cpu(), fb_r[..., f]) < TOL, f
        sb = ours.score_batchwise(ue_o[users], ie_o, users)           # = score_batchwise_ltr (instance re-binding)
        assert rel_err(sb.cpu(), ref.score_batchwise(ue_r[users], ie_r, users)) < TOL
        pu = torch.from_numpy(np.random.default_rng(2).integers(ds_cpu.n_users, size=50))
        pi = torch.from_numpy(np.random.default_rng(3).integers(ds_cpu.n_items, size=50))
        sp = ours.score_pairwise(ue_o[pu.to(env["dev"])], ie_o[pi.to(env["dev"])], pu.to(env["dev"]), pi.to(env["dev"]))
        sp_r = ref.score_pairwise(ue_r[pu], ie_r[pi], pu, pi)
        assert sp.shape == sp_r.shape == (50, 1) and rel_err(sp.cpu(), sp_r) < TOL
        fp = ours.get_features_pairwise(ours.get_user_vectors(ue_o[pu.to(env["dev"])], pu.to(env["dev"])), ours.get_item_vectors(ie_o[pi.to(env["dev"])], pi.to(env["dev"])))
        fp_r = ref.get_features_pairwise(ref.get_user_vectors(ue_r[pu], pu), ref.get_item_vectors(ie_r[pi], pi))
        assert rel_err(fp.cpu(), fp_r) < TOL
    batch = MG.make_batch(ds_cpu, np.random.default_rng(9), 64, 1)
    for training in (False, True):
        l_r, g_r = _step(ref, batch, 21, training)
        l_o, g_o = _step(ours, batch, 21, training)
        assert abs(l_o - l_r) <= TOL * max(abs(l_r), 1e-3)
        assert set(g_o) == set(g_r) and all(n.startswith("layers.") for n in g_r)
        for name in g_r:
            assert rel_err(g_o[name], g_r[name]) < 5 * TOL, (training, name)
    p_r, s_r = ref.predict(ref.test_users, with_scores=True)
    p_o, s_o = ours.predict(ours.test_users, with_scores=True)
    _same_lists(p_o, s_o, p_r, s_r)
    batches = [MG.make_batch(ds_cpu, np.random.default_rng(200 + i), 64, 1) for i in range(4)]
    for m in (ref, ours):
        torch.manual_seed(7)
        m.fit(batches)                     # the reference's loop; evaluate -> evaluate_ltr (feature-weight logging) -> fused predict
    for name, v in ref.state_dict().items():
        assert torch.allclose(ours.state_dict()[name].cpu(), v, rtol=0, atol=5e-6), name
    assert set(ours.state_dict()) == set(ref.state_dict())
    for name in ref.metrics:
        assert np.allclose(ours.metrics_logger[name], ref.metrics_logger[name], rtol=0, atol=1e-9), name


def scenario_adv(env):
    T = env["T"]
    ref, ours, ds_cpu, ds_gpu, a_cpu, a_gpu = _pair(env, "adv_sampling", T.AdvSamplDataset, uid="adv")
    random.seed(5)
    rows = torch.stack([ds_cpu[int(i)] for i in np.random.default_rng(5).integers(len(ds_cpu), size=48)])   # (48, 1 + 90)
    with torch.no_grad():
        ue_r, ie_r = ref.representation
        ue_o, ie_o = ours.representation
        users, items = rows[:, 0], rows[:, 1:]
        adv = ours.score_pairwise_adv(ue_o[users.to(env["dev"])], ie_o[items.to(env["dev"])])
        assert adv.shape == (48, items.shape[1])
        assert rel_err(adv.cpu(), ref.score_pairwise_adv(ue_r[users], ie_r[items])) < TOL
        one = ours.score_pairwise_adv(ue_o[users[:1].to(env["dev"])], ie_o[items[:1].to(env["dev"])])
        assert one.shape == (1, items.shape[1])                       # no squeeze to 1-D (G14)
    for training in (False, True):
        # the same Python-RNG state drives random.sample(positives) in both (advanced_sampling.py:64), the same torch seed
        # drives both dropout draws of the step (G12)
        l_r, g_r = _step(ref, rows, 31, training)
        l_o, g_o = _step(ours, rows, 31, training)
        assert abs(l_o - l_r) <= TOL * abs(l_r), (training, l_o, l_r)
        for name in g_r:
            assert rel_err(g_o[name], g_r[name]) < TOL, (training, name)
