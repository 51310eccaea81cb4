"""GPU: the training step replayed from one CUDA graph (textgcn_b200.train_graph) against the eager step — same dropout
draws (device counter vs host counter), same Adam trajectory (bias corrections computed on the device vs on the host)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from helpers import StubDataset, load_weights, params_from_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batches(g, n, batch, seed):
    rng = np.random.default_rng(seed)
    nu, ni = int(g["n_users"]), int(g["n_items"])
    tu, ti = g["train_u"], g["train_i"]
    out = []
    for _ in range(n):
        e = rng.integers(len(tu), size=batch)
        out.append(torch.from_numpy(np.stack([tu[e], ti[e], rng.integers(ni, size=batch)], 1).astype(np.int64)))
    return out


@pytest.mark.parametrize("dropout", [0.0, 0.4])
def test_graphed_step_follows_the_eager_trajectory(dropout):
    from textgcn_b200.models import BaseModel
    from textgcn_b200.optim import FusedAdam
    from textgcn_b200.train_graph import GraphedTrainStep
    g = load_golden("small_lgcn_d64")
    batches = _batches(g, 9, 64, seed=3) + _batches(g, 1, 40, seed=4)  # the last batch is ragged: it runs eagerly
    weights, losses = {}, {}
    for mode in ("eager", "graph"):
        torch.manual_seed(0)
        model = BaseModel(params_from_golden(g, lr=1e-2, dropout=dropout, dropout_rng="device"), StubDataset(g, DEV))
        load_weights(model, g)
        model.train()
        model.training = True
        model._loss_values = {"bpr": 0.0, "reg": 0.0}
        seen = []
        if mode == "eager":
            opt = FusedAdam(model.parameters(), lr=1e-2)
            for data in batches:
                opt.zero_grad(set_to_none=False)
                loss = model.get_loss(data)
                loss.backward()
                opt.step()
                seen.append(float(loss))
        else:
            opt = FusedAdam(model.parameters(), lr=1e-2, capturable=True)
            step = GraphedTrainStep(model, opt)
            for data in batches:
                seen.append(float(step(data)))
            assert step.graph is not None and step.shape == (64, 3)
            assert abs(float(step.loss_sums.sum()) - sum(seen)) < 1e-4 * abs(sum(seen)) + 1e-6
        weights[mode] = (model.embedding_user.weight.detach().cpu().numpy().copy(),
                         model.embedding_item.weight.detach().cpu().numpy().copy())
        losses[mode] = seen
    assert np.allclose(losses["graph"], losses["eager"], rtol=1e-4, atol=1e-7), (losses["graph"], losses["eager"])
    assert rel_err(weights["graph"][0], weights["eager"][0]) < 1e-4
    assert rel_err(weights["graph"][1], weights["eager"][1]) < 1e-4


def test_fit_with_cuda_graph_reduces_loss():
    from textgcn_b200.models import BaseModel
    g = load_golden("small_lgcn_d64")
    torch.manual_seed(0)
    batches = [torch.from_numpy(g["batch"])] * 8
    model = BaseModel(params_from_golden(g, epochs=3, evaluate_every=3, lr=1e-2, dropout=0.0, cuda_graph=True, dropout_rng="device"),
                      StubDataset(g, DEV))
    load_weights(model, g)
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    before = float(model.get_loss(batches[0]))
    model.fit(batches)
    model.training = False
    model._loss_values = {"bpr": 0.0, "reg": 0.0}
    assert float(model.get_loss(batches[0])) < before
