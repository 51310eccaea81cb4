"""Launch list of ONE adv_sampling training step at c2 (use under: ncu --profile-from-start off --metrics gpu__time_duration.sum)."""
import logging
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from textgcn_b200 import ops  # noqa: E402
from textgcn_b200.models import AdvSamplModel, make_params  # noqa: E402
from textgcn_b200.optim import FusedAdam  # noqa: E402
from textgcn_b200.sampler import AdvEpochSampler  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    w = bench.build_workload("c2", dev)
    graph = ops.Graph(w["nu"], w["ni"], w["rowptr"], w["col"], w["val"])

    class DS:
        pass

    ds = DS()
    ds.n_users, ds.n_items, ds.graph, ds.norm_matrix = w["nu"], w["ni"], graph, None
    ds.test_users, ds.true_test_lil = [0], [[0]]
    batch = 2048
    params = make_params(emb_size=w["d"], n_layers=w["L"], k=[20], batch_size=batch, fused_adam=True, dropout_rng="device",
                         positive_sampler="device", device=dev, logger=logging.getLogger("probe"))
    model = AdvSamplModel(params, ds)
    opt = FusedAdam(model.parameters(), lr=params.lr)
    smp = AdvEpochSampler(graph, batch_size=batch, seed=0)
    users = torch.arange(batch, dtype=torch.int32, device=dev)
    model.train()
    model.training = True

    def step():
        data = smp.sample(users, 1234)
        opt.zero_grad(set_to_none=False)
        triples = model.select_triples(data)
        loss = super(AdvSamplModel, model).get_loss(triples)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    print("step ms", e0.elapsed_time(e1))
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
