"""One invocation of every hot-path kernel at the Electronics shape (c2), for `ncu --set full`.

    ncu --set full --clock-control none --import-source on -k regex:'spmm|bpr_|adam_|dropout_|adv_select|ltr_|tf32_split|
        eval_topk|topk_merge|dense_nt|pairwise_adv|topk_metrics|layer_mean|sample_' -o gpurun_out/r02_kernels python tools/profile_kernels.py
Every op runs exactly once (cold), so each kernel shows up once per launch in the report."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_workload, make_batch  # noqa: E402
from textgcn_b200 import metrics as M  # noqa: E402
from textgcn_b200 import ops  # noqa: E402
from textgcn_b200.models import _FusedBprFn  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    w = build_workload(os.environ.get("PROFILE_WORKLOAD", "c2"), dev)
    nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
    g = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
    gen = torch.Generator(device=dev).manual_seed(0)
    # a2-a6: propagate, eval mode then train mode (keep mask) + the Horner backward through the transposed mask (tperm)
    emb = ops.propagate_fwd(g, w["uw"], w["iw"], L)
    keep = ops.dropout_mask(nnz, 0.4, 1234, dev)                                 # n2: dropout_mask_kernel
    emb_t = ops.propagate_fwd(g, w["uw"], w["iw"], L, keep=keep, dropout=0.4)
    grad = ops.propagate_bwd(g, emb_t, L, keep=keep, dropout=0.4)
    # a7-a10: fused BPR + grads, fused Adam
    batch = make_batch(w, 2048, dev, torch)
    uw, iw = w["uw"].clone().requires_grad_(True), w["iw"].clone().requires_grad_(True)
    losses = _FusedBprFn.apply(uw, iw, g, L, False, keep, 0.4, ops.as_index(batch[:, 0], dev), ops.as_index(batch[:, 1], dev),
                               ops.as_index(batch[:, 2:].t(), dev), 1e-4)
    losses.sum().backward()
    m, v = torch.zeros_like(uw), torch.zeros_like(uw)
    ops.adam_step(uw.detach(), uw.grad, m, v, 1e-3, 0.9, 0.999, 1e-8, 1)
    # a11-a13: fused eval (3xTF32 + exact fp32 on a slice), dense score matrix, metrics
    users = torch.arange(nu, dtype=torch.int32, device=dev)
    ids, sc = ops.eval_topk(g, emb[:nu], emb[nu:], 20, users=users)
    ops.eval_topk(g, emb[:nu], emb[nu:], 20, users=users[:8192].contiguous(), precision="fp32")
    ops.score_batchwise(emb[:2048], emb[nu:])
    truth = M.TruthCSR.from_pairs(torch.arange(nu, device=dev), torch.randint(0, ni, (nu,), generator=gen, device=dev), nu)
    ops.topk_metrics(ids, truth.ptr, truth.ids, [20])
    # a14-a16: samplers + hardest-negative selection
    from textgcn_b200.sampler import AdvEpochSampler, BprEpochSampler
    BprEpochSampler(g, 2048).sample(users[:2048].contiguous(), 7)
    data = AdvEpochSampler(g, 2048).sample(users[:2048].contiguous(), 7)
    ops.adv_select(g, emb, ops.as_index(data[:, 0], dev), ops.as_index(data[:, 1:], dev), 20)
    ops.score_pairwise_adv(emb[:256], emb[nu:][data[:256, 1:]])
    # a17-a21: LTR pairwise features, packing, streamed wide contraction
    D = 768
    ur, ud = torch.randn(nu, D, generator=gen, device=dev), torch.randn(nu, D, generator=gen, device=dev)
    ir, idesc = torch.randn(ni, D, generator=gen, device=dev), torch.randn(ni, D, generator=gen, device=dev)
    pu, pi = ops.as_index(batch[:, 0], dev), ops.as_index(batch[:, 1], dev)
    ops.ltr_pairwise_features(nu, emb, pu, pi, ur, ud, ir, idesc)
    items_p = ops.ltr_pack_items(emb[nu:], ir, idesc, [0.3, 0.2, 0.1, 0.05, 0.07])
    ru = users[:18944].contiguous()
    users_p = ops.ltr_pack_users(ru, emb[:nu], ur, ud)
    ops.eval_topk(g, users_p, items_p, 20, users=ru, by_position=True,
                  user_bias=torch.zeros(ru.numel(), device=dev), item_bias=torch.zeros(ni, device=dev))
    ops.ltr_features_rows(users_p[:2048, :d], items_p[:2048, :d], users_p[:2048, d:d + D], users_p[:2048, d + D:], items_p[:2048, d:d + D],
                          items_p[:2048, d + D:])
    torch.cuda.synchronize()
    print("profiled ops ran")


if __name__ == "__main__":
    main()
