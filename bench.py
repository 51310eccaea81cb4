#!/usr/bin/env python
"""Benchmark of the LightGCN hot path (BASELINE.json metric: propagation edges/s + eval users/s, top-k@20).

    python bench.py --gpus N --steps K --warmup W            # ours (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own BaseModel on the host cores, rank 0 only

Workload.  BASELINE.json quotes the metric as a 1/2/4/8-GPU series, and the only config that names that series is
configs[4] — the 200M-edge graph (10M users, 2M items, emb 128, 4 layers), which also fits one GPU — so EVERY N runs
"c5" at the top level (strong scaling; the driver's per-N ratios are like-for-like).  The N = 1 line additionally
carries `c2`: BASELINE.json configs[1] (Electronics-shaped, 190k users, 63k items, 1.7M edges, emb 64, 3 layers) with its
own propagation / eval / e2e / training-step / adv_sampling / LTR numbers and CPU baseline.  `--workload c2` puts c2
at the top level instead.

A "step" is one pass of the hot path over the whole graph: ``representation`` = L fused SpMM layers + layer mean.
``value`` = nnz(Â)·L / t with inputs resident in HBM; ``e2e`` = the same through the host-buffer C-ABI call
(``tgcn_propagate_host``: H2D of E0, L layers, D2H of the result).
  N > 1 : default `--mg-scheme grid`: G feature slices x R user partitions (1xN below 8 GPUs, 2x4 at 8), one NCCL
          all-reduce of the item-table slice per hop inside each row group, the result exchange fused into the last
          passes as peer-memory stores.  `--mg-scheme bipartite` / `rowblock` are the earlier schemes.
          Eval: user-range sharding (comm-free) and the item-range variant with a cross-GPU top-k merge.
Between timed iterations L2 is flushed (a 256 MiB write); timing is CUDA events on the launching stream, max over
ranks.  The JSON line also carries ``roofline`` (dominant kernel: spmm_group_kernel, HBM bound), ``cpu_baseline``
(the reference's own BaseModel on the host cores, N = 1 only), ``torch_cuda_reference`` (the reference's torch.sparse / matmul / topk ops on
the same GPU), ``eval`` (users/s) and ``clocks``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# c5 eval leg: 8 x 148 SMs x 128-user tiles = 151 552 of the 10M users, the same at every N (whole waves of CTAs at N = 1..8)
EVAL_USERS_C5 = 8 * 18944
L2_CAP_BYTES_PER_CLK = 6300.0
TF32_DENSE_PEAK_TFLOPS = 1125.0  # B200 nominal dense TF32 (half of the 2.25 PFLOP/s bf16 figure); MEASURED_PEAKS.json has no TF32 entry

METRIC = "propagation edges/s (directed nnz x layers per second; eval users/s top-k@20 in `eval`)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "c5", "tiny"])
    ap.add_argument("--eval-users", type=int, default=0, help="users ranked in the eval leg (0 = workload default)")
    ap.add_argument("--eval-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="N = 1: skip the extra Electronics-shaped (configs[1]) legs")
    ap.add_argument("--no-torch-ref", action="store_true", help="skip the torch.sparse-on-CUDA comparator")
    ap.add_argument("--no-extras", action="store_true", help="skip the adv_sampling / LTR legs (BASELINE.json configs[2], [3])")
    ap.add_argument("--topk", type=int, default=20)
    ap.add_argument("--ar-chunks", type=int, default=1, help="bipartite scheme: split the item-table all-reduce into this many chunks")
    ap.add_argument("--mg-scheme", default="grid", choices=["grid", "bipartite", "rowblock"],
                    help="multi-GPU propagation: G feature slices x R user partitions with a peer-memory result exchange (grid), "
                         "users partitioned + item-table all-reduce (bipartite = grid 1xN without the final exchange), or row "
                         "blocks + all-gather of layer embeddings (rowblock)")
    ap.add_argument("--grid", default="auto", help="grid scheme shape(s) GxR[:hp][,GxR...] (G·R = N; the first is the headline; "
                    ":hp = row-group all-reduce on a high-priority stream); "
                    "auto = 1x2, 1x4, 2x4 for N = 2, 4, 8 (the fastest measured)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "collective"], help="grid scheme: result exchange through peer "
                    "memory (fused into the last passes) or NCCL all-to-all / all-gather")
    ap.add_argument("--no-nccl-high-priority", dest="nccl_high_priority", action="store_false",
                    help="grid scheme: row-group all-reduce on a normal-priority stream (default: high priority, measured 18.2 vs "
                         "18.8 ms per step at 2x4 on 8 GPUs)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def spmm_layer_bytes(nnz, n, d):
    """Algorithmic bytes of one SpMM layer: gathered rows + col/val + output rows + rowptr."""
    return nnz * (4 * d + 8) + n * 4 * d + (n + 1) * 4


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------------
def build_workload(name, device):
    import torch
    from textgcn_b200.graph import norm_adj_csr
    from textgcn_b200.synthetic import WORKLOADS, interactions
    nu, ni, ne, d, L = WORKLOADS[name]
    tu, ti = interactions(nu, ni, ne, device, seed=0)
    rowptr, col, val = norm_adj_csr(tu, ti, nu, ni)
    del tu, ti
    torch.manual_seed(0)
    gen = torch.Generator(device=device).manual_seed(0)
    uw = torch.randn(nu, d, generator=gen, device=device) * 0.1  # base_model.py:68-69
    iw = torch.randn(ni, d, generator=gen, device=device) * 0.1
    return dict(name=name, nu=nu, ni=ni, ne=ne, d=d, L=L, rowptr=rowptr.contiguous(), col=col.contiguous(),
                val=val.contiguous(), uw=uw, iw=iw, nnz=int(col.numel()))


def timed_steps(fn, steps, warmup, flush, torch):
    """W untimed + K timed calls; L2 flushed (untimed) before every timed call; per-step CUDA events."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s, e in ev:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return [s.elapsed_time(e) for s, e in ev]


def make_batch(w, batch, device, torch, seed=1):
    """(B, 3) int64 rows [user, positive, negative]: a random train interaction per row + a uniform random item."""
    gen = torch.Generator(device=device).manual_seed(seed)
    e = torch.randint(0, int(w["rowptr"][w["nu"]]), (batch,), generator=gen, device=device)
    users = torch.searchsorted(w["rowptr"][:w["nu"] + 1].to(torch.int64), e, right=True) - 1
    pos = w["col"][e].to(torch.int64) - w["nu"]
    neg = torch.randint(0, w["ni"], (batch,), generator=gen, device=device)
    return torch.stack([users, pos, neg], dim=1)


def train_leg(w, graph, dev, flush, torch, batch=2048, steps=10, warmup=3):
    import logging
    from textgcn_b200.models import BaseModel, make_params

    class DS:
        pass

    ds = DS()
    ds.n_users, ds.n_items, ds.graph, ds.norm_matrix = w["nu"], w["ni"], graph, None
    ds.test_users, ds.true_test_lil = [0], [[0]]
    params = make_params(emb_size=w["d"], n_layers=w["L"], k=[20], batch_size=batch, fused_adam=True, dropout_rng="device",
                         device=dev, logger=logging.getLogger("bench"))
    model = BaseModel(params, ds)
    from textgcn_b200.optim import FusedAdam
    opt = FusedAdam(model.parameters(), lr=params.lr)
    data = make_batch(w, batch, dev, torch)
    model.train()
    model.training = True

    def step():
        opt.zero_grad(set_to_none=False)
        loss = model.get_loss(data)
        loss.backward()
        opt.step()

    t = timed_steps(step, steps, warmup, flush, torch)
    eager_ms = sum(t) / len(t)
    res = {"eager_ms_per_step": eager_ms}
    try:  # the same step replayed from one CUDA graph (device-side draw counter and Adam step count)
        from textgcn_b200.train_graph import GraphedTrainStep
        gopt = FusedAdam(model.parameters(), lr=params.lr, capturable=True)
        gstep = GraphedTrainStep(model, gopt)
        for _ in range(5):
            gstep(data)
        assert gstep.graph is not None
        t = timed_steps(lambda: gstep(data), steps, warmup, flush, torch)
        res["ms_per_step"] = sum(t) / len(t)
        res["mode"] = "one CUDA graph per step (GraphedTrainStep)"
    except Exception as exc:
        res["ms_per_step"] = eager_ms
        res["mode"] = "eager"
        res["graph_error"] = str(exc)[:200]
    # the same eager step with the reference's default mask source: torch.rand(nnz) on the HOST generator + H2D every step
    # (base_model.py:82; dropout_rng="host" reproduces the reference's draws bit for bit, at this price)
    try:
        model.dropout_rng = "host"
        t = timed_steps(step, 3, 1, flush, torch)
        res["host_rng_eager_ms_per_step"] = sum(t) / len(t)
    except Exception as exc:
        res["host_rng_error"] = str(exc)[:200]
    model.dropout_rng = "device"
    nnz, n, d, L, keep_p = w["nnz"], w["nu"] + w["ni"], w["d"], w["L"], 1.0 - params.dropout
    spmm_pass = nnz * 9 + keep_p * nnz * 4 * d + n * 4 * d            # col/val + mask byte + gathers of the kept edges + output rows
    step_bytes = (nnz                                                  # dropout_mask_kernel
                  + L * spmm_pass + L * n * 4 * d                      # forward passes + the layer-mean epilogue's addend reads
                  + 2 * n * 4 * d                                      # zero fill of the two gradient tables
                  + batch * 3 * 2 * 4 * d * 2                          # bpr_kernel: 3 rows x (emb + E0) read, the same in atomics
                  + 6 * nnz                                            # permute_mask_kernel
                  + L * (spmm_pass + n * 4 * d)                        # Horner backward: each pass also reads G as addend
                  + 7 * n * 4 * d)                                     # adam_kernel over both tables
    ach = step_bytes / (res["ms_per_step"] * 1e-3) / 1e9
    big = n * 4 * d > (256 << 20)
    res["roofline"] = {"bound": "hbm" if big else "l2 / hbm (tables of 65 MB: gathers hit L2, the gradient / Adam streams do not)",
                       "algorithmic_bytes": step_bytes,
                       "achieved": ach, "unit": "GB/s", "frac_of_hbm_peak": ach / peaks()[0],
                       "frac_of_l2_cap": ach / (L2_CAP_BYTES_PER_CLK * 1965e6 / 1e9),
                       "note": "sum of the per-kernel byte models of profiles/r02/kernel_table_c2.md over one step / its CUDA-event time"}
    res["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
    res.update({"batch": batch, "dropout": params.dropout, "steps_per_s": 1e3 / res["ms_per_step"],
                "includes": "device dropout draw, L-layer propagate, fused BPR(SELU)+L2 kernel, Horner backward (L transposed SpMM), "
                            "fused Adam over both tables"})
    return res


def torch_cuda_reference(w, dev, flush, torch, topk, n_eval=8192):
    """The reference's own ops on the SAME GPU: torch.sparse.mm x L + stack/mean (base_model.py:93-106, :141-157), and
    predict as matmul -> index_put(-inf) -> topk in batches of 2048 users (:235-261).  The real comparator of the kernels."""
    nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
    n = nu + ni
    rowptr = w["rowptr"].to(dev)
    counts = (rowptr[1:] - rowptr[:-1]).to(torch.int64)
    row = torch.repeat_interleave(torch.arange(n, device=dev), counts)
    col = w["col"].to(dev)
    norm = torch.sparse_coo_tensor(torch.stack([row, col.to(torch.int64)]), w["val"].to(dev), (n, n)).coalesce()
    del row

    def representation():
        cur = torch.cat([w["uw"], w["iw"]])
        layers = [cur]
        for _ in range(L):
            cur = torch.sparse.mm(norm, cur)
            layers.append(cur)
        return torch.mean(torch.stack(layers), dim=0)

    t = timed_steps(representation, 3, 1, flush, torch)
    ms = sum(t) / len(t)
    emb = representation()
    ue, ie = emb[:nu], emb[nu:]
    n_eval = min(n_eval, nu)
    kept = []

    def predict():
        kept.clear()
        for s0 in range(0, n_eval, 2048):
            users = torch.arange(s0, min(s0 + 2048, n_eval), device=dev)
            scores = ue[users] @ ie.T
            lo, hi = int(rowptr[s0]), int(rowptr[min(s0 + 2048, n_eval)])
            rows = torch.repeat_interleave(torch.arange(users.numel(), device=dev), counts[s0:s0 + users.numel()])
            scores[rows, col[lo:hi].to(torch.int64) - nu] = float("-inf")
            kept.append(torch.topk(scores, topk, dim=1))

    te = timed_steps(predict, 2, 1, flush, torch)
    ems = sum(te) / len(te)
    res = {"representation_ms": ms, "edges_per_s": nnz * L / (ms * 1e-3), "eval_users_per_s": n_eval / (ems * 1e-3),
           "n_users_ranked": n_eval, "ops": "torch.sparse.mm (cuSPARSE) x L + stack/mean; matmul (cuBLAS) + index_put + topk, device-side mask"}
    return res, emb, torch.cat([t.indices for t in kept]), torch.cat([t.values for t in kept])


def metrics_leg(ops, ids, n_users, n_items, k, dev, flush, torch, sample=20000):
    """SURVEY.md §8f n4 at the scale it was asked for: recall / precision / hit / ndcg / f1 @k for ALL n_users of the config in one
    pass of tgcn_topk_metrics over a (n_users, k) id table and a CSR of held-out items (one per user, §8d) — no Python loop over
    users.  The table tiles the ranked sample; a 20 000-row sample is checked against a closed-form torch evaluation (one true item
    per row: recall = hit, precision = hit / k, ndcg = hit / log2(rank + 2), f1 from those)."""
    from textgcn_b200 import metrics as M
    reps = -(-n_users // ids.shape[0])
    table = ids.repeat(reps, 1)[:n_users].contiguous()
    gen = torch.Generator(device=dev).manual_seed(5)
    # make ~half of the rows hits: the true item is one of the row's predictions at a random rank, or a random item
    rank_pos = torch.randint(0, k, (n_users,), generator=gen, device=dev)
    truth = torch.where(torch.rand(n_users, generator=gen, device=dev) < 0.5, table[torch.arange(n_users, device=dev), rank_pos].long(),
                        torch.randint(0, n_items, (n_users,), generator=gen, device=dev))
    csr = M.TruthCSR.from_pairs(torch.arange(n_users, device=dev), truth, n_users)
    t = timed_steps(lambda: ops.topk_metrics(table, csr.ptr, csr.ids, [k]), 3, 1, flush, torch)
    ms = sum(t) / len(t)
    vals = ops.topk_metrics(table[:sample].contiguous(), csr.ptr[:sample + 1].contiguous(), csr.ids[:sample].contiguous(), [k])[0]
    eq = table[:sample].long() == truth[:sample, None]
    hit = eq.any(1).double()
    pos = torch.where(eq.any(1), eq.double().argmax(1), torch.zeros(sample, dtype=torch.int64, device=dev))
    ndcg = hit / torch.log2(pos.double() + 2)
    prec = hit / k
    f1 = torch.where(hit > 0, 2 * hit * prec / (hit + prec), torch.zeros_like(hit))
    want = torch.stack([hit.mean(), prec.mean(), hit.mean(), ndcg.mean(), f1.mean()])
    return {"rows": n_users, "k": k, "ms": ms, "users_per_s": n_users / (ms * 1e-3), "bytes_read": n_users * (k * 4 + 12),
            "sample_rows_checked": sample, "sample_max_abs_err": float((vals - want).abs().max())}


EVAL_PATHS = {
    "screen": (1, "screen_prep_items_kernel + [gather_rows_kernel] + eval_topk_tc_kernel<SCREEN, INS> (one TF32 tcgen05.mma per product on "
                  "the raw tables, TMEM accumulators, TMA operands, CTA pairs, lists on inserter warps fed through shared-memory rings; "
                  "exact fp32 re-scoring + certificate in the kernel) + device-gated 3xTF32 second pass for uncertified rows"),
    "3xtf32": (3, "tf32_split_kernel x2 + eval_topk_tc_kernel (3xTF32 tcgen05.mma, TMEM accumulators, TMA operands) + topk_merge_kernel"),
    "fp32": (0, "eval_topk_simt_kernel (exact fp32 FMA) + topk_merge_kernel"),
}


def eval_roofline(tensor_flops_per_gpu, ms, mma_per_product=3):
    """Tensor-pipe roofline of the fused eval call on one GPU: the TF32 MMA flops the call EXECUTES (3 per product for 3xTF32, 1 for
    the screened path) against the nominal dense TF32 peak and against half the MEASURED dense bf16 rate (MEASURED_PEAKS.json; cuBLAS
    8192^3, the only measured tensor figure).  The screened path is bound by its epilogue (the per-score maximum scan and the list
    updates), not by the tensor pipe: its fraction is lower although the call is faster."""
    ach = tensor_flops_per_gpu / (ms * 1e-3) / 1e12
    half_bf16 = None
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            half_bf16 = float(json.load(f).get("bf16_tflops", 0)) / 2 or None
    return {"bound": "tensor", "achieved": ach, "peak": TF32_DENSE_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / TF32_DENSE_PEAK_TFLOPS,
            "peak_source": "nominal dense TF32 (2.25 PFLOP/s bf16 / 2)", "frac_of_half_measured_bf16": (ach / half_bf16) if half_bf16 else None,
            "per_gpu": True, "traffic": None,
            "mma_per_product": mma_per_product,
            "note": f"achieved = {mma_per_product} x 2 x K x n_items x n_users TF32 flops / CUDA-event time of the whole call (every kernel "
                    "of the call), per GPU"}


def norm_rel_err(a, b, torch, chunk=1 << 22):
    """max |a - b| / max |b| over (n, d) device tables, in row chunks (no table-sized temporaries)."""
    num = den = 0.0
    for s0 in range(0, a.shape[0], chunk):
        x, y = a[s0:s0 + chunk], b[s0:s0 + chunk]
        num = max(num, float((x - y).abs().max()))
        den = max(den, float(y.abs().max()))
    return num / max(den, 1e-30)


def compare_topk(ids, sc, ref_ids, ref_sc, torch, rtol=1e-5, atol=1e-6):
    """Tie-aware comparison of our (n, k) table with torch.topk's on the same rows: the reference rows are first put in
    the canonical order (score desc, id asc; torch.topk's tie order is unspecified); a row is `exact` when
    the id lists agree, `tied` when they differ only where the position-wise scores agree within rtol/atol (fp32
    summation-order near-ties), else `bad`."""
    o = torch.sort(ref_ids, dim=1, stable=True).indices
    ref_ids, ref_sc = torch.gather(ref_ids, 1, o), torch.gather(ref_sc, 1, o)
    o = torch.sort(ref_sc, dim=1, descending=True, stable=True).indices
    ref_ids, ref_sc = torch.gather(ref_ids, 1, o), torch.gather(ref_sc, 1, o)
    same = (ids.to(torch.int64) == ref_ids).all(1)
    close = ((sc - ref_sc).abs() <= atol + rtol * ref_sc.abs()) | (torch.isinf(sc) & torch.isinf(ref_sc))
    tied = ~same & close.all(1)
    return {"topk_rows": int(ids.shape[0]), "topk_exact": int(same.sum()), "topk_tied": int(tied.sum()),
            "topk_bad": int((~same & ~close.all(1)).sum())}


def sharded_eval_agreement(a_ids, a_sc, b_ids, b_sc, torch):
    """User-range sharding (every rank: its users against ALL items) against item-range sharding (every rank: all users against ITS
    items, then exchange + merge) on the same users.  Both compute exact fp32 scores in the same FMA order, so the tables are
    normally bit-identical; a row that needed the screened path's second pass in one variant only can swap near-tied items (within
    the 3xTF32 error, ~1e-6) at the k-th place, which the tie-aware comparison accepts and anything else fails."""
    bit = bool(torch.equal(a_ids, b_ids) and torch.equal(a_sc, b_sc))
    res = {"item_sharded_bit_identical": bit}
    if bit:
        res["item_sharded_matches_user_sharded"] = True
    else:
        cmp = compare_topk(b_ids, b_sc, a_ids.to(torch.int64), a_sc, torch)
        res["item_sharded_vs_user_sharded"] = cmp
        res["item_sharded_matches_user_sharded"] = cmp["topk_bad"] == 0
    return res


def c2_leg(args, dev, flush, torch, hbm_peak):
    """BASELINE.json configs[1] (and [2], [3] on the same graph) on one GPU: propagation, eval, e2e, training step,
    adv_sampling step, LTR ranking, the torch-on-CUDA comparator and the CPU baseline."""
    from textgcn_b200 import ops
    w = build_workload("c2", dev)
    nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
    n = nu + ni
    k = args.topk
    graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    t = timed_steps(lambda: ops.propagate_fwd(graph, w["uw"], w["iw"], L, out=out), max(args.steps, 10), args.warmup, flush, torch)
    ms = sum(t) / len(t)
    step_bytes = L * spmm_layer_bytes(nnz, n, d) + L * n * 4 * d
    ach = step_bytes / (ms * 1e-3) / 1e9
    res = {"workload": "c2", "n_users": nu, "n_items": ni, "nnz": nnz, "emb": d, "layers": L, "ms_per_step": ms,
           "edges_per_s": nnz * L / (ms * 1e-3), "roofline_achieved_GBs": ach, "roofline_frac": ach / hbm_peak,
           "roofline_note": "the 64.8 MB table is L2-resident, so the no-reuse HBM model is exceeded; compulsory DRAM traffic is "
                            f"{(2 * n * 4 * d + nnz * 8 + n * 4) / 1e6:.0f} MB per layer; the bound that applies is L2 -> SM bandwidth",
           # every gathered row still crosses the L2 -> SM fabric (ncu: L1 hit rate 17 %): the algorithmic bytes against the
           # measured L2 slice throughput cap (~6300 B/clk full chip, /opt/skills/guides/B300_MICROARCH.md "L2 cache") at the max SM clock
           "l2_roofline": {"bound": "l2", "achieved": ach, "peak": L2_CAP_BYTES_PER_CLK * 1965e6 / 1e9, "unit": "GB/s",
                           "frac": ach / (L2_CAP_BYTES_PER_CLK * 1965e6 / 1e9),
                           "peak_source": "LTS throughput cap 6300 B/clk (B300_MICROARCH.md, measured on B300; same L2 design) x 1965 MHz"}}
    if not args.no_e2e:
        h_u = torch.empty((nu, d), dtype=torch.float32).pin_memory().copy_(w["uw"].cpu())
        h_i = torch.empty((ni, d), dtype=torch.float32).pin_memory().copy_(w["iw"].cpu())
        h_o = torch.empty((n, d), dtype=torch.float32).pin_memory()
        stage = torch.empty((2 * n, d), dtype=torch.float32, device=dev)
        te = timed_steps(lambda: ops.propagate_host(graph, h_u, h_i, h_o, L, stage), 5, 3, flush, torch)
        torch.cuda.synchronize()
        res["e2e"] = {"value": nnz * L / (sum(te) / len(te) * 1e-3), "unit": "edges/s", "ms_per_step": sum(te) / len(te),
                      "h2d_bytes_per_step": n * d * 4, "d2h_bytes_per_step": n * d * 4,
                      "matches_device_result": bool(torch.equal(h_o, out.cpu()))}
        del stage
    if not args.no_eval:
        users = torch.arange(nu, dtype=torch.int32, device=dev)
        st = {}
        ops.eval_topk(graph, out[:nu], out[nu:], k, users=users, stats=st)
        mma = EVAL_PATHS[st["precision"]][0]
        te = timed_steps(lambda: ops.eval_topk(graph, out[:nu], out[nu:], k, users=users), args.eval_steps, 3, flush, torch)
        ems = sum(te) / len(te)
        res["eval"] = {"users_per_s": nu / (ems * 1e-3), "ms": ems, "k": k, "n_users_ranked": nu, "precision": st["precision"],
                       "second_pass_rows": st["second_pass_rows"], "kernel": EVAL_PATHS[st["precision"]][1],
                       "tensor_flops_per_s": mma * 2.0 * d * ni * nu / (ems * 1e-3),
                       "roofline": eval_roofline(mma * 2.0 * d * ni * nu, ems, mma)}
        n_f = min(nu, 32768)
        tf = timed_steps(lambda: ops.eval_topk(graph, out[:nu], out[nu:], k, users=users[:n_f].contiguous(), precision="fp32"),
                         1, 1, flush, torch)
        res["eval"]["fp32_simt_users_per_s"] = n_f / (sum(tf) / len(tf) * 1e-3)
    if not args.no_train:
        try:
            res["train"] = train_leg(w, graph, dev, flush, torch)
        except Exception as exc:
            res["train"] = {"error": str(exc)[:300]}
    if not args.no_extras:
        try:
            res["configs"] = extras_leg(w, graph, dev, flush, torch)
        except Exception as exc:
            res["configs"] = {"error": str(exc)[:300]}
    if not args.no_torch_ref:
        try:
            res["torch_cuda_reference"], ref_emb, ref_ids, ref_sc = torch_cuda_reference(w, dev, flush, torch, k)
            res["parity"] = parity_vs_torch(ops, graph, out, nu, k, ref_emb, ref_ids, ref_sc, torch)
            del ref_emb, ref_ids, ref_sc
        except Exception as exc:
            res["torch_cuda_reference"] = {"error": str(exc)[:300]}
    if not args.no_cpu_baseline:
        res["cpu_baseline"] = cpu_baseline(w, k)
    return res


PARITY_TOL = 1e-5  # north_star: propagated embeddings and scores within 1e-5 relative (norm-wise)


def parity_vs_torch(ops, graph, out, nu, k, ref_emb, ref_ids, ref_sc, torch):
    """Our result against the reference's own ops on the same GPU in the same run: `out` vs torch.sparse's representation
    (norm-wise), the fused top-k vs matmul + index_put + topk on the same users (tie-aware)."""
    n_ref = ref_ids.shape[0]
    users = torch.arange(n_ref, dtype=torch.int32, device=out.device)
    ids, sc = ops.eval_topk(graph, out[:nu], out[nu:], k, users=users)
    res = {"prop_rel_err": norm_rel_err(out, ref_emb, torch), "tolerance": PARITY_TOL}
    res.update(compare_topk(ids, sc, ref_ids, ref_sc, torch))
    res["ok"] = bool(res["prop_rel_err"] <= PARITY_TOL and res["topk_bad"] == 0)
    return res


def extras_leg(w, graph, dev, flush, torch, batch=2048):
    """BASELINE.json configs[2] (adv_sampling step: device candidate + positive sampling, hardest-of-1000 selection, BPR on
    up to B·5·k triples, backward, fused Adam) and configs[3] (LTR ranking with random 768-d text tables)."""
    import logging
    from textgcn_b200.models import AdvSamplModel, LTRLinearWPop, make_params
    from textgcn_b200.optim import FusedAdam
    from textgcn_b200.sampler import AdvEpochSampler

    class DS:
        pass

    ds = DS()
    ds.n_users, ds.n_items, ds.graph, ds.norm_matrix = w["nu"], w["ni"], graph, None
    ds.test_users, ds.true_test_lil = [0], [[0]]
    log = logging.getLogger("bench")
    out = {}
    params = make_params(emb_size=w["d"], n_layers=w["L"], k=[20], batch_size=batch, fused_adam=True, dropout_rng="device",
                         positive_sampler="device", device=dev, logger=log)
    model = AdvSamplModel(params, ds)
    opt = FusedAdam(model.parameters(), lr=params.lr)
    smp = AdvEpochSampler(graph, batch_size=batch, seed=0)
    users = torch.arange(batch, dtype=torch.int32, device=dev)
    model.train()
    model.training = True
    n_triples = []

    def adv_step():
        data = smp.sample(users, 1234)
        opt.zero_grad(set_to_none=False)
        triples = model.select_triples(data)
        n_triples.append(triples.shape[0])
        loss = super(AdvSamplModel, model).get_loss(triples)
        loss.backward()
        opt.step()

    t = timed_steps(adv_step, 5, 2, flush, torch)
    adv_ms = sum(t) / len(t)
    nnz, n, d, L, keep_p, T = w["nnz"], w["nu"] + w["ni"], w["d"], w["L"], 1.0 - params.dropout, n_triples[-1]
    spmm_pass = nnz * 9 + keep_p * nnz * 4 * d + n * 4 * d
    adv_bytes = (batch * (1 + smp.n_cand) * 8 + 2 * nnz                # candidate sampler, two dropout draws (G12)
                 + 2 * (L * spmm_pass + L * n * 4 * d)                 # two propagations per step
                 + batch * smp.n_cand * (4 * d + 4)                    # adv_select_kernel gathers
                 + 2 * n * 4 * d + T * 3 * 2 * 4 * d * 2               # gradient zero fill, bpr_kernel on T triples
                 + 6 * nnz + L * (spmm_pass + n * 4 * d) + 7 * n * 4 * d)   # backward + Adam
    out["adv_sampling"] = {"ms_per_step": adv_ms, "batch_users": batch, "candidates": smp.n_cand, "k": 20,
                           "triples_per_step": n_triples[-1],
                           "roofline": {"bound": "l2 / hbm", "algorithmic_bytes": adv_bytes, "achieved": adv_bytes / (adv_ms * 1e-3) / 1e9,
                                        "unit": "GB/s", "frac_of_hbm_peak": adv_bytes / (adv_ms * 1e-3) / 1e9 / peaks()[0],
                                        "note": "byte models of the step's kernels / CUDA-event time; the step also contains torch glue "
                                                "(nonzero / stack building the (T, 3) triples) and a host sync for T"},
                           "includes": "candidate + positive sampling kernels, propagate, adv_select_kernel, second propagate + fused BPR, "
                                       "Horner backward, fused Adam"}
    del model, opt
    D = 768
    gen = torch.Generator(device=dev).manual_seed(3)
    for name_, shape in (("items_as_avg_reviews", (w["ni"], D)), ("items_as_desc", (w["ni"], D)), ("users_as_avg_reviews", (w["nu"], D)),
                         ("users_as_avg_desc", (w["nu"], D))):
        setattr(ds, name_, torch.randn(*shape, generator=gen, device=dev))
    ds.popularity_users = torch.rand(w["nu"], 1, generator=gen, device=dev)
    ds.popularity_items = torch.rand(w["ni"], 1, generator=gen, device=dev)
    ltr = LTRLinearWPop(make_params(emb_size=w["d"], n_layers=w["L"], k=[20], device=dev, logger=log), ds)
    n_eval = 18944 if os.environ.get("TGCN_LTR_EVAL_USERS") is None else int(os.environ["TGCN_LTR_EVAL_USERS"])
    t = timed_steps(lambda: ltr.predict_device(torch.arange(n_eval, dtype=torch.int32, device=dev)), 2, 1, flush, torch)
    ms = sum(t) / len(t)
    out["ltr_pop"] = {"users_per_s": n_eval / (ms * 1e-3), "ms": ms, "n_users_ranked": n_eval, "text_dim": D, "contraction_width": w["d"] + 2 * D,
                      "tensor_flops_per_s": 3 * 2.0 * (w["d"] + 2 * D) * w["ni"] * n_eval / (ms * 1e-3),
                      "includes": "propagate, ltr_pack_items/users, tf32_split_kernel x2, eval_topk_tc_kernel<256,20,2,stream> (3xTF32, user and "
                                  "item K-chunks streamed together, bias terms folded into one extra K-chunk)"}
    ltr.eval_precision = "fp32"
    t = timed_steps(lambda: ltr.predict_device(torch.arange(n_eval, dtype=torch.int32, device=dev)), 2, 1, flush, torch)
    out["ltr_pop"]["fp32_simt_users_per_s"] = n_eval / (sum(t) / len(t) * 1e-3)
    return out


def reference_sample(w, torch, max_user_nnz=8_000_000):
    """Bounded CPU sample of a workload for the reference's own BaseModel: the sub-matrix of Â induced by the first U_s
    users and ALL items — their user rows AND the matching entries of every item row, so the long item rows are timed
    too — with U_s the largest prefix whose user rows hold <= max_user_nnz non-zeros (the whole graph at c2).  Values are
    Â's own (same work per non-zero).  Returns CPU numpy CSR arrays with columns renumbered to the sample."""
    nu, ni = w["nu"], w["ni"]
    rowptr = w["rowptr"].to(torch.int64)
    full = int(rowptr[nu]) <= max_user_nnz
    us = nu if full else int(torch.searchsorted(rowptr[:nu + 1].contiguous(), torch.tensor(max_user_nnz, device=rowptr.device), right=True)) - 1
    n_u = int(rowptr[us])
    lo = int(rowptr[nu])
    icol, ival = w["col"][lo:], w["val"][lo:]
    keep = icol < us
    counts = rowptr[nu + 1:] - rowptr[nu:-1]
    rows = torch.repeat_interleave(torch.arange(ni, device=rowptr.device), counts)[keep]
    ip = torch.zeros(ni + 1, dtype=torch.int64, device=rowptr.device)
    ip[1:] = torch.cumsum(torch.bincount(rows, minlength=ni), 0)
    rp = torch.cat([rowptr[:us + 1], n_u + ip[1:]])
    col = torch.cat([w["col"][:n_u].to(torch.int64) - nu + us, icol[keep].to(torch.int64)])
    val = torch.cat([w["val"][:n_u], ival[keep]])
    return dict(n_users=us, n_items=ni, full=full, nnz=int(col.numel()), rowptr=rp.cpu().numpy(), col=col.cpu().numpy(),
                val=val.cpu().numpy(), user_w=w["uw"][:us].cpu(), item_w=w["iw"].cpu())


def reference_cpu_run(w, topk, steps, warmup, n_predict, budget_s=150.0):
    """The reference's CPU path on the box's host cores: its own ``BaseModel`` (unmodified, /root/reference or the copy in
    baseline/_ref) where importable (kind "reference"), else the restated op sequence of oracle/lightgcn_oracle.py (kind
    "port").  ``representation`` timed `steps` times after `warmup` calls; ``predict`` on the first n_predict users once."""
    import numpy as np
    import torch
    from oracle import lightgcn_oracle as O
    from oracle import reference_cpu as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    smp = reference_sample(w, torch)
    L, nnz_s = w["L"], smp["nnz"]
    what = (f"{'full graph' if smp['full'] else 'sub-matrix of Â induced by the first ' + str(smp['n_users']) + ' users and all items'} of "
            f"{w['name']}: {smp['n_users'] + smp['n_items']} rows, {nnz_s} nnz (user rows and item rows), emb {w['d']}, {L} layers")
    n_predict = min(n_predict, smp["n_users"])
    users = np.arange(n_predict)
    if R.reference_root() is not None:
        model = R.build_model(smp["n_users"], smp["n_items"], smp["rowptr"], smp["col"], smp["val"], smp["user_w"], smp["item_w"], L,
                              [topk], batch_size=2048)
        times = R.time_representation(model, steps, warmup, budget_s)
        t_pred, t_rep, _ = R.time_predict(model, users)
        kind = "reference"
        how = (f"the reference's own BaseModel.representation (base_model.py:93-106), unmodified, device='cpu', {len(times)} timed calls "
               f"after {warmup} warm-up; BaseModel.predict (:235-276) on {n_predict} users x {smp['n_items']} items = {t_pred * 1e3:.0f} ms "
               f"including its own representation call ({t_rep * 1e3:.0f} ms)")
        pred_only = max(t_pred - t_rep, 1e-9)
    else:
        row = np.repeat(np.arange(len(smp["rowptr"]) - 1), np.diff(smp["rowptr"]))
        norm = O.sparse_tensor(row, smp["col"], smp["val"], len(smp["rowptr"]) - 1)
        times = []
        for it in range(warmup + steps):
            t = time.perf_counter()
            ue, ie = O.propagate(norm, smp["user_w"], smp["item_w"], L)
            if it >= warmup:
                times.append(time.perf_counter() - t)
        ucol = smp["col"][:smp["rowptr"][smp["n_users"]]] - smp["n_users"]
        lists = np.split(ucol, smp["rowptr"][1:smp["n_users"]])
        t = time.perf_counter()
        O.predict_topk_torch(ue, ie, users, lists, topk)
        pred_only = time.perf_counter() - t
        kind = "port"
        how = f"oracle port (reference not importable): torch.sparse.mm x L + mean, {len(times)} timed calls; predict on {n_predict} users"
    mean_s = sum(times) / len(times)
    return {"value": nnz_s * L / mean_s, "unit": "edges/s", "cores": cores, "kind": kind, "sample": what + "; " + how,
            "ms_per_step": mean_s * 1e3, "best_ms": min(times) * 1e3, "steps": len(times), "warmup": warmup,
            "eval_users_per_s": n_predict / pred_only, "eval_users_per_s_incl_representation": n_predict / (pred_only + mean_s),
            "sample_nnz": nnz_s, "workload_nnz": w["nnz"]}


def cpu_baseline(w, topk):
    """cpu_baseline leg of our arm (N = 1, rank 0): the reference's CPU path on a bounded sample, best of 3."""
    res = reference_cpu_run(w, topk, steps=3, warmup=1, n_predict=8192 if w["nnz"] <= 16_000_000 else 2048, budget_s=60.0)
    res["value"] = res["sample_nnz"] * w["L"] / (res["best_ms"] * 1e-3)   # best of 3
    return res


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, on a bounded sample of OUR
    arm's workload (same config / metric / unit; throughput is per non-zero, so the sample and the full graph compare)."""
    if rank != 0:
        return
    import torch
    name = args.workload if args.workload != "auto" else "c5"
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    w = build_workload(name, dev)
    res = reference_cpu_run(w, args.topk, steps=args.steps, warmup=args.warmup, n_predict=8192 if name != "c5" else 2048)
    value = res["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": args.gpus, "steps": res["steps"],
        "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "n_users": w["nu"], "n_items": w["ni"], "nnz": w["nnz"], "emb": w["d"], "layers": w["L"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "eval": {"users_per_s": res["eval_users_per_s"], "k": args.topk,
                 "users_per_s_incl_representation": res["eval_users_per_s_incl_representation"]},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from textgcn_b200 import ops
    from textgcn_b200 import dist as tdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL logs (e.g. its version banner at NCCL_DEBUG=VERSION) go to stdout by default: keep stdout to the ONE JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload if args.workload != "auto" else "c5"
    w = build_workload(name, dev)
    nu, ni, d, L, nnz = w["nu"], w["ni"], w["d"], w["L"], w["nnz"]
    n = nu + ni
    if world > 1:
        # every rank builds the workload itself: they must agree bit for bit (exact integer / float64 checksums), else the
        # multi-GPU result is meaningless — refuse to time it
        sums = torch.stack([w["rowptr"].long().sum(), w["col"].long().sum()] +
                           [w[key].view(torch.int32).long().sum() for key in ("val", "uw", "iw")])   # int64 sums of the bit patterns: exact
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            # never time ranks that disagree on the graph: adopt rank 0's copy everywhere (same shapes: the generator emits
            # exactly n_edges rows) and say so in the line
            for key in ("rowptr", "col", "val", "uw", "iw"):
                dist.broadcast(w[key], src=0)
            workload_note = f"ranks built different workloads (checksum spread {(hi - lo).tolist()}): rank 0's copy was broadcast"
        else:
            workload_note = None
    else:
        workload_note = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm_peak, peak_src = peaks()
    sampler = ClockSampler(local_rank)

    step_bytes = L * spmm_layer_bytes(nnz, n, d) + L * n * 4 * d  # + mean epilogue reads of E0..E_{L-1}
    extra = {}
    if world == 1:
        graph = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
        out = torch.empty((n, d), dtype=torch.float32, device=dev)

        def step():
            ops.propagate_fwd(graph, w["uw"], w["iw"], L, out=out)

        sampler.start()
        times = timed_steps(step, args.steps, args.warmup, flush, torch)
        launches_per_step = L
        parallelism = "single GPU"
        scaling = "strong"
    elif args.mg_scheme == "grid":
        shapes = [{2: (1, 2), 4: (1, 4), 8: (2, 4)}.get(world, (1, world))] if args.grid == "auto" else \
            [tuple(int(x) for x in sh.lower().split(":")[0].split("x")) + (sh.lower().endswith(":hp"),) for sh in args.grid.split(",")]
        shapes = [sh if len(sh) == 3 else sh + (args.nccl_high_priority,) for sh in shapes]

        def make_grid(G, R, hp):
            assert G * R == world, f"--grid {G}x{R} does not match {world} ranks"
            gpart = tdist.GridPartition(w["rowptr"], nu, ni, d, G, R)
            gg, rr = gpart.coords(rank)
            row_group = None
            for g_id in range(G):  # every rank creates every row group, in the same order
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True) if hp else None
                grp = dist.new_group(gpart.row_group_ranks(g_id), pg_options=opts) if R > 1 else None
                if g_id == gg:
                    row_group = grp
            u0, u1 = gpart.rows.users(rr)
            ug = ops.Graph(nu, ni, *gpart.rows.user_block(rr, w["rowptr"], w["col"], w["val"]), row_begin=u0, block=True)
            ig = ops.Graph(nu, ni, *gpart.rows.item_block(rr, w["rowptr"], w["col"], w["val"]), row_begin=nu, block=True)
            prop = tdist.GridPropagator(gpart, rank, ug, ig, L, dev, row_group=row_group, exchange=args.exchange)
            c0, c1 = gpart.cols(gg)
            return gpart, prop, w["uw"][u0:u1, c0:c1].contiguous(), w["iw"][:, c0:c1].contiguous()

        # extra shapes first (comparison points), the headline shape last so its tables are the ones eval reads
        for G, R, hp in shapes[1:]:
            gpart_x, prop_x, xu, xi = make_grid(G, R, hp)
            dist.barrier()
            tx = timed_steps(lambda: prop_x.propagate(xu, xi), args.steps, args.warmup, flush, torch)
            mx = torch.tensor([sum(tx)], dtype=torch.float64, device=dev)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            extra.setdefault("other_grids", {})[f"{G}x{R}" + (":hp" if hp else "")] = {"ms_per_step": float(mx) / args.steps,
                                                                "value": nnz * L / (float(mx) / args.steps * 1e-3)}
            dist.barrier()
            prop_x.close()
            del gpart_x, prop_x, xu, xi
            torch.cuda.empty_cache()
        G, R, hp = shapes[0]
        gpart, prop, e0_u, e0_i = make_grid(G, R, hp)

        def step():
            prop.propagate(e0_u, e0_i)

        dist.barrier()
        sampler.start()
        times = timed_steps(step, args.steps, args.warmup, flush, torch)
        dist.barrier()
        if os.environ.get("TGCN_GRID_TIMING"):
            prop.timing_report()
            step()
            extra["grid_timing_rank0"] = prop.timing_report()
            dist.barrier()
        launches_per_step = 2 * L + 1
        parallelism = (f"grid {G}x{R}: {G} feature slices of {gpart.ds} columns x {R} user partitions by nnz; per hop one NCCL "
                       f"all-reduce of the (I, {gpart.ds}) item slice inside each row group of {R}; result exchange "
                       f"{'fused into the last passes as peer-memory stores (CUDA IPC, NVLink)' if args.exchange == 'p2p' else 'by NCCL all-to-all + all-gather'}")
        scaling = "strong"
        extra["comm_bytes_per_hop_per_rank"] = prop.comm_bytes_per_hop
        f0, f1 = gpart.final_users(rank)
        rows_local = (f1 - f0) + ni
        out_u, out_i = prop.out_u, prop.out_i
    elif args.mg_scheme == "rowblock":
        part = tdist.RowPartition(w["rowptr"], world)
        rp, col, val = part.local_block(rank, w["rowptr"], w["col"], w["val"])
        s, e = part.rows(rank)
        lgraph = ops.Graph(nu, ni, rp, col, val, row_begin=s, block=True)
        prop = tdist.DistPropagator(part, rank, lgraph, d, L, dev)
        e0 = torch.cat([w["uw"], w["iw"]])[s:e].contiguous()
        out_local = torch.empty((e - s, d), dtype=torch.float32, device=dev)

        def step():
            prop.propagate(e0, out_local)

        dist.barrier()
        sampler.start()
        times = timed_steps(step, args.steps, args.warmup, flush, torch)
        dist.barrier()
        launches_per_step = L
        parallelism = f"row-block x{world}, NCCL all-gather of layer embeddings between hops"
        scaling = "strong"
        extra["comm_bytes_per_hop_per_rank"] = prop.comm_bytes_per_hop
        rows_local = e - s
    else:
        part = tdist.BipartitePartition(w["rowptr"], nu, ni, world)
        u0, u1 = part.users(rank)
        ugraph = ops.Graph(nu, ni, *part.user_block(rank, w["rowptr"], w["col"], w["val"]), row_begin=u0, block=True)
        ugraph.set_mask_col_offset(0)
        igraph = ops.Graph(nu, ni, *part.item_block(rank, w["rowptr"], w["col"], w["val"]), row_begin=nu, block=True)
        chunks = None
        if args.ar_chunks > 1:  # chunked item rows: each chunk's all-reduce starts as soon as its SpMM is enqueued
            irp, icol, ival = igraph.rowptr, igraph.col, igraph.val
            chunks = []
            for c in range(args.ar_chunks):
                r0, r1 = ni * c // args.ar_chunks, ni * (c + 1) // args.ar_chunks
                lo, hi = int(irp[r0]), int(irp[r1])
                h = ops.Graph(nu, ni, (irp[r0:r1 + 1] - irp[r0]).contiguous(), icol[lo:hi].contiguous(), ival[lo:hi].contiguous(),
                              row_begin=nu + r0, block=True)
                chunks.append((r0, r1, h))
        prop = tdist.BipartitePropagator(part, rank, ugraph, igraph, d, L, dev, item_chunks=chunks)
        e0_u = w["uw"][u0:u1].contiguous()
        out_u = torch.empty((u1 - u0, d), dtype=torch.float32, device=dev)
        out_i = torch.empty((ni, d), dtype=torch.float32, device=dev)

        def step():
            prop.propagate(e0_u, w["iw"], out_u, out_i)

        dist.barrier()
        sampler.start()
        times = timed_steps(step, args.steps, args.warmup, flush, torch)
        dist.barrier()
        launches_per_step = 2 * L + 1 + (args.ar_chunks - 1) * L
        parallelism = f"users partitioned x{world} by nnz, item table replicated: NCCL all-reduce of (I, d) per hop overlapped with the user-row SpMM"
        scaling = "strong"
        extra["comm_bytes_per_hop_per_rank"] = prop.comm_bytes_per_hop
        rows_local = (u1 - u0) + ni
    torch.cuda.synchronize()
    total_ms = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms) / args.steps
    value = nnz * L / (ms_per_step * 1e-3)

    # ---- e2e through the host-buffer C-ABI entry (N = 1) / host-staged shard (N > 1) -------------------------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            h_u = torch.empty((nu, d), dtype=torch.float32).pin_memory().copy_(w["uw"].cpu())
            h_i = torch.empty((ni, d), dtype=torch.float32).pin_memory().copy_(w["iw"].cpu())
            h_o = torch.empty((n, d), dtype=torch.float32).pin_memory()
            stage = torch.empty((2 * n, d), dtype=torch.float32, device=dev)

            def e2e_step():
                ops.propagate_host(graph, h_u, h_i, h_o, L, stage)
        elif args.mg_scheme == "grid":
            h_u = torch.empty_like(e0_u, device="cpu").pin_memory().copy_(e0_u.cpu())
            h_i = torch.empty_like(e0_i, device="cpu").pin_memory().copy_(e0_i.cpu())
            h_ou = torch.empty((f1 - f0, d), dtype=torch.float32).pin_memory()
            # items_emb is replicated after the exchange: every rank returns 1/P of it (round 1 copied all of it from every rank)
            i0, i1 = tdist.item_shard(ni, world, rank)
            h_oi = torch.empty((i1 - i0, d), dtype=torch.float32).pin_memory()
            d_u, d_i = torch.empty_like(e0_u), torch.empty_like(e0_i)
            rows_out = (f1 - f0) + (i1 - i0)

            def e2e_step():
                d_u.copy_(h_u, non_blocking=True)
                d_i.copy_(h_i, non_blocking=True)
                ou, oi = prop.propagate(d_u, d_i)
                h_ou.copy_(ou, non_blocking=True)
                h_oi.copy_(oi[i0:i1], non_blocking=True)
        elif args.mg_scheme == "rowblock":
            h_e0 = torch.empty_like(e0, device="cpu").pin_memory().copy_(e0.cpu())
            h_o = torch.empty_like(out_local, device="cpu").pin_memory()
            d_e0 = torch.empty_like(e0)

            def e2e_step():
                d_e0.copy_(h_e0, non_blocking=True)
                prop.propagate(d_e0, out_local)
                h_o.copy_(out_local, non_blocking=True)
        else:
            h_u = torch.empty_like(e0_u, device="cpu").pin_memory().copy_(e0_u.cpu())
            h_i = torch.empty_like(w["iw"], device="cpu").pin_memory().copy_(w["iw"].cpu())
            h_ou = torch.empty_like(out_u, device="cpu").pin_memory()
            h_oi = torch.empty_like(out_i, device="cpu").pin_memory()
            d_u, d_i = torch.empty_like(e0_u), torch.empty_like(w["iw"])

            def e2e_step():
                d_u.copy_(h_u, non_blocking=True)
                d_i.copy_(h_i, non_blocking=True)
                prop.propagate(d_u, d_i, out_u, out_i)
                h_ou.copy_(out_u, non_blocking=True)
                h_oi.copy_(out_i, non_blocking=True)
        if world > 1:
            dist.barrier()
        et = timed_steps(e2e_step, max(3, args.steps // 2), 3, flush, torch)
        e_ms = torch.tensor([sum(et) / len(et)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        if world == 1:
            rows_local = n
        e2e = {"value": nnz * L / (float(e_ms) * 1e-3), "unit": "edges/s", "ms_per_step": float(e_ms),
               "h2d_bytes_per_step": (e0_u.numel() + e0_i.numel()) * 4 if (world > 1 and args.mg_scheme == "grid") else rows_local * d * 4,
               "d2h_bytes_per_step": (rows_out if (world > 1 and args.mg_scheme == "grid") else rows_local) * d * 4,
               "api": "tgcn_propagate_host (pinned host tables -> device, L layers, result -> pinned host; uploads / downloads on copy "
                      "streams overlapping the first / last layer)" if world == 1
               else "host-pinned E0 shard -> device, L hops with the collective, result shard (own user rows + 1/P of the item table) "
                    "-> pinned host; bytes are per rank"}
        if world == 1:   # the host-buffer call must return exactly what the device-resident call computed (first and last 1M rows)
            torch.cuda.synchronize()
            e2e["matches_device_result"] = bool(torch.equal(h_o[:1 << 20], out[:1 << 20].cpu()) and
                                                torch.equal(h_o[-(1 << 20):], out[-(1 << 20):].cpu()))

    # ---- eval leg: fused score + mask + top-k ----------------------------------------------------------------
    ev = None
    if not args.no_eval:
        k = args.topk
        if world == 1:
            emb = out
            n_eval = args.eval_users or (nu if name != "c5" else EVAL_USERS_C5)
            users = torch.arange(n_eval, dtype=torch.int32, device=dev)
            h_users = torch.arange(n_eval, dtype=torch.int32).pin_memory()
            h_ids = torch.empty((n_eval, k), dtype=torch.int32).pin_memory()
            h_sc = torch.empty((n_eval, k), dtype=torch.float32).pin_memory()

            def eval_step():
                return ops.eval_topk(graph, emb[:nu], emb[nu:], k, users=users)

            def eval_e2e():
                du = h_users.to(dev, non_blocking=True)
                ids, sc = ops.eval_topk(graph, emb[:nu], emb[nu:], k, users=du)
                h_ids.copy_(ids, non_blocking=True)
                h_sc.copy_(sc, non_blocking=True)
        else:
            # headline: user-range sharding (comm-free): every rank ranks a slice of ITS users against all items
            n_eval = args.eval_users or (nu if name != "c5" else EVAL_USERS_C5)   # the SAME user count at every N: strong scaling
            n_eval = (n_eval + world - 1) // world * world
            per = n_eval // world
            if args.mg_scheme == "rowblock":
                full = prop.gather_full(out_local)
                u_tab, i_tab = full[:nu], full[nu:]
                sample = torch.arange(rank * per, (rank + 1) * per, dtype=torch.int32, device=dev)
                mrows = int(w["rowptr"][nu])
                mgraph = ops.Graph(nu, ni, w["rowptr"][:nu + 1].contiguous(), w["col"][:mrows].contiguous(),
                                   w["val"][:mrows].contiguous(), row_begin=0, block=True)

                def eval_step():
                    return ops.eval_topk(mgraph, u_tab, i_tab, k, users=sample)
            elif args.mg_scheme == "grid":
                per = min(per, gpart.final_users(world - 1)[1] - gpart.final_users(world - 1)[0])
                n_eval = per * world
                sample = torch.arange(f0, f0 + per, dtype=torch.int32, device=dev)  # global ids of this rank's first users
                mrows = int(w["rowptr"][nu])
                mgraph = ops.Graph(nu, ni, w["rowptr"][:nu + 1].contiguous(), w["col"][:mrows].contiguous(),
                                   w["val"][:mrows].contiguous(), row_begin=0, block=True)
                out_u_eval = out_u[:per]

                def eval_step():
                    return ops.eval_topk(mgraph, out_u_eval, out_i, k, users=sample, by_position=True)

                all_users = torch.empty(n_eval, dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(all_users, sample)
                all_vecs = torch.empty((n_eval, d), dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(all_vecs, out_u_eval.contiguous())

                def eval_item_sharded():
                    return tdist.sharded_eval_topk(mgraph, all_vecs, out_i, all_users, k, rank, world, by_position=True)

                a_ids, a_sc = eval_step()
                b_ids, b_sc = eval_item_sharded()
                extra.update(sharded_eval_agreement(a_ids, a_sc, b_ids, b_sc, torch))
                t_is = timed_steps(eval_item_sharded, args.eval_steps, 3, flush, torch)
                is_ms = torch.tensor([sum(t_is) / len(t_is)], dtype=torch.float64, device=dev)
                dist.all_reduce(is_ms, op=dist.ReduceOp.MAX)
                extra["eval_item_sharded"] = {"users_per_s": n_eval / (float(is_ms) * 1e-3), "ms": float(is_ms),
                                              "sharding": f"item range x{world}, all-to-all of partial top-k + merge"}
            else:
                per = min(per, u1 - u0)
                n_eval = per * world
                sample = torch.arange(u0, u0 + per, dtype=torch.int32, device=dev)
                mrows = int(w["rowptr"][nu])
                mgraph = ops.Graph(nu, ni, w["rowptr"][:nu + 1].contiguous(), w["col"][:mrows].contiguous(),
                                   w["val"][:mrows].contiguous(), row_begin=0, block=True)

                def eval_step():
                    return ops.eval_topk(ugraph, out_u, out_i, k, users=sample, by_position=True)

                # north_star variant: item-range sharding + cross-GPU merge over the same users
                all_users = torch.empty(n_eval, dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(all_users, sample)
                all_vecs = torch.empty((n_eval, d), dtype=torch.float32, device=dev)
                dist.all_gather_into_tensor(all_vecs, out_u[:per].contiguous())

                def eval_item_sharded():
                    return tdist.sharded_eval_topk(mgraph, all_vecs, out_i, all_users, k, rank, world, by_position=True)

                a_ids, a_sc = eval_step()
                b_ids, b_sc = eval_item_sharded()
                extra.update(sharded_eval_agreement(a_ids, a_sc, b_ids, b_sc, torch))
                t_is = timed_steps(eval_item_sharded, args.eval_steps, 3, flush, torch)
                is_ms = torch.tensor([sum(t_is) / len(t_is)], dtype=torch.float64, device=dev)
                dist.all_reduce(is_ms, op=dist.ReduceOp.MAX)
                extra["eval_item_sharded"] = {"users_per_s": n_eval / (float(is_ms) * 1e-3), "ms": float(is_ms),
                                              "sharding": f"item range x{world}, all-to-all of partial top-k + merge"}
            eval_e2e = None
        t_ev = timed_steps(eval_step, args.eval_steps, 3, flush, torch)
        ev_ms = torch.tensor([sum(t_ev) / len(t_ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ev_ms, op=dist.ReduceOp.MAX)
        prec = ops.eval_resolve_precision(ni, d, k)  # which path the call takes (auto precision)
        mma = EVAL_PATHS[prec][0]
        ev = {"users_per_s": n_eval / (float(ev_ms) * 1e-3), "k": k, "n_users_ranked": n_eval, "n_items": ni,
              "ms": float(ev_ms), "score_flops_per_s": 2.0 * d * ni * n_eval / (float(ev_ms) * 1e-3), "precision": prec,
              "kernel": EVAL_PATHS[prec][1],
              "tensor_flops_per_s": mma * 2.0 * d * ni * n_eval / (float(ev_ms) * 1e-3), "scaling": "strong",
              "sharding": "single GPU" if world == 1 else f"user range x{world} (comm-free); item-range variant in eval_item_sharded"}
        ev["roofline"] = eval_roofline(mma * 2.0 * d * ni * n_eval / world, float(ev_ms), mma)
        if world == 1:
            st = {}
            ops.eval_topk(graph, emb[:nu], emb[nu:], k, users=users, stats=st)
            ev["second_pass_rows"] = st["second_pass_rows"]
            if prec == "screen":  # the 3xTF32 variant on the same users, for comparison
                t_3 = timed_steps(lambda: ops.eval_topk(graph, emb[:nu], emb[nu:], k, users=users, precision="3xtf32"),
                                  max(1, args.eval_steps - 1), 1, flush, torch)
                ms3 = sum(t_3) / len(t_3)
                ev["3xtf32"] = {"users_per_s": n_eval / (ms3 * 1e-3), "ms": ms3, "tensor_flops_per_s": 3 * 2.0 * d * ni * n_eval / (ms3 * 1e-3),
                                "roofline": eval_roofline(3 * 2.0 * d * ni * n_eval, ms3, 3)}
            n_f = min(n_eval, 32768)
            users_f = users[:n_f].contiguous()
            t_f = timed_steps(lambda: ops.eval_topk(graph, emb[:nu], emb[nu:], k, users=users_f, precision="fp32"),
                              max(1, args.eval_steps - 1), 1, flush, torch)
            ev["fp32_simt_users_per_s"] = n_f / (sum(t_f) / len(t_f) * 1e-3)
            ev["fp32_simt_note"] = f"eval_topk_simt_kernel (exact fp32 FMA) on {n_f} users"
        if world == 1 and name == "c5":
            ev["metrics_10M_users"] = metrics_leg(ops, eval_step()[0], nu, ni, k, dev, flush, torch)
        if eval_e2e is not None:
            t_e = timed_steps(eval_e2e, args.eval_steps, 3, flush, torch)
            ev["e2e_users_per_s"] = n_eval / (sum(t_e) / len(t_e) * 1e-3)
            ev["e2e_h2d_bytes"] = n_eval * 4
            ev["e2e_d2h_bytes"] = n_eval * k * 8
    sampler.stop_flag = True
    sampler.join(timeout=1)

    # ---- training step (a7-a10): dropout draw + propagate + fused BPR + Horner backward + fused Adam ---------
    if world == 1 and not args.no_train:
        try:
            torch.cuda.empty_cache()
            # c5: the whole fused step at the headline scale (12 M x 128 tables, 400 M nnz; ~98 GB peak), fewer timed steps
            extra["train"] = train_leg(w, graph, dev, flush, torch) if name != "c5" else train_leg(w, graph, dev, flush, torch, steps=5, warmup=3)
        except Exception as exc:
            extra["train"] = {"error": str(exc)[:300]}
        torch.cuda.empty_cache()

    if world == 1 and not args.no_extras and name != "c5":
        try:
            extra["configs"] = extras_leg(w, graph, dev, flush, torch)
        except Exception as exc:
            extra["configs"] = {"error": str(exc)[:300]}

    # ---- same workload on ONE GPU, measured by rank 0 in the same run (for honest strong-scaling ratios) ----
    parity = None
    if world > 1:
        # every rank recomputes the whole result on its own GPU with the single-GPU path and checks ITS shard of the
        # multi-GPU result against it (rank 0 also times it: the honest strong-scaling denominator)
        try:
            g1 = ops.Graph(nu, ni, w["rowptr"], w["col"], w["val"])
            out1 = torch.empty((n, d), dtype=torch.float32, device=dev)
            ops.propagate_fwd(g1, w["uw"], w["iw"], L, out=out1)
            if args.mg_scheme == "grid":
                shard = [(out_u[:f1 - f0], out1[f0:f1]), (out_i, out1[nu:])]
            elif args.mg_scheme == "rowblock":
                shard = [(out_local, out1[s:e])]
            else:
                shard = [(out_u, out1[u0:u1]), (out_i, out1[nu:])]
            errs = [norm_rel_err(a, b, torch) if a.numel() else 0.0 for a, b in shard] + [0.0]
            err = torch.tensor([max(errs), errs[0], errs[1]], dtype=torch.float64, device=dev)
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            parity = {"mg_vs_n1_rel_err": float(err[0]), "users_emb_rel_err": float(err[1]), "items_emb_rel_err": float(err[2]),
                      "tolerance": PARITY_TOL, "checked": "every rank's shard of users_emb / items_emb "
                      "against the single-GPU tgcn_propagate_fwd result recomputed on the same GPU", "ok": bool(float(err[0]) <= PARITY_TOL)}
            if rank == 0:
                t1 = timed_steps(lambda: ops.propagate_fwd(g1, w["uw"], w["iw"], L, out=out1), max(2, args.steps // 4), 1, flush, torch)
                extra["n1_same_workload"] = {"value": nnz * L / (sum(t1) / len(t1) * 1e-3), "ms_per_step": sum(t1) / len(t1)}
            del g1, out1, shard
        except Exception as exc:  # e.g. out of memory on a shared device
            parity = {"error": str(exc)[:200], "ok": False}
        dist.barrier()

    if world == 1 and not args.no_torch_ref:
        try:
            extra["torch_cuda_reference"], ref_emb, ref_ids, ref_sc = torch_cuda_reference(w, dev, flush, torch, args.topk,
                                                                                           n_eval=2048 if name == "c5" else 8192)
            parity = parity_vs_torch(ops, graph, out, nu, args.topk, ref_emb, ref_ids, ref_sc, torch)
            del ref_emb, ref_ids, ref_sc
        except Exception as exc:
            extra["torch_cuda_reference"] = {"error": str(exc)[:300]}
            parity = {"error": str(exc)[:200], "ok": False}
        torch.cuda.empty_cache()

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline(w, args.topk)
    if world == 1 and name == "c5" and not args.no_c2:
        try:
            del graph, out
            for key in ("rowptr", "col", "val", "uw", "iw"):
                w[key] = None
            torch.cuda.empty_cache()
            extra["c2"] = c2_leg(args, dev, flush, torch, hbm_peak)
        except Exception as exc:
            extra["c2"] = {"error": str(exc)[:300]}

    if rank == 0:
        per_rank_bytes = step_bytes / world
        achieved = per_rank_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if world == 1 and os.path.exists(tpath):  # the ncu capture is of the single-GPU launch; N > 1 launches differ
            with open(tpath) as f:
                traffic = json.load(f).get(name)
        line = {
            "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": name, "n_users": nu, "n_items": ni, "train_interactions": w["ne"], "nnz": nnz, "emb": d,
                       "layers": L, "parallelism": parallelism, "l2": "flushed between timed iterations (256 MiB write)",
                       "interactions_per_s": value / 2},
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": traffic, "kernel": "spmm_group_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": step_bytes / L / world, "launches_per_step": L,
                         "note": "achieved = algorithmic (no-reuse) bytes of the L SpMM launches of a step / their CUDA-event time; "
                                 "traffic = ncu dram bytes read+write of one such launch on one GPU; popular rows hit in L2, so "
                                 "achieved can exceed the copy peak",
                         "frac_of_nominal_8TBs": achieved / 8000.0, "per_gpu": world > 1},
            "eval": ev, "clocks": sampler.summary(),
        }
        line.update(extra)
        if workload_note:
            line["config"]["workload_note"] = workload_note
        tref = extra.get("torch_cuda_reference") or {}
        if tref.get("edges_per_s"):
            # the comparator that matters: the reference's own ops (torch.sparse / cuBLAS / topk) on the SAME GPU in the same run;
            # `cpu_baseline` / --impl reference time the reference's CPU path on a bounded sample (throughput per non-zero)
            line["vs_torch_cuda_reference"] = {"propagation": value / tref["edges_per_s"],
                                               "eval": (ev["users_per_s"] / tref["eval_users_per_s"]) if (ev and tref.get("eval_users_per_s")) else None}
        if parity is not None:
            line["parity"] = parity
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    bad = [p for p in (parity, (extra.get("c2") or {}).get("parity")) if p is not None and not p.get("ok", False)]
    if ((extra.get("c2") or {}).get("e2e") or {}).get("matches_device_result") is False:
        bad.append({"c2_e2e_matches_device_result": False})
    if extra.get("item_sharded_matches_user_sharded") is False:
        bad.append({"item_sharded_matches_user_sharded": False})
    if e2e is not None and e2e.get("matches_device_result") is False:
        bad.append({"e2e_matches_device_result": False})
    if bad:   # the line above is still printed, but a result out of tolerance fails the run
        print(f"bench.py: PARITY FAILURE {bad}", file=sys.stderr, flush=True)
        sys.exit(3)


if __name__ == "__main__":
    main()
