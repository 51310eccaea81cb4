"""LTR-shaped ranking call (K = 1600 + bias terms, 63 k items): the streamed screened variant against 3xTF32, rows sent to the second pass."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from textgcn_b200 import ops  # noqa: E402


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(5)
    n_rank, n_items, K, k = int(sys.argv[1]) if len(sys.argv) > 1 else 18944, 63000, 1600, 20
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    ue = torch.randn(n_rank, K, generator=gen, device=dev)
    ie = torch.randn(n_items, K, generator=gen, device=dev) * scale * 0.02
    ub = torch.rand(n_rank, generator=gen, device=dev)
    ib = torch.rand(n_items, generator=gen, device=dev) * 0.1
    res = {"n_rank": n_rank, "n_items": n_items, "K": K}
    for prec in ("screen", "3xtf32"):
        st = {}
        ops.eval_topk(None, ue, ie, k, user_bias=ub, item_bias=ib, by_position=True, precision=prec, stats=st)
        res[prec] = {"ms": round(timed(lambda: ops.eval_topk(None, ue, ie, k, user_bias=ub, item_bias=ib, by_position=True, precision=prec)), 3),
                     "second_pass_rows": st["second_pass_rows"]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
