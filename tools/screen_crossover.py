"""Where the screened eval path starts to pay: screen vs 3xTF32 on random embeddings over item counts and widths."""
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
    gen = torch.Generator(device=dev).manual_seed(1)
    n_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
    out = []
    for d in (64, 128):
        ue = torch.randn(n_rank, d, generator=gen, device=dev) * 0.3
        for ni in (32768, 65536, 131072, 262144, 524288, 1048576):
            ie = torch.randn(ni, d, generator=gen, device=dev) * (0.1 + torch.rand(ni, 1, generator=gen, device=dev))
            row = {"d": d, "n_items": ni, "n_rank": n_rank}
            for prec in ("screen", "3xtf32"):
                st = {}
                ops.eval_topk(None, ue, ie, 20, precision=prec, stats=st)
                row[prec + "_ms"] = round(timed(lambda: ops.eval_topk(None, ue, ie, 20, precision=prec)), 3)
                if prec == "screen":
                    row["second_pass_rows"] = st["second_pass_rows"]
            row["speedup"] = round(row["3xtf32_ms"] / row["screen_ms"], 3)
            out.append(row)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
