"""One fused training step at the headline scale (c5: 12 M nodes x 128, 400 M nnz): does it fit, what does it cost, where does memory peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from textgcn_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    w = bench.build_workload("c5", dev)
    graph = ops.Graph(w["nu"], w["ni"], w["rowptr"], w["col"], w["val"])
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    torch.cuda.reset_peak_memory_stats()
    res = bench.train_leg(w, graph, dev, flush, torch, batch=2048, steps=3, warmup=1)
    res["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
