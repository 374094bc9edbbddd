"""Calibration on the GPU box: (1) what a plain device copy achieves at the byte counts of one
step's kernels, (2) the fixed cost of one kernel node.  20 launches per CUDA graph on rotating
buffers (cold L2), like bench.py.  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_runner, timed  # noqa: E402

dev = torch.device("cuda:0")
side = torch.cuda.Stream()
res = {}
tiny = torch.zeros(32, device=dev)
run = make_runner([lambda: tiny.add_(1.0)] * 20, 400, True, side)
run()
res["tiny_kernel_node_us"] = round(timed(run, False) / 400 * 1e3, 2)
for mb in (4, 9, 18, 36, 72, 144):
    n = mb * (1 << 20) // 4
    nset = max(4, (3 * 126) // (2 * mb) + 1)
    a = [torch.randn(n, device=dev) for _ in range(nset)]
    b = [torch.empty(n, device=dev) for _ in range(nset)]
    fns = [(lambda i=i: b[i].copy_(a[i])) for i in range(nset)]
    iters = 20 * nset
    run = make_runner(fns, iters, True, side)
    run()
    us = timed(run, False) / iters * 1e3
    res[f"copy_{mb}MiB"] = {"us": round(us, 2), "traffic_MB": round(2 * mb * 1.048576, 1),
                            "GBps_rw": round(2 * mb * (1 << 20) / us / 1e3, 1)}
    del a, b, fns, run
for mb in (28, 84, 168):                      # write-only streams (what the gathers mostly are)
    n = mb * (1 << 20) // 4
    nset = max(4, (3 * 126) // mb + 1)
    b = [torch.empty(n, device=dev) for _ in range(nset)]
    fns = [(lambda i=i: b[i].fill_(1.0)) for i in range(nset)]
    iters = 20 * nset
    run = make_runner(fns, iters, True, side)
    run()
    us = timed(run, False) / iters * 1e3
    res[f"fill_{mb}MiB"] = {"us": round(us, 2), "GBps_w": round(mb * (1 << 20) / us / 1e3, 1)}
    del b, fns, run
for mb in (28, 84):                           # read-only streams
    n = mb * (1 << 20) // 4
    nset = max(4, (3 * 126) // mb + 1)
    a = [torch.randn(n, device=dev) for _ in range(nset)]
    o = torch.empty(nset, device=dev)
    fns = [(lambda i=i: torch.sum(a[i], dim=0, out=o[i])) for i in range(nset)]
    iters = 20 * nset
    run = make_runner(fns, iters, True, side)
    run()
    us = timed(run, False) / iters * 1e3
    res[f"sum_{mb}MiB"] = {"us": round(us, 2), "GBps_r": round(mb * (1 << 20) / us / 1e3, 1)}
    del a, fns, run
print(json.dumps(res))
