"""Plain-copy calibration sweep (run on a B200 through gpurun): time per launch of (a) dst.copy_(src) (the driver's
D2D memcpy node) and (b) an SM copy kernel (torch.add(src, 0, out=dst)) for total byte counts (read + written)
between 10 and 120 MB, in the bench harness (CUDA graph, rotating pairs larger than 3x L2)."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import L2_BYTES, make_runner, timed  # noqa: E402

dev = torch.device("cuda:0")
side = torch.cuda.Stream()
for total_mb in (10, 20, 30, 30.2, 36.6, 40, 40.5, 44.8, 60, 80, 101, 120):
    half = int(total_mb * 1e6 / 2) // 4 * 4
    n_pairs = max(2, -(-3 * L2_BYTES // (2 * half)))
    pairs = [(torch.randn(half // 4, device=dev), torch.empty(half // 4, device=dev)) for _ in range(n_pairs)]
    res = {"total_MB": total_mb, "pairs": n_pairs}
    for name, fn in (("memcpy", lambda a, b: b.copy_(a)), ("kernel", lambda a, b: torch.add(a, 0.0, out=b))):
        fns = [(lambda a=a, b=b: fn(a, b)) for a, b in pairs]
        iters = 5 * len(fns)
        run = make_runner(fns, iters, True, side)
        run()
        us = statistics.median(timed(run, False) for _ in range(5)) / iters * 1e3
        res[name + "_us"] = round(us, 2)
        res[name + "_GBps"] = round(2 * half / us / 1e3)
    print(json.dumps(res), flush=True)
    del pairs
