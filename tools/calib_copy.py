"""Calibration: what a plain device copy achieves at the byte counts of one step's kernels
(cold L2, CUDA-graph replay, rotating buffers).  Prints GB/s (read+write) per size."""
import json
import torch

dev = torch.device("cuda:0")
res = {}
for mb in (4, 9, 18, 36, 72, 144, 512):
    n = mb * (1 << 20) // 4
    nset = max(4, (3 * 126) // (2 * mb) + 1)
    a = [torch.randn(n, device=dev) for _ in range(nset)]
    b = [torch.empty(n, device=dev) for _ in range(nset)]
    gs = []
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(nset):
            b[i].copy_(a[i])
        torch.cuda.synchronize()
        for i in range(nset):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                b[i].copy_(a[i])
            gs.append(g)
    for i in range(3 * nset):
        gs[i % nset].replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 40 * nset
    s.record()
    for i in range(iters):
        gs[i % nset].replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / iters * 1e3
    res[f"{mb}MiB"] = {"us": round(us, 2), "GBps_rw": round(2 * mb * (1 << 20) / us / 1e3, 1)}
    del a, b, gs
print(json.dumps(res))
