"""Planar ring merge with the denoise epilogue fused (ds_merge_nearest_ddim) against the unfused chain on
reference-order tensors: CFG + ddim_step as torch elementwise ops producing x_prev / pred_x0, then merge_all.
BASELINE configs 1 and 5 (ring windows 40x64x16, fp32).  One JSON line per workload on stdout.

    python tools/ring_ddim_bench.py [--iters 30]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dynamicscaler_b200.step import PanoramaStepper, ddim_coefficients  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    for name in ("cfg1_ring_64x128", "cfg5_ring_512x1024"):
        wl = bench.WORKLOADS[name]
        st = PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev)
        plan = st.plan_ring(bench.ring_windows_of(wl, 1), wl["w"], wl["h"])
        gen = torch.Generator(device=dev).manual_seed(0)
        lat = torch.randn((1, wl["C"], wl["F"], wl["H"], wl["W"]), device=dev, generator=gen)
        den = torch.zeros_like(lat)
        x = st.gather_all(lat, plan)
        ec, eu = (torch.randn(x.shape, device=dev, generator=gen) for _ in range(2))
        coef = ddim_coefficients(0.5, 0.6, 0.0, 0.5 ** 0.5, 7.5)
        sa, s1m, sp, sd, sig, g = coef

        def unfused():
            e = eu + g * (ec - eu)
            x0 = (x - s1m * e) / sa
            xp = sp * x0 + sd * e
            st.merge_all([lat, den], [xp, x0], plan)

        xp_ready, x0_ready = torch.randn_like(x), torch.randn_like(x)
        t_unfused = timed(unfused, args.iters)
        t_merge_only = timed(lambda: st.merge_all([lat, den], [xp_ready, x0_ready], plan), args.iters)
        t_fused = timed(lambda: st.merge_denoised_ref([lat, den], x, ec, eu, coef, plan), args.iters)
        V = x.numel() * 4
        surv = plan.coverage * lat.numel() * 4
        print(json.dumps({"workload": name, "windows": plan.n_jobs, "window_tensor_bytes": V, "surviving_bytes": int(surv),
                          "unfused_torch_ddim_plus_merge_us": round(t_unfused, 1), "merge_only_us": round(t_merge_only, 1),
                          "fused_merge_ddim_us": round(t_fused, 1),
                          "fused_necessary_GBps": round((3 * surv + 2 * surv) / t_fused / 1e3, 1),
                          "note": "fused reads x/ec/eu for surviving elements only and writes 2 panoramas; the unfused chain "
                                  "streams every window element through 7 elementwise kernels first"}))


if __name__ == "__main__":
    main()
