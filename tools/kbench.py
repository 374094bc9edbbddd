"""Kernel-level autotuning bench for the pixel-major path (run on a B200 through gpurun).
Times each of the four cfg2 kernels alone (CUDA graph of 20 launches on rotating buffer
sets, cold L2) for a list of ds_tune knob settings and prints one JSON line per setting."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, Workset, make_plan, make_runner, modes_of, n_phases, timed  # noqa: E402
from dynamicscaler_b200 import lib  # noqa: E402
from dynamicscaler_b200.step import PanoramaStepper  # noqa: E402

GATHER_CPS = 0          # DS_TUNE_PM_GATHER_CTAS_PER_SM


def main():
    wl_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_sphere_128x256"
    wl = WORKLOADS[wl_name]
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    st = PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    modes, NP = modes_of(wl), n_phases(wl)
    plans = {m: [make_plan(st, wl, m, i) for i in range(NP)] for m in modes}
    n_jobs = plans[modes[0]][0].n_jobs
    n_sets = 4 if n_jobs < 100 else 2
    sets = [Workset(wl, n_jobs, dev, gen, st if wl["kind"] == "sphere" else None) for _ in range(n_sets)]
    combos = [(i % NP, i % n_sets) for i in range(2 * NP)]
    side = torch.cuda.Stream()

    def t_kernel(kind, mode, iters=200):
        fns = []
        for off, s in combos:
            ws, plan = sets[s], plans[mode][off]
            if kind == "gather":
                fns.append(lambda ws=ws, plan=plan: st.gather_all(ws.pano, plan, out=ws.wins))
            else:
                fns.append(lambda ws=ws, plan=plan: ws.merge(st, plan))
        run = make_runner(fns, iters, True, side)
        run()
        return timed(run, False) / iters * 1e3

    settings = [{"gather_cps": 0, "pdl": m} for m in (11, 0, 15)]
    if os.environ.get("KBENCH_DEFAULT_ONLY"):
        settings = settings[:1]
    if os.environ.get("KBENCH_SETTINGS"):                                 # JSON list of {gather_cps, pdl, pipe}
        settings = json.loads(os.environ["KBENCH_SETTINGS"])
    for cfg in settings:
        lib.check(lib.ds_tune(GATHER_CPS, cfg.get("gather_cps", 0)))
        lib.check(lib.ds_tune(2, cfg.get("pdl", 11)))                     # DS_TUNE_PDL
        lib.check(lib.ds_tune(3, cfg.get("pipe", 2)))                     # DS_TUNE_PM_GATHER_PIPELINE
        lib.check(lib.ds_tune(4, cfg.get("hw_const", 1)))                 # DS_TUNE_PM_HW_CONST
        res = dict(cfg, workload=wl_name)
        for mode in modes:                                  # whole steps (gather -> merge chains)
            fns = []
            for off, s_ in combos:
                ws, plan = sets[s_], plans[mode][off]

                def full(ws=ws, plan=plan):
                    st.gather_all(ws.pano, plan, out=ws.wins)
                    ws.merge(st, plan)
                fns.append(full)
            it = 200 if n_jobs < 100 else 40
            run = make_runner(fns, it, True, side)
            run()
            res[f"step_{mode}_us"] = round(timed(run, False) / it * 1e3, 2)
        for mode in modes:
            for kind in ("gather", "merge"):
                us = t_kernel(kind, mode, iters=200 if n_jobs < 100 else 40)
                gb, mb = st.algorithmic_bytes(plans[mode][0], 4, 4, k=2)
                res[f"{kind}_{mode}_us"] = round(us, 2)
                res[f"{kind}_{mode}_GBps"] = round((gb if kind == "gather" else mb) / us / 1e3, 1)
        if wl["kind"] == "sphere" and wl["f_len"] == 16 and not os.environ.get("KBENCH_NO_DENOISE"):
            # ddim_step + CFG fused into the merge vs the same chain as torch ops + merge_all
            from dynamicscaler_b200.step import ddim_coefficients
            coef = ddim_coefficients(0.5, 0.55, 0.0, 0.7071, 7.5)
            eps_c = [torch.randn_like(ws.x0) for ws in sets]
            eps_u = [torch.randn_like(ws.x0) for ws in sets]
            sa, s1m, sp, sd, sig, g = coef
            for mode in modes:
                fused, unfused = [], []
                for off, s_ in combos:
                    ws, plan = sets[s_], plans[mode][off]
                    fused.append(lambda ws=ws, plan=plan, s_=s_: st.merge_denoised(ws.state, ws.wins, eps_c[s_], eps_u[s_], coef, plan))

                    def chain(ws=ws, plan=plan, s_=s_):
                        e = eps_u[s_] + g * (eps_c[s_] - eps_u[s_])
                        x0 = (ws.wins - s1m * e) / sa
                        xp = sp * x0 + sd * e
                        st.merge_all(ws.state, [xp, x0], plan)
                    unfused.append(chain)
                for name, fns in (("fused", fused), ("torch_chain", unfused)):
                    it = 200 if n_jobs < 100 else 40
                    run = make_runner(fns, it, True, side)
                    run()
                    res[f"denoise_merge_{name}_{mode}_us"] = round(timed(run, False) / it * 1e3, 2)
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
