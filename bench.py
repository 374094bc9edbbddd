#!/usr/bin/env python
"""bench.py -- panorama denoise-step blends/s on B200 (driver contract in the task spec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (torch port)

Workload (BASELINE.json configs[1]): T2V sphere panorama, latent [1,4,16,128,256] fp32,
44 default views (gen_pano_360.py:444-455) of 40x64 at fov 120, longitude offset
(i % 10) * 12 per step.  One STEP = gather_all (every window cut from the step-start
panorama) -> injected payloads (x_prev, pred_x0; stand-ins for UNet + ddim_step) ->
merge_all (latent <- x_prev, denoised <- pred_x0, mask <- 1).  value = steps/s * 44.
The headline mode is "bilinear" (bilinear resample + weighted splat/normalise, the
"window resample + weighted blend" the config names); "nearest" (the variant the reference
pipelines call) is measured in the same run and reported under "variants".
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs[1] -- the headline: T2V sphere panorama, 44 views/step
    "cfg2_sphere_128x256": dict(kind="sphere", C=4, F=16, f_len=16, H=128, W=256, w=64, h=40, fov=120),
    # north-star shape (configs[2] without the UNet)
    "cfg3_sphere_256x512": dict(kind="sphere", C=4, F=16, f_len=16, H=256, W=512, w=64, h=40, fov=120),
    # configs[3]: 64-frame panorama, 13 wrapping frame windows x 44 views = 572 jobs/step
    "cfg4_sphere_f64": dict(kind="sphere", C=4, F=64, f_len=16, H=128, W=256, w=64, h=40, fov=120, frame_ring=True),
    # configs[0] / configs[4]: planar ring shift windows
    "cfg1_ring_64x128": dict(kind="ring", C=4, F=16, f_len=16, H=64, W=128, w=64, h=40, grid="stride8"),
    "cfg5_ring_512x1024": dict(kind="ring", C=4, F=16, f_len=16, H=512, W=1024, w=64, h=40, grid="16x16"),
}
L2_BYTES = 126 * 1024 * 1024


def default_views(i):
    from dynamicscaler_b200.geometry import sphere_views
    return sphere_views(i)


def n_phases(wl):
    return 10 if wl["kind"] == "sphere" else 16


def modes_of(wl):
    return ("bilinear", "nearest") if wl["kind"] == "sphere" else ("ring",)


def ring_windows_of(wl, i):
    """Planar windows of step i (latent units), reference loop order."""
    from dynamicscaler_b200.geometry import ring_windows
    if wl["grid"] == "stride8":      # 16 x 8 windows, stride 8, per-step offset (4, 1)
        ow, oh = (i % 16) * 4, (i % 16) * 1
        return [(ow + 8 * a, oh + 8 * b) for a in range(16) for b in range(8)]
    return ring_windows(i, 16, 16, wl["W"] * 8, wl["H"] * 8)


def make_plan(st, wl, mode, i):
    if wl["kind"] == "ring":
        return st.plan_ring(ring_windows_of(wl, i), wl["w"], wl["h"])
    kw = {}
    if wl.get("frame_ring"):
        from dynamicscaler_b200.geometry import frame_windows
        kw = dict(frame_begins=frame_windows(i, frames=wl["f_len"], total_f=wl["F"]), frames=wl["f_len"])
    return st.plan_sphere(default_views(i), wl["w"], wl["h"], fov=wl["fov"], mode=mode, **kw)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kern, mode, pm):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json, written by tools/summarise_ncu.py); None when not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not (pm and os.path.exists(p)):
        return None
    names = {("gather", "bilinear"): "pm_gather_kernel<float, 4,", ("gather", "nearest"): "pm_gather_kernel<float, 1,",
             ("merge", "bilinear"): "pm_merge_weighted_tiled_kernel<float", ("merge", "nearest"): "pm_merge_nearest_kernel<float"}
    want = names.get((kern, mode))
    for k, v in json.load(open(p)).items():
        if want and k.startswith(want):
            return v["dram_bytes_per_launch"]
    return None


# ------------------------------------------------------------------ clock sampler ----
class ClockSampler(threading.Thread):
    """NVML samples of SM clock + throttle reasons while ``self.load`` is set."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.load, self.stop_flag = [], set(), False, False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv is not None and not self.stop_flag:
            if self.load:
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.reasons |= {n for bit, n in self.REASONS.items() if mask & bit}
                except Exception:
                    pass
            time.sleep(0.002)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ reference arm ----
def cpu_step_time(wl, mode, reps):
    """Times oracle/torch_port.{sphere,ring}_step (the reference's per-call torch op sequence)
    on the host cores.  Returns (seconds per step, cores, sample description, windows/step)."""
    from oracle import torch_port as tp
    if wl.get("frame_ring"):
        return None                                    # no CPU port of the frame-ring composition
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    shape = (1, wl["C"], wl["F"], wl["H"], wl["W"])
    vshape = (1, wl["C"], wl["f_len"], wl["h"], wl["w"])
    n_win = len(default_views(0)) if wl["kind"] == "sphere" else len(ring_windows_of(wl, 0))
    lat = torch.randn(shape, generator=g)
    xp = [torch.randn(vshape, generator=g) for _ in range(n_win)]
    x0 = [torch.randn(vshape, generator=g) for _ in range(n_win)]
    ts = []
    for i in range(reps + 1):
        t = time.perf_counter()
        if wl["kind"] == "sphere":
            tp.sphere_step(lat, torch.zeros_like(lat), default_views(i), xp, x0, fov=wl["fov"], width=wl["w"],
                           height=wl["h"], mode=mode)
        else:
            tp.ring_step(lat, torch.zeros_like(lat), ring_windows_of(wl, i), xp, x0, w=wl["w"], h=wl["h"])
        ts.append(time.perf_counter() - t)
    return (statistics.mean(ts[1:]), torch.get_num_threads(),
            f"{reps} full steps ({n_win} windows each) after 1 warm-up", n_win)


def run_reference(args, wl, wl_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    reps = max(1, min(args.steps, 8 if args.mode == "bilinear" else 20))     # about 10-20 s of CPU work
    res = cpu_step_time(wl, args.mode, reps)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": f"no CPU port of workload {wl_name}"}))
        return
    sec, cores, sample, n_win = res
    value = n_win / sec
    print(json.dumps({
        "impl": "reference", "metric": "panorama denoise-step window-blends/sec", "value": value, "unit": "blends/s",
        "n_gpus": args.gpus, "steps": reps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "mode": args.mode, "latent": [1, wl["C"], wl["F"], wl["H"], wl["W"]],
                   "windows_per_step": n_win, "window": [wl["h"], wl["w"]]},
        "cpu_baseline": {"value": value, "unit": "blends/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "blends/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------ our arm ----
class Workset:
    """One rotating buffer set (panoramas + windows + payloads), so consecutive steps touch
    different memory and the aggregate footprint exceeds L2."""

    def __init__(self, wl, n_local, dev, gen, st=None, peer=False, gen_pano=None, dtype=torch.float32):
        shape = (1, wl["C"], wl["F"], wl["H"], wl["W"])
        vshape = (n_local, 1, wl["C"], wl["f_len"], wl["h"], wl["w"])
        self.lat = torch.randn(shape, device=dev, generator=gen_pano if gen_pano is not None else gen).to(dtype)
        self.den = torch.zeros(shape, device=dev, dtype=dtype)
        # T2V sphere: [1,H,W] mask plane (t2v_sphere_panorama_pipeline.py:159); frame ring (I2V) and the
        # planar ring loops: latent-shaped mask (i2v_sphere_panorama_pipeline.py:220,810)
        plane_mask = wl["kind"] == "sphere" and not wl.get("frame_ring")
        self.mask = torch.zeros((1, wl["H"], wl["W"]) if plane_mask else shape, device=dev, dtype=dtype)
        self.wins = torch.empty(vshape, device=dev, dtype=dtype)
        self.x_prev = torch.randn(vshape, device=dev, generator=gen).to(dtype)
        self.x0 = torch.randn(vshape, device=dev, generator=gen).to(dtype)
        # resident pixel-major state (layout "pm"): gather/merge work on it in place
        self.state = st.make_state([self.lat, self.den], mask=self.mask, peer=peer) if st is not None else None
        self.pano = self.state if self.state is not None else self.lat

    def merge(self, st, plan):
        if self.state is not None:
            st.merge_all(self.state, [self.x_prev, self.x0], plan)
        else:
            st.merge_all([self.lat, self.den], [self.x_prev, self.x0], plan, mask=self.mask)

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.lat, self.den, self.mask, self.wins, self.x_prev, self.x0))


def timed(run, dist_on):
    """CUDA-event time (ms) of one call of ``run`` bracketed by barrier + synchronize."""
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize()
    if dist_on:
        torch.distributed.barrier()
    return a.elapsed_time(b)


def make_runner(fns, K, use_graph, stream):
    """``run()`` executes exactly K steps, step i = fns[i % len(fns)].  With graphs, whole
    cycles of len(fns) consecutive steps are captured into ONE CUDA graph (plus one graph for
    the remainder), so the launch-bound inner loop costs one graph launch per cycle."""
    n = len(fns)
    if not use_graph:
        def run():
            for i in range(K):
                fns[i % n]()
        return run

    def cap(count):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(count):
                fns[i % n]()
        return g

    with torch.cuda.stream(stream):
        full, rem = divmod(K, n)
        g_full = cap(n) if full else None
        g_rem = cap(rem) if rem else None

    def run():
        for _ in range(full):
            g_full.replay()
        if g_rem is not None:
            g_rem.replay()
    return run


def run_ours(args, wl, wl_name):
    import torch.distributed as dist
    from dynamicscaler_b200 import lib
    from dynamicscaler_b200.step import PanoramaStepper

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=dev)
    assert lib.ds_abi_version() == 3                     # fails loudly if the .so is missing

    sampler = ClockSampler(local)
    sampler.start()
    st = PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev, rank=rank, world=world)
    if args.multicast != "auto":
        st.use_multicast = args.multicast == "on"
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    modes, NP = modes_of(wl), n_phases(wl)
    if args.mode not in modes:
        args.mode = modes[0]
    plans = {m: [make_plan(st, wl, m, i) for i in range(NP)] for m in modes}
    n_local = st.local_jobs(plans[modes[0]][0])[1]
    n_win = plans[modes[0]][0].n_jobs
    st_pm = st if (args.layout == "pm" and wl["kind"] == "sphere") else None
    peer = dist_on and st_pm is not None and args.exchange == "peer"
    gen_pano = torch.Generator(device=dev).manual_seed(99)      # replicas start identical on every rank
    if peer:
        lib.check(lib.ds_tune(1, 30_000_000_000))            # DS_TUNE_PEER_TIMEOUT_NS: ranks may start seconds apart
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if args.dtype == "bf16" else 4
    try:
        one = Workset(wl, n_local, dev, gen, st_pm, peer, gen_pano, dtype)
    except Exception as e:                                   # no symmetric memory on this system
        if not peer:
            raise
        print(f"[bench] peer exchange unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce path", file=sys.stderr)
        peer = False
        one = Workset(wl, n_local, dev, gen, st_pm, peer, gen_pano, dtype)
    n_sets = max(2, min(4, -(-3 * L2_BYTES // one.nbytes())))   # aggregate footprint >= 3x L2
    sets = [one] + [Workset(wl, n_local, dev, gen, st_pm, peer, gen_pano, dtype) for _ in range(n_sets - 1)]
    combos = [(i % NP, i % n_sets) for i in range(NP * n_sets // np.gcd(NP, n_sets))]

    def make_step(mode, off, s):
        ws, plan = sets[s], plans[mode][off]

        def step():
            st.gather_all(ws.pano, plan, out=ws.wins)
            ws.merge(st, plan)
        return step

    def make_gather(mode, off, s):
        ws, plan = sets[s], plans[mode][off]
        return lambda: st.gather_all(ws.pano, plan, out=ws.wins)

    def make_merge(mode, off, s):
        ws, plan = sets[s], plans[mode][off]
        return lambda: ws.merge(st, plan)

    # warm-up outside graphs (module load, NCCL communicator setup)
    side = torch.cuda.Stream()
    for mode in modes:
        for c in combos[:3]:
            make_step(mode, *c)()
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    launches_before = lib.launches

    def runner(maker, mode, count):
        fns = [maker(mode, *c) for c in combos]
        try:
            return make_runner(fns, count, use_graph, side), use_graph
        except Exception as e:                             # e.g. a collective that cannot be captured
            torch.cuda.synchronize()
            print(f"[bench] graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            return make_runner(fns, count, False, side), False

    results = {}
    K, Wm = args.steps, max(args.warmup, 3)
    for mode in modes:
        launches_before = lib.launches
        make_step(mode, *combos[0])()                       # count this arm's own kernel launches per step
        launches_per_step = lib.launches - launches_before
        run, graphed = runner(make_step, mode, K)
        warm, _ = runner(make_step, mode, Wm)
        warm()
        sampler.load = True
        ms = timed(run, dist_on)
        sampler.load = False
        if dist_on:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        results[mode] = {"ms_total": ms, "ms_per_step": ms / K, "graph": graphed,
                         "launches_per_step": launches_per_step}
        # per-kernel pass (same rotation) for the roofline
        if not dist_on and not args.no_kernel_pass:
            iters = max(K, 5 * len(combos))
            for name, maker in (("gather", make_gather), ("merge", make_merge)):
                krun, _ = runner(maker, mode, iters)
                krun()
                sampler.load = True
                results[mode][name + "_us"] = timed(krun, False) / iters * 1e3
                sampler.load = False
    # make sure the clock sampler saw the load: replay the headline step for ~0.3 s
    run, _ = runner(make_step, args.mode, 20 * len(combos))
    sampler.load = True
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        run()
        torch.cuda.synchronize()
    sampler.load = False

    # ---- e2e: host buffers in, host buffers out, through the public API ---------------
    # Every step uploads ITS payloads from pinned host memory and downloads ITS windows.  The device
    # side of the payloads is double-buffered, so the upload of step i+1 runs on the copy stream
    # while step i computes (the uploads are the critical path: PCIe), and the window download has
    # its own stream.  Exactly e2e_steps uploads and downloads happen inside the timed region.
    mode = args.mode
    ws = sets[0]
    h_xp = torch.randn(ws.x_prev.shape).to(dtype).pin_memory()
    h_x0 = torch.randn(ws.x0.shape).to(dtype).pin_memory()
    h_wins = torch.empty(ws.wins.shape, dtype=dtype).pin_memory()
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    dev_pay = [(ws.x_prev, ws.x0), (torch.empty_like(ws.x_prev), torch.empty_like(ws.x0))]
    ready = [torch.cuda.Event(), torch.cuda.Event()]       # payload buffer b holds a fresh upload
    free = [torch.cuda.Event(), torch.cuda.Event()]        # the merge that consumed buffer b is done
    wins_read = torch.cuda.Event()                         # the download of the previous windows is done
    e2e_steps = max(3, min(K, 30))
    main = torch.cuda.current_stream()

    def upload(b):
        with torch.cuda.stream(up_stream):
            up_stream.wait_event(free[b])
            dev_pay[b][0].copy_(h_xp, non_blocking=True)
            dev_pay[b][1].copy_(h_x0, non_blocking=True)
            ready[b].record(up_stream)

    def e2e_step(i):
        plan = plans[mode][i % NP]
        b = i & 1
        upload(b ^ 1)                                       # next step's inputs, overlapped with this step
        main.wait_event(wins_read)                          # the window buffer is free again
        st.gather_all(ws.pano, plan, out=ws.wins)
        gathered = torch.cuda.Event()
        gathered.record(main)
        with torch.cuda.stream(down_stream):                # this step's result: device -> pinned host
            down_stream.wait_event(gathered)
            h_wins.copy_(ws.wins, non_blocking=True)
            wins_read.record(down_stream)
        main.wait_event(ready[b])                           # this step's inputs have arrived
        if ws.state is not None:
            st.merge_all(ws.state, list(dev_pay[b]), plan)
        else:
            st.merge_all([ws.lat, ws.den], list(dev_pay[b]), plan, mask=ws.mask)
        free[b].record(main)

    for ev in free + [wins_read]:
        ev.record(main)
    upload(0)
    for i in range(4):
        e2e_step(i)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    a, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(e2e_steps):
        e2e_step(i)
    main.wait_stream(down_stream)                           # the last download is part of the job
    main.wait_stream(up_stream)
    b_ev.record()
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(b_ev) / e2e_steps
    if dist_on:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    sampler.stop_flag = True

    if rank == 0:
        head = results[args.mode]
        value = n_win / (head["ms_per_step"] * 1e-3)
        peak, peak_src = peaks()
        line = {
            "metric": "panorama denoise-step window-blends/sec", "value": value, "unit": "blends/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": wl_name, "mode": args.mode, "latent": [1, wl["C"], wl["F"], wl["H"], wl["W"]],
                       "windows_per_step": n_win, "window": [wl["h"], wl["w"]], "payloads": 2,
                       "l2_policy": f"inputs larger than L2: {n_sets} rotating buffer sets, {n_sets * one.nbytes() / 2**20:.0f} MiB",
                       "cuda_graph": head["graph"], "steps_per_graph": len(combos) if head["graph"] else 0,
                       "layout": "resident pixel-major [H*W][F][4] (csrc/pm.cu)" if st_pm is not None
                                 else "reference [B,C,F,H,W] (csrc/sphere.cu, csrc/ring.cu)",
                       "maps": "precomputed on host (static geometry), cached on device",
                       "sharding": ("none" if not dist_on else
                                    "windows split over ranks; merge kernels push rows into every replica over NVLink ("
                                    + ("NVLS multicast stores" if st.use_multicast else "per-replica TMA row copies")
                                    + ", release/acquire counters)" if peer else
                                    "windows split over ranks, 1 NCCL all-reduce/step")},
            "steps_per_s": 1e3 / head["ms_per_step"],
            "peer_wait_timeouts": bool(peer and any(w.state.peer_error() for w in sets)),
            "gpu_launches": int(head["launches_per_step"] * K),
            "clocks": sampler.summary(),
            "e2e": {"value": n_win / (e2e_ms * 1e-3), "unit": "blends/s",
                    "h2d_bytes_per_step": int((h_xp.numel() + h_x0.numel()) * esz),
                    "d2h_bytes_per_step": int(h_wins.numel() * esz), "ms_per_step": e2e_ms, "steps": e2e_steps},
            "variants": {m: {"ms_per_step": r["ms_per_step"], "blends_per_s": n_win / (r["ms_per_step"] * 1e-3),
                             **{k: v for k, v in r.items() if k.endswith("_us")}} for m, r in results.items()},
        }
        if not dist_on and not args.no_kernel_pass:
            P = wl["C"]
            gb, mb = st.algorithmic_bytes(plans[args.mode][0], P, esz, k=2)
            kern = "merge" if head["merge_us"] >= head["gather_us"] else "gather"
            bytes_launch = mb if kern == "merge" else gb
            dur = head[kern + "_us"] * 1e-6
            line["roofline"] = {
                "bound": "hbm", "kernel": f"{kern}_{args.mode}", "achieved": bytes_launch / dur / 1e9, "peak": peak,
                "unit": "GB/s", "frac": bytes_launch / dur / 1e9 / peak,
                "traffic": ncu_traffic(kern, args.mode, st_pm is not None and args.dtype == "f32"), "peak_source": peak_src,
                "algorithmic_bytes": {"gather": gb, "merge": mb, "coverage": plans[args.mode][0].coverage},
                "step_frac": (gb + mb) / (head["ms_per_step"] * 1e-3) / 1e9 / peak,
                "all": {m: {kn: (st.algorithmic_bytes(plans[m][0], P, esz)[0 if kn == "gather" else 1]
                                 / (results[m][kn + "_us"] * 1e-6) / 1e9) for kn in ("gather", "merge")}
                        for m in results},
            }
            reps = 8 if args.mode == "bilinear" else 20                 # about 10-20 s of CPU work
            cpu = None if (args.no_cpu or args.dtype != "f32") else cpu_step_time(wl, args.mode, reps)
            if cpu is not None:
                sec, cores, sample, _ = cpu
                line["cpu_baseline"] = {"value": n_win / sec, "unit": "blends/s", "cores": cores, "kind": "port",
                                        "sample": sample, "ms_per_step": sec * 1e3}
        print(json.dumps(line))
    if dist_on:
        # Leave without tearing NCCL down: destroying a communicator whose collectives were
        # captured in still-alive CUDA graphs can block forever; the JSON line is out, so exit.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bilinear", choices=["bilinear", "nearest", "ring"],
                    help="sphere workloads: bilinear (headline) or nearest; ring workloads always run 'ring'")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="latent dtype: f32 is the reference's (headline); bf16 halves every byte count (maps stay fp32)")
    ap.add_argument("--layout", default="pm", choices=["pm", "ref"],
                    help="pm: panoramas resident in pixel-major order (PanoramaState); ref: reference order")
    ap.add_argument("--exchange", default="peer", choices=["peer", "allreduce"],
                    help="N > 1: fused peer-store exchange in the merge kernels, or NCCL all-reduce + select")
    ap.add_argument("--multicast", default="auto", choices=["auto", "on", "off"],
                    help="peer exchange: NVLS multicast stores (one store reaches every replica) or per-replica copies")
    ap.add_argument("--workload", default="cfg2_sphere_128x256", choices=list(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-kernel-pass", action="store_true",
                    help="skip the per-kernel timing pass and the roofline (profiling runs: launch list = steps only)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.mode not in modes_of(wl):
        args.mode = modes_of(wl)[0]
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
