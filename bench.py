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
def cpu_inputs(wl, views_per_step=None):
    """Seeded CPU inputs of one step: (latent, x_prev list, x0 list, jobs per step)."""
    g = torch.Generator().manual_seed(0)
    shape = (1, wl["C"], wl["F"], wl["H"], wl["W"])
    vshape = (1, wl["C"], wl["f_len"], wl["h"], wl["w"])
    if wl["kind"] == "sphere":
        n_views = len(default_views(0)) if views_per_step is None else views_per_step
        n_win = n_views * (len(frame_begins_of(wl, 0)) if wl.get("frame_ring") else 1)
    else:
        n_win = len(ring_windows_of(wl, 0)) if views_per_step is None else views_per_step
    lat = torch.randn(shape, generator=g)
    # payload tensors are shared between windows beyond 64 (memory): the arithmetic per window is unchanged
    pool = min(n_win, 64)
    xp = [torch.randn(vshape, generator=g) for _ in range(pool)]
    x0 = [torch.randn(vshape, generator=g) for _ in range(pool)]
    return lat, [xp[i % pool] for i in range(n_win)], [x0[i % pool] for i in range(n_win)], n_win


def frame_begins_of(wl, i):
    from dynamicscaler_b200.geometry import frame_windows
    return frame_windows(i, frames=wl["f_len"], total_f=wl["F"])


def cpu_one_step(wl, mode, i, lat, xp, x0, views_per_step=None):
    """One gather-all -> merge-all step of oracle/torch_port (the reference's per-call torch op
    sequence) on the host cores.  ``views_per_step``: bounded sample = the first m views / windows."""
    from oracle import torch_port as tp
    if wl["kind"] == "sphere":
        views = default_views(i)[:views_per_step]
        if wl.get("frame_ring"):
            tp.frame_ring_sphere_step(lat, torch.zeros_like(lat), frame_begins_of(wl, i), wl["f_len"], views, xp, x0,
                                      fov=wl["fov"], width=wl["w"], height=wl["h"], mode=mode)
        else:
            tp.sphere_step(lat, torch.zeros_like(lat), views, xp, x0, fov=wl["fov"], width=wl["w"],
                           height=wl["h"], mode=mode)
    else:
        tp.ring_step(lat, torch.zeros_like(lat), ring_windows_of(wl, i)[:views_per_step], xp, x0, w=wl["w"], h=wl["h"])


def cpu_step_time(wl, mode, reps, views_per_step=None, warmup=1):
    """Times ``reps`` steps after ``warmup``.  Returns (seconds per step, cores, sample text, windows/step)."""
    torch.set_num_threads(os.cpu_count() or 1)
    lat, xp, x0, n_win = cpu_inputs(wl, views_per_step)
    ts = []
    for i in range(reps + warmup):
        t = time.perf_counter()
        cpu_one_step(wl, mode, i, lat, xp, x0, views_per_step)
        ts.append(time.perf_counter() - t)
    what = "full steps" if views_per_step is None else f"steps sampled to the first {views_per_step} views/windows"
    return (statistics.mean(ts[warmup:]), torch.get_num_threads(),
            f"{reps} {what} ({n_win} windows each) after {warmup} warm-up", n_win)


def run_reference(args, wl, wl_name):
    """--impl reference: the reference's own CPU implementation of the path (torch port issuing the
    reference's op sequence per call, oracle/torch_port.py) on all host threads; exactly --steps timed
    steps after --warmup, each step a bounded sample of the workload when a full step would make the
    run longer than a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, Wm = max(1, args.steps), max(0, args.warmup)
    budget_s = 150.0
    # size the per-step sample: time one full-size probe of a few views, extrapolate
    probe_views = 4
    sec_probe, _, _, n_probe = cpu_step_time(wl, args.mode, 1, views_per_step=probe_views, warmup=1)
    per_win = sec_probe / n_probe
    full_views = len(default_views(0)) if wl["kind"] == "sphere" else len(ring_windows_of(wl, 0))
    jobs_per_view = (len(frame_begins_of(wl, 0)) if wl.get("frame_ring") else 1)
    fit = int(budget_s / ((K + Wm) * per_win * jobs_per_view))
    m = None if fit >= full_views else max(1, fit)
    sec, cores, sample, n_win = cpu_step_time(wl, args.mode, K, views_per_step=m, warmup=Wm)
    value = n_win / sec
    line = {
        "impl": "reference", "metric": "panorama denoise-step window-blends/sec", "value": value, "unit": "blends/s",
        "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_name, "mode": args.mode, "latent": [1, wl["C"], wl["F"], wl["H"], wl["W"]],
                   "windows_per_step": full_views * jobs_per_view, "window": [wl["h"], wl["w"]]},
        "cpu_baseline": {"value": value, "unit": "blends/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "blends/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if wl["kind"] == "sphere" and args.mode == "bilinear":
        # the mode the reference pipelines call (get_view_tensor_no_interpolate + set_view_tensor_no_interpolation)
        reps = max(2, min(K, 10))
        sec_n, _, sample_n, n_win_n = cpu_step_time(wl, "nearest", reps, views_per_step=m, warmup=1)
        line["variants"] = {"nearest": {"value": n_win_n / sec_n, "unit": "blends/s", "ms_per_step": sec_n * 1e3,
                                        "sample": sample_n}}
    print(json.dumps(line))


# ------------------------------------------------------------------------ our arm ----
class Workset:
    """One rotating buffer set (panoramas + windows + payloads), so consecutive steps touch
    different memory and the aggregate footprint exceeds L2."""

    def __init__(self, wl, n_local, dev, gen, st=None, peer=False, gen_pano=None, dtype=torch.float32, ring_st=None):
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
        # planar ring path on N > 1 GPUs: reference-order replicas in symmetric memory, peer-store exchange
        self.ring_state = ring_st.make_ring_state([self.lat, self.den], mask=self.mask) if ring_st is not None else None
        self.pano = self.state if self.state is not None else (self.ring_state if self.ring_state is not None else self.lat)

    def merge(self, st, plan):
        if self.state is not None:
            st.merge_all(self.state, [self.x_prev, self.x0], plan)
        elif self.ring_state is not None:
            st.merge_all(self.ring_state, [self.x_prev, self.x0], plan)
        else:
            st.merge_all([self.lat, self.den], [self.x_prev, self.x0], plan, mask=self.mask)

    def peer_error(self):
        for s_ in (self.state, self.ring_state):
            if s_ is not None and getattr(s_, "peer_table", None) is not None and s_.peer_error():
                return True
        return False

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


def timed_regions(run, dist_on, regions, dev):
    """``regions`` separately timed executions of ``run`` (each one exactly K steps, bracketed by
    barrier + synchronize); every region's time is the MAX over ranks.  Returns the list (ms)."""
    ms = [timed(run, dist_on) for _ in range(regions)]
    if dist_on:
        t = torch.tensor(ms, device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = [float(v) for v in t.tolist()]
    return ms


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


_SLEEP_CAL = {}


def sleep_cycles_per_us(dev):
    """torch.cuda._sleep spins for a number of SM cycles; calibrate cycles per microsecond once."""
    if dev not in _SLEEP_CAL:
        torch.cuda._sleep(1000)
        torch.cuda.synchronize()
        n = 20_000_000
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        torch.cuda._sleep(n)
        b.record()
        torch.cuda.synchronize()
        _SLEEP_CAL[dev] = n / (a.elapsed_time(b) * 1e3)
    return _SLEEP_CAL[dev]


def dropin_step_time(wl, dev, steps=10):
    """Regime (i) of SURVEY §7: the reference's UNMODIFIED sequential call sites (per view: get latent, get
    mask, mix, set latent / denoised / mask -- t2v_sphere_panorama_pipeline.py:176-291) through the drop-in
    classes, one kernel launch per call, eager Python.  Returns (ms per step, calls per step)."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy, PanoramaTensor
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    H, W, F, C, h, w = wl["H"], wl["W"], wl["F"], wl["C"], wl["h"], wl["w"]
    lat = torch.randn(1, C, F, H, W, device=dev)
    h_lat, h_den = PanoramaLatentProxy(lat), PanoramaLatentProxy(torch.zeros_like(lat))
    xp = torch.randn(1, C, F, h, w, device=dev)
    x0 = torch.randn(1, C, F, h, w, device=dev)
    ones = torch.ones(1, h, w, device=dev)

    def step(i):
        h_msk = PanoramaTensor(torch.zeros(1, H, W, device=dev))          # re-created every step (:173)
        for ph, th in default_views(i):
            vw, _ = h_lat.get_view_tensor_no_interpolate(fov=wl["fov"], theta=th, phi=ph, width=w, height=h)
            mk, _ = h_msk.get_view_tensor_no_interpolate(fov=wl["fov"], theta=th, phi=ph, width=w, height=h)
            vw = mix_latents_with_mask(latent_1=vw, latent_to_add=vw, mask=mk, mix_ratio=0.5)
            h_lat.set_view_tensor_no_interpolation(view_tensor=xp, fov=wl["fov"], theta=th, phi=ph)
            h_den.set_view_tensor_no_interpolation(view_tensor=x0, fov=wl["fov"], theta=th, phi=ph)
            h_msk.set_view_tensor_no_interpolation(view_tensor=ones, fov=wl["fov"], theta=th, phi=ph)

    for i in range(10):
        step(i)                                                           # fills the map caches of the 10 offsets
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / steps * 1e3, len(default_views(0)) * 6


# ---------------------------------------------------------------- multi-GPU parity ----
def parity_check(wl, mode, dev, rank, world, pm, peer, dtype, n_steps, multicast, publish_in_kernel=False):
    """N > 1: run ``n_steps`` steps of the sharded path on fresh, seeded inputs (identical on every
    rank), export every rank's panoramas + mask, and compare their sha256 with a SINGLE-rank run of
    the same steps done locally on every rank.  Payloads depend on the gathered windows, so a wrong
    gather shows up in the state as well.  Returns the ``parity`` object of the JSON line."""
    import hashlib
    import torch.distributed as dist
    from dynamicscaler_b200.step import PanoramaStepper

    shape = (1, wl["C"], wl["F"], wl["H"], wl["W"])
    vshape = (1, wl["C"], wl["f_len"], wl["h"], wl["w"])
    gen = torch.Generator(device=dev).manual_seed(777)                 # same stream on every rank
    lat0 = torch.randn(shape, device=dev, generator=gen).to(dtype)
    plane_mask = wl["kind"] == "sphere" and not wl.get("frame_ring")
    mshape = (1, wl["H"], wl["W"]) if plane_mask else shape

    def drive(st, peer_):
        plans = [make_plan(st, wl, mode, i) for i in range(min(n_steps, n_phases(wl)))]
        n_win = plans[0].n_jobs
        g2 = torch.Generator(device=dev).manual_seed(778)
        base_a = torch.randn((n_win,) + vshape, device=dev, generator=g2).to(dtype)
        base_b = torch.randn((n_win,) + vshape, device=dev, generator=g2).to(dtype)
        lat, den, msk = lat0.clone(), torch.zeros_like(lat0), torch.zeros(mshape, device=dev, dtype=dtype)
        state = st.make_state([lat, den], mask=msk, peer=peer_) if pm else None
        if not pm and peer_:
            state = st.make_ring_state([lat, den], mask=msk)           # planar ring path, peer-store exchange
        for i in range(n_steps):
            plan = plans[i % len(plans)]
            ids = torch.tensor(plan.job_ids, device=dev, dtype=torch.long)
            wins = st.gather_all(state if state is not None else lat, plan)
            xp = wins * 0.5 + base_a[ids] * float(1 + i % 3)
            x0 = wins * 0.25 - base_b[ids]
            if state is not None:
                st.merge_all(state, [xp, x0], plan)
            else:
                st.merge_all([lat, den], [xp, x0], plan, mask=msk)
        outs = (state.export() + [state.export_mask()]) if state is not None else [lat, den, msk]
        torch.cuda.synchronize()
        h = hashlib.sha256()
        for t in outs:
            h.update(t.contiguous().view(torch.uint8).cpu().numpy().tobytes())
        if state is not None and getattr(state, "peer_table", None) is not None:
            assert not state.peer_error(), "a cross-GPU wait timed out during the parity run"
        return h.digest()

    st_n = PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev, rank=rank, world=world)
    st_n.use_multicast = multicast
    st_n.merge_publish_in_kernel = publish_in_kernel
    mine = drive(st_n, peer)
    single = drive(PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev), False)
    both = torch.tensor(list(mine) + list(single), device=dev, dtype=torch.uint8)
    allr = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(allr, both)
    digests = [bytes(t.cpu().tolist()) for t in allr]
    ok = all(d[:32] == digests[0][32:] and d[32:] == digests[0][32:] for d in digests)
    return {"ok": bool(ok), "ranks": world, "steps": n_steps, "mode": mode, "digest": digests[0][32:].hex()[:16],
            "rank_digests": [d[:32].hex()[:16] for d in digests],
            "against": "single-rank run of the same steps (every rank's exported latent, denoised, mask; sha256)"}


# ------------------------------------------------------------------------ our arm ----
def run_ours(args, wl, wl_name):
    import torch.distributed as dist
    from dynamicscaler_b200 import _lib, lib
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
    assert lib.ds_abi_version() == _lib.ABI_VERSION          # fails loudly if the .so is missing

    sampler = ClockSampler(local)
    sampler.start()
    st = PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev, rank=rank, world=world)
    if args.multicast != "auto":
        st.use_multicast = args.multicast == "on"
    if args.merge_publish == "kernel":
        st.merge_publish_in_kernel = True
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    modes, NP = modes_of(wl), n_phases(wl)
    if args.mode not in modes:
        args.mode = modes[0]
    if args.only_mode:
        modes = (args.mode,)
    # host-side plan construction (index maps + merge plans of every distinct window list; static geometry, built
    # once per stage and cached): timed so its cost is visible next to the step time
    t_plan = time.perf_counter()
    plans = {m: [make_plan(st, wl, m, i) for i in range(NP)] for m in modes}
    torch.cuda.synchronize()
    plan_build_ms = (time.perf_counter() - t_plan) * 1e3
    n_local = st.local_jobs(plans[modes[0]][0])[1]
    n_win = plans[modes[0]][0].n_jobs
    st_pm = st if (args.layout == "pm" and wl["kind"] == "sphere") else None
    peer = dist_on and st_pm is not None and args.exchange == "peer"
    ring_peer = dist_on and wl["kind"] == "ring" and args.exchange == "peer"
    ring_st = st if ring_peer else None
    gen_pano = torch.Generator(device=dev).manual_seed(99)      # replicas start identical on every rank
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if args.dtype == "bf16" else 4
    try:
        one = Workset(wl, n_local, dev, gen, st_pm, peer, gen_pano, dtype, ring_st)
    except Exception as e:                                   # no symmetric memory on this system
        if not (peer or ring_peer):
            raise
        print(f"[bench] peer exchange unavailable ({type(e).__name__}: {e}); using the NCCL all-reduce path", file=sys.stderr)
        peer, ring_peer, ring_st = False, False, None
        one = Workset(wl, n_local, dev, gen, st_pm, peer, gen_pano, dtype)
    n_sets = max(2, min(4, -(-3 * L2_BYTES // one.nbytes())))   # aggregate footprint >= 3x L2
    sets = [one] + [Workset(wl, n_local, dev, gen, st_pm, peer, gen_pano, dtype, ring_st) for _ in range(n_sets - 1)]
    combos = [(i % NP, i % n_sets) for i in range(NP * n_sets // np.gcd(NP, n_sets))]

    # --denoise-us T: a busy kernel of T microseconds per LOCAL window between gather and merge, the
    # stand-in for the UNet + ddim_step of BASELINE config 3 (the reference module cannot travel here)
    denoise_cycles = int(args.denoise_us * n_local * sleep_cycles_per_us(dev)) if args.denoise_us > 0 else 0
    # --balance-evals (with --denoise-us, N > 1): T is the time of the TWO CFG evaluations of a window; single uncond
    # evaluations move between ranks exactly as loop.sphere_denoise_loop(balance_evals=True) moves them (the window
    # travels to the receiver, the result back: real NCCL point-to-point transfers inside the step)
    balance = None
    if args.balance_evals and denoise_cycles and dist_on and wl["kind"] == "sphere":
        from dynamicscaler_b200.loop import exchange_windows, plan_eval_balance
        n_views = len(default_views(0))
        n_fw = plans[modes[0]][0].n_fwin
        per_rank = [len(range(r, n_views, world)) * n_fw for r in range(world)]
        moves = plan_eval_balance(per_rank)
        evals = 2 * n_local - sum(1 for d, q, j in moves if d == rank) + sum(1 for d, q, j in moves if q == rank)
        denoise_cycles = int(args.denoise_us / 2 * evals * sleep_cycles_per_us(dev))
        balance = dict(out=[(j, q) for d, q, j in moves if d == rank], inn=[d for d, q, j in moves if q == rank], evals=evals)
        vshape1 = (1, wl["C"], wl["f_len"], wl["h"], wl["w"])
        balance["x_in"] = [torch.empty(vshape1, device=dev, dtype=dtype) for _ in balance["inn"]]
        balance["eu_in"] = [torch.empty(vshape1, device=dev, dtype=dtype) for _ in balance["out"]]

    def make_step(mode, off, s):
        ws, plan = sets[s], plans[mode][off]

        def step():
            st.gather_all(ws.pano, plan, out=ws.wins)
            if balance is not None:                     # donated windows travel to their receivers
                exchange_windows([(ws.wins[j], q) for j, q in balance["out"]],
                                 [(t, d) for t, d in zip(balance["x_in"], balance["inn"])])
            if denoise_cycles:
                torch.cuda._sleep(denoise_cycles)
            if balance is not None:                     # ... and the uncond results back to the owners
                exchange_windows([(t, d) for t, d in zip(balance["x_in"], balance["inn"])],
                                 [(t, q) for t, (j, q) in zip(balance["eu_in"], balance["out"])])
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

    def runner(maker, mode, count):
        fns = [maker(mode, *c) for c in combos]
        try:
            return make_runner(fns, count, use_graph, side), use_graph
        except Exception as e:                             # e.g. a collective that cannot be captured
            torch.cuda.synchronize()
            print(f"[bench] graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            return make_runner(fns, count, False, side), False

    results = {}
    K, Wm, R = args.steps, max(args.warmup, 3), max(args.regions, 1)
    for mode in modes:
        launches_before = lib.launches
        make_step(mode, *combos[0])()                       # count this arm's own kernel launches per step
        launches_per_step = lib.launches - launches_before - (0 if not denoise_cycles else 0)
        run, graphed = runner(make_step, mode, K)
        warm, _ = runner(make_step, mode, Wm)
        warm()
        sampler.load = True
        ms_list = timed_regions(run, dist_on, R, dev)       # R regions of exactly K steps each
        sampler.load = False
        ms = statistics.median(ms_list)
        results[mode] = {"ms_total": ms, "ms_per_step": ms / K, "graph": graphed,
                         "launches_per_step": launches_per_step,
                         "regions_ms_per_step": [round(v / K, 6) for v in ms_list]}
        # per-kernel pass (same rotation) for the roofline
        if not dist_on and not args.no_kernel_pass and not denoise_cycles:
            iters = max(K, 5 * len(combos))
            for name, maker in (("gather", make_gather), ("merge", make_merge)):
                krun, _ = runner(maker, mode, iters)
                krun()
                sampler.load = True
                results[mode][name + "_us"] = statistics.median(timed_regions(krun, False, 5, dev)) / iters * 1e3
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
    # Every step uploads ITS payloads from pinned host memory and downloads ITS windows; the copies
    # are inside the timed region.  The device side of the payloads is double-buffered, so the upload
    # of step i+1 overlaps step i (the uploads are the critical path: PCIe), and the window download
    # runs on its own stream.  The whole e2e_steps-step pipeline (memcpy nodes included) is captured
    # into ONE CUDA graph, so the host launches nothing per step; eager fallback if capture fails.
    mode = args.mode
    ws = sets[0]
    h_xp = torch.randn(ws.x_prev.shape).to(dtype).pin_memory()
    h_x0 = torch.randn(ws.x0.shape).to(dtype).pin_memory()
    h_wins = torch.empty(ws.wins.shape, dtype=dtype).pin_memory()
    up_stream, down_stream = torch.cuda.Stream(), torch.cuda.Stream()
    dev_pay = [(ws.x_prev, ws.x0), (torch.empty_like(ws.x_prev), torch.empty_like(ws.x0))]
    e2e_steps = max(3, min(K, 30))

    def e2e_pipeline(main, n):
        """n steps on streams main / up_stream / down_stream; dependencies only through events
        recorded INSIDE this call (capturable)."""
        ready, free, wins_read = [None, None], [None, None], None
        up_stream.wait_stream(main)
        down_stream.wait_stream(main)
        for i in range(n):
            plan = plans[mode][i % NP]
            b = i & 1
            with torch.cuda.stream(up_stream):              # this step's inputs: pinned host -> device
                if free[b] is not None:
                    up_stream.wait_event(free[b])           # the merge that consumed buffer b is done
                dev_pay[b][0].copy_(h_xp, non_blocking=True)
                dev_pay[b][1].copy_(h_x0, non_blocking=True)
                ready[b] = torch.cuda.Event()
                ready[b].record(up_stream)
            if wins_read is not None:
                main.wait_event(wins_read)                  # the window buffer is free again
            st.gather_all(ws.pano, plan, out=ws.wins)
            gathered = torch.cuda.Event()
            gathered.record(main)
            with torch.cuda.stream(down_stream):            # this step's result: device -> pinned host
                down_stream.wait_event(gathered)
                h_wins.copy_(ws.wins, non_blocking=True)
                wins_read = torch.cuda.Event()
                wins_read.record(down_stream)
            main.wait_event(ready[b])                       # this step's inputs have arrived
            if ws.state is not None:
                st.merge_all(ws.state, list(dev_pay[b]), plan)
            elif ws.ring_state is not None:
                st.merge_all(ws.ring_state, list(dev_pay[b]), plan)
            else:
                st.merge_all([ws.lat, ws.den], list(dev_pay[b]), plan, mask=ws.mask)
            free[b] = torch.cuda.Event()
            free[b].record(main)
        main.wait_stream(down_stream)                       # the last download is part of the job
        main.wait_stream(up_stream)

    e2e_graph = None
    if use_graph:
        try:
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    e2e_pipeline(torch.cuda.current_stream(), e2e_steps)
            e2e_graph = g
        except Exception as e:
            torch.cuda.synchronize()
            print(f"[bench] e2e graph capture failed ({type(e).__name__}: {e}); running the e2e loop eagerly", file=sys.stderr)
    e2e_run = e2e_graph.replay if e2e_graph is not None else (lambda: e2e_pipeline(torch.cuda.current_stream(), e2e_steps))
    e2e_run()                                               # warm-up
    e2e_ms = statistics.median(timed_regions(e2e_run, dist_on, 5, dev)) / e2e_steps
    sampler.stop_flag = True

    parity = None
    if dist_on and not args.no_parity:
        parity = parity_check(wl, args.mode, dev, rank, world, st_pm is not None, peer or ring_peer, dtype,
                              max(4, min(K, 10)), st.use_multicast, st.merge_publish_in_kernel)

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
                       "timing": f"median of {R} separately timed regions of exactly {K} steps (CUDA events, max over ranks)",
                       "layout": "resident pixel-major [H*W][F][4] (csrc/pm.cu)" if st_pm is not None
                                 else "reference [B,C,F,H,W] (csrc/sphere.cu, csrc/ring.cu)",
                       "maps": "precomputed on host (static geometry), cached on device",
                       "denoise_us_per_window": args.denoise_us,
                       "balance_evals": (None if balance is None else
                                         f"{balance['evals']} CFG evaluations on rank 0 after moving single uncond evaluations "
                                         "between ranks (NCCL P2P inside the step)"),
                       "host_plan_build_ms": {"total": plan_build_ms, "plans": NP * len(modes),
                                              "note": "index maps + merge plans of all distinct window lists, built once per "
                                                      "stage before the loop and cached (outside the timed region)"},
                       "sharding": ("none" if not dist_on else
                                    "windows split over ranks; merge kernels push rows into every replica over NVLink ("
                                    + ("NVLS multicast stores" if st.use_multicast else "per-replica TMA row copies")
                                    + ", release/acquire counters)" if peer else
                                    "windows split over ranks; merge kernel stores won elements into every reference-order "
                                    "replica over NVLink, 2 cross-GPU barriers/step" if ring_peer else
                                    "windows split over ranks, 1 NCCL all-reduce/step")},
            "steps_per_s": 1e3 / head["ms_per_step"],
            "regions_ms_per_step": head["regions_ms_per_step"],
            "peer_wait_timeouts": bool((peer or ring_peer) and any(w.peer_error() for w in sets)),
            "gpu_launches": int(head["launches_per_step"] * K),
            "clocks": sampler.summary(),
            "e2e": {"value": n_win / (e2e_ms * 1e-3), "unit": "blends/s",
                    "h2d_bytes_per_step": int((h_xp.numel() + h_x0.numel()) * esz),
                    "d2h_bytes_per_step": int(h_wins.numel() * esz), "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "cuda_graph": e2e_graph is not None},
            "variants": {m: {"ms_per_step": r["ms_per_step"], "blends_per_s": n_win / (r["ms_per_step"] * 1e-3),
                             **{k: v for k, v in r.items() if k.endswith("_us")}} for m, r in results.items()},
        }
        if parity is not None:
            line["parity"] = parity
        if not dist_on and wl["kind"] == "sphere" and not wl.get("frame_ring") and args.dtype == "f32" and not args.no_kernel_pass:
            d_ms, d_calls = dropin_step_time(wl, dev)
            line["variants"]["dropin_sequential"] = {
                "ms_per_step": d_ms, "blends_per_s": n_win / (d_ms * 1e-3), "calls_per_step": d_calls,
                "us_per_call": d_ms * 1e3 / d_calls,
                "what": "the reference's unmodified per-view call sites through the drop-in classes (nearest), eager, "
                        "one launch per call"}
        if not dist_on and not args.no_kernel_pass and not denoise_cycles:
            P = wl["C"]
            plan0 = plans[args.mode][0]
            gb, mb = st.algorithmic_bytes(plan0, P, esz, k=2)
            ngb, nmb, ndetail = st.necessary_bytes(plan0, P, esz, k=2)
            kern = "merge" if head["merge_us"] >= head["gather_us"] else "gather"
            bytes_launch = mb if kern == "merge" else gb
            nec_launch = nmb if kern == "merge" else ngb
            dur = head[kern + "_us"] * 1e-6
            step_s = head["ms_per_step"] * 1e-3
            line["roofline"] = {
                "bound": "hbm", "kernel": f"{kern}_{args.mode}",
                # frac: the bytes the kernel cannot avoid (surviving payload sectors + rows written + maps/plan)
                "achieved": nec_launch / dur / 1e9, "peak": peak, "unit": "GB/s", "frac": nec_launch / dur / 1e9 / peak,
                "necessary_bytes": nec_launch,
                # the SURVEY §8d formula (charges the whole payload of every window, read or not)
                "achieved_algorithmic": bytes_launch / dur / 1e9, "frac_algorithmic": bytes_launch / dur / 1e9 / peak,
                "traffic": ncu_traffic(kern, args.mode, st_pm is not None and args.dtype == "f32"), "peak_source": peak_src,
                "algorithmic_bytes": {"gather": gb, "merge": mb, "coverage": plan0.coverage},
                "necessary": {"gather": ngb, "merge": nmb, **ndetail},
                "step_frac": (ngb + nmb) / step_s / 1e9 / peak,
                "step_frac_algorithmic": (gb + mb) / step_s / 1e9 / peak,
                "all": {m: {kn: {"us": results[m][kn + "_us"],
                                 "algorithmic_GBps": st.algorithmic_bytes(plans[m][0], P, esz)[0 if kn == "gather" else 1]
                                 / (results[m][kn + "_us"] * 1e-6) / 1e9,
                                 "necessary_GBps": st.necessary_bytes(plans[m][0], P, esz)[0 if kn == "gather" else 1]
                                 / (results[m][kn + "_us"] * 1e-6) / 1e9} for kn in ("gather", "merge")}
                        for m in results},
            }
            # context for the fractions: what a PLAIN elementwise copy kernel moving the same necessary byte
            # count reaches in the same harness (CUDA graph, rotating buffers larger than L2) -- at 10-15 us per launch
            # even a memcpy is far from the 6.5 TB/s it reaches on gigabyte transfers
            def plain_copy_us(nbytes):
                half = max(int(nbytes) // 8 * 4, 1 << 20)                           # bytes read = bytes written = half
                n_pairs = max(2, -(-3 * L2_BYTES // (2 * half)))
                pairs = [(torch.empty(half // 4, device=dev), torch.empty(half // 4, device=dev)) for _ in range(n_pairs)]
                # an SM copy kernel: b.copy_(a) would become a driver memcpy node, which only reaches 2.6 TB/s inside
                # a graph at these sizes (tools/calib_copy_sweep.py, profiles/r02e_copy_sweep.jsonl)
                cfns = [(lambda a=a, b=b: torch.add(a, 0.0, out=b)) for a, b in pairs]
                iters = max(K, 5 * len(cfns))
                crun = make_runner(cfns, iters, use_graph, side)
                crun()
                return statistics.median(timed_regions(crun, False, 5, dev)) / iters * 1e3

            try:
                for m in results:
                    nb = st.necessary_bytes(plans[m][0], P, esz)
                    for kn, nbytes in (("gather", nb[0]), ("merge", nb[1])):
                        cu = plain_copy_us(nbytes)
                        line["roofline"]["all"][m][kn]["plain_copy_same_bytes_us"] = cu
                        line["roofline"]["all"][m][kn]["kernel_vs_plain_copy"] = cu / results[m][kn + "_us"]
                dom = line["roofline"]["all"][args.mode][kern]
                line["roofline"]["plain_copy_same_bytes"] = {
                    "us": dom["plain_copy_same_bytes_us"], "GBps": nec_launch / (dom["plain_copy_same_bytes_us"] * 1e-6) / 1e9,
                    "kernel_vs_copy": dom["kernel_vs_plain_copy"],
                    "what": "elementwise SM copy kernel (torch.add(src, 0, out=dst)) moving the same necessary byte count (half "
                            "read, half written) in the same harness; kernel_vs_copy = copy time / kernel time"}
            except Exception as e:                                                  # never let the calibration break the line
                print(f"[bench] plain-copy calibration skipped ({type(e).__name__}: {e})", file=sys.stderr)
            if not args.no_cpu and args.dtype == "f32":
                # bounded sample: about 10-20 s of CPU work on this box's host cores
                full_views = len(default_views(0)) if wl["kind"] == "sphere" else len(ring_windows_of(wl, 0))
                m_views = None
                if wl.get("frame_ring") or wl["H"] >= 512:
                    m_views = 4 if wl.get("frame_ring") else 64
                reps = 8 if args.mode == "bilinear" else 20
                if m_views is not None:
                    reps = 2
                sec, cores, sample, n_cpu = cpu_step_time(wl, args.mode, reps, views_per_step=m_views)
                line["cpu_baseline"] = {"value": n_cpu / sec, "unit": "blends/s", "cores": cores, "kind": "port",
                                        "sample": sample, "ms_per_step": sec * 1e3}
                if wl["kind"] == "sphere" and args.mode == "bilinear" and m_views is None:
                    sec_n, _, sample_n, n_cpu_n = cpu_step_time(wl, "nearest", 10)
                    line["cpu_baseline"]["nearest"] = {"value": n_cpu_n / sec_n, "unit": "blends/s",
                                                       "ms_per_step": sec_n * 1e3, "sample": sample_n}
        print(json.dumps(line))
    rc = 0 if (parity is None or parity["ok"]) else 3
    if dist_on:
        # Leave without tearing NCCL down: destroying a communicator whose collectives were
        # captured in still-alive CUDA graphs can block forever; the JSON line is out, so exit.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(rc)
    if rc:
        sys.exit(rc)


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
    ap.add_argument("--merge-publish", default="host", choices=["kernel", "host"],
                    help="peer exchange ordering: 'host' = one ds_pm_peer_sync launch after every gather / merge; 'kernel' = "
                         "deferred: kernels count themselves as issued, the next kernel publishes the count in its prologue")
    ap.add_argument("--workload", default="cfg2_sphere_128x256", choices=list(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--regions", type=int, default=11,
                    help="separately timed regions of exactly --steps steps; the median is reported")
    ap.add_argument("--denoise-us", type=float, default=0.0,
                    help="busy kernel of this many microseconds per local window between gather and merge "
                         "(UNet + ddim_step stand-in for the scaling study; 0 = the blend-only headline)")
    ap.add_argument("--balance-evals", action="store_true",
                    help="with --denoise-us and N > 1: move single uncond evaluations between ranks (88 = 8 x 11)")
    ap.add_argument("--only-mode", action="store_true", help="time only --mode (skip the other variant)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-vs-single-rank parity run")
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
