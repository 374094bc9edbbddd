"""Host-side control flow of the loop drivers (dynamicscaler_b200/loop.py) on the CPU: grouping, job numbering,
per-window condition callables and the arguments handed to the stepper, with a recording stand-in for
``PanoramaStepper`` (no kernels run; the GPU tests check the numerics)."""
import types

import numpy as np
import torch

from dynamicscaler_b200 import loop as L


class FakePlan:
    def __init__(self, kind, views, fbs, f_len, h, w, rank=0, world=1):
        self.kind, self.h, self.w, self.f_len = kind, h, w, f_len
        nv = len(views)
        local = list(range(rank, nv, world)) if kind != "ring" else None
        if kind == "ring":
            n = nv * len(fbs)
            base, rem = divmod(n, world)
            j0 = rank * base + min(rank, rem)
            self.job_ids = list(range(j0, j0 + base + (1 if rank < rem else 0)))
        else:
            self.job_ids = [s * nv + v for s in range(len(fbs)) for v in local]
        self.n_local = len(self.job_ids)
        self.extra = {"vms": list(views), "n_fwin": len(fbs)}
        self.views, self.fbs = list(views), list(fbs)

    @property
    def n_fwin(self):
        return self.extra["n_fwin"]


class FakeStepper:
    def __init__(self, F, rank=0, world=1):
        self.F, self.rank, self.world, self.device, self.log = F, rank, world, torch.device("cpu"), []

    def plan_sphere(self, views, width, height, fov=120, mode="nearest", frame_begins=None, frames=None):
        fbs = [0] if frame_begins is None else list(frame_begins)
        p = FakePlan("sphere_nearest", views, fbs, self.F if frame_begins is None else frames, height, width, self.rank, self.world)
        self.log.append(("plan", tuple(views), tuple(fbs)))
        return p

    def plan_ring(self, windows, w, h, frame_begins=None, frames=None):
        fbs = [0] if frame_begins is None else list(frame_begins)
        p = FakePlan("ring", windows, fbs, self.F if frame_begins is None else frames, h, w, self.rank, self.world)
        self.log.append(("plan", tuple(windows), tuple(fbs)))
        return p

    def local_jobs(self, plan):
        return (plan.job_ids[0] if plan.job_ids else 0, plan.n_local) if plan.kind == "ring" else (0, plan.n_local)

    def gather_all(self, src, plan, renoise=None, **kw):
        x = torch.zeros(plan.n_local, 1, 4, plan.f_len, plan.h, plan.w)
        self.log.append(("gather", plan.n_local, None if renoise is None else sorted(k for k in renoise if renoise[k] is not None)))
        return x if renoise is None else (x, x if renoise.get("want_raw") else None, x if renoise.get("want_mask") else None)

    def merge_denoised(self, state, x, ec, eu, coef, plan, **kw):
        self.log.append(("merge", x.shape[0], sorted(k for k, v in kw.items() if v is not None)))

    merge_denoised_ref = merge_denoised


SCHED = dict(ddim_timesteps=np.array([1, 251, 501, 751]), alphas_cumprod=np.linspace(0.999, 0.01, 1000).astype(np.float32),
             ddim_alphas=np.array([0.9, 0.7, 0.4, 0.1]), ddim_alphas_prev=np.array([0.999, 0.9, 0.7, 0.4]),
             ddim_sigmas=np.zeros(4), ddim_sqrt_one_minus_alphas=np.sqrt(1 - np.array([0.9, 0.7, 0.4, 0.1])))


def run_sphere(group_size, frame_begins, views, cond, **kw):
    st = FakeStepper(8)
    state = types.SimpleNamespace(shape=(1, 4, 8, 16, 32), dtype=torch.float32, device=torch.device("cpu"), reset_mask=lambda: None)
    seen = []
    model = lambda x, ts, c_crossattn=None, fps=None, **k: seen.append((x.shape[0], ts.numel(), c_crossattn[0].shape[0])) or x
    ids = []
    L.sphere_denoise_step(st, state, views, 12, 8, model=model, t=500, step_i=1, schedule=SCHED, cond=cond,
                          uncond={"c_crossattn": [torch.zeros(1, 7, 3)], "fps": 16}, guidance_scale=7.5, frames=4,
                          group_size=group_size, frame_begins=frame_begins, on_input=lambda i, j, x: ids.append(list(j)), **kw)
    return st.log, ids, seen


def test_sphere_step_groups_and_job_ids():
    views = [(90, 0), (0, 0), (0, 120), (0, 240), (-45, 60)]
    cond = {"c_crossattn": [torch.zeros(1, 7, 3)], "fps": 16}
    # whole step, no frame windows
    log, ids, seen = run_sphere(None, None, views, cond)
    assert ids == [[0, 1, 2, 3, 4]] and seen == [(10, 10 * 4, 10)]
    # groups of 2 views
    log, ids, _ = run_sphere(2, None, views, cond)
    assert ids == [[0, 1], [2, 3], [4]]
    # whole step with frame windows: frame-window-major numbering
    log, ids, _ = run_sphere(None, [1, 3, 5], views, cond)
    assert ids == [list(range(15))] and log[0] == ("plan", tuple(views), (1, 3, 5))
    # groups never span frame windows; the reference's sequential order is f outer, views inner (i2v :284-325)
    log, ids, _ = run_sphere(1, [1, 3, 5], views, cond)
    assert ids == [[n] for n in range(15)]
    assert [e[2] for e in log if e[0] == "plan"] == [(1,)] * 5 + [(3,)] * 5 + [(5,)] * 5
    log, ids, _ = run_sphere(2, [1, 3], views, cond)
    assert ids == [[0, 1], [2, 3], [4], [5, 6], [7, 8], [9]]


def test_sphere_step_condition_callable_and_mixes():
    views = [(90, 0), (0, 0), (0, 120)]
    boxes_seen = []

    def cond_fn(i, job_ids, boxes):
        boxes_seen.append((list(job_ids), list(boxes)))
        return {"c_crossattn": [torch.zeros(len(job_ids), 11, 3)], "fps": 16}

    log, ids, seen = run_sphere(2, [2, 6], views, cond_fn, cfg_batch=False, renoise_ratio=0.6, prev_ratio=0.4,
                                eps_fn=lambda i, j, shape, dev: torch.zeros(shape))
    assert boxes_seen[0] == ([0, 1], [(90, 0, 2), (0, 0, 2)]) and boxes_seen[-1] == ([5], [(0, 120, 6)])
    assert seen[0] == (2, 8, 2) and seen[1] == (2, 8, 2)                     # cond call, then uncond call (expanded context)
    gathers = [e for e in log if e[0] == "gather"]
    assert all("eps" in e[2] and "want_raw" in e[2] and "want_mask" in e[2] for e in gathers)
    assert all(e[2] == ["prev", "prev_mask", "prev_ratio"] for e in log if e[0] == "merge")


def test_ring_step_groups_condition_and_mask():
    st = FakeStepper(8)
    windows, fbs = [(0, 0), (0, 4), (9, 0), (9, 4)], [1, 5]
    lat = torch.zeros(1, 4, 8, 16, 30)
    ids, boxes_seen = [], []

    def cond_fn(i, job_ids, boxes):
        boxes_seen.append(list(boxes))
        return {"c_crossattn": [torch.zeros(len(job_ids), 11, 3)], "fps": 16}

    model = lambda x, ts, c_crossattn=None, fps=None, **k: x
    L.ring_denoise_step(st, [lat, lat.clone()], windows, fbs, 12, 8, 4, model=model, t=500, index=2, step_i=1, schedule=SCHED,
                        cond=cond_fn, uncond=None, guidance_scale=7.5, group_size=3, on_input=lambda i, j, x: ids.append(list(j)),
                        mask=torch.zeros_like(lat), renoise=(0.9, 0.3, 0.6), prev_ratio=0.5,
                        eps_fn=lambda i, j, shape, dev: torch.zeros(shape))
    assert ids == [[0, 1, 2], [3], [4, 5, 6], [7]]
    assert boxes_seen[0] == [(0, 0, 1), (0, 4, 1), (9, 0, 1)] and boxes_seen[-1] == [(9, 4, 5)]
    assert [e[2] for e in st.log if e[0] == "plan"] == [(1,), (1,), (5,), (5,)]
    merges = [e for e in st.log if e[0] == "merge"]
    assert all(set(m[2]) == {"mask", "prev", "prev_mask", "prev_ratio"} for m in merges)
    # whole step as one group
    st2, ids2 = FakeStepper(8), []
    L.ring_denoise_step(st2, [lat, lat.clone()], windows, fbs, 12, 8, 4, model=model, t=500, index=2, step_i=1, schedule=SCHED,
                        cond={"c_crossattn": [torch.zeros(1, 7, 3)], "fps": 16}, uncond=None, guidance_scale=7.5,
                        on_input=lambda i, j, x: ids2.append(list(j)))
    assert ids2 == [list(range(8))] and [e for e in st2.log if e[0] == "merge"][0][2] == []


def test_ring_loop_switches_mixes_off_at_the_last_step():
    st = FakeStepper(8)
    lat = torch.zeros(1, 4, 8, 16, 30)
    mask = torch.ones_like(lat)
    model = lambda x, ts, c_crossattn=None, fps=None, **k: x
    L.ring_denoise_loop(st, [lat, lat.clone()], model=model, schedule=SCHED, schedule_fn=lambda i: ([(0, 0), (9, 4)], [0, 4]),
                        width=12, height=8, frames=4, cond={"c_crossattn": [torch.zeros(1, 7, 3)], "fps": 16}, uncond=None,
                        mask=mask, renoise_ratio=0.6, prev_ratio_list=[0.1, 0.2, 0.3, 0.4],
                        eps_fn=lambda i, j, shape, dev: torch.zeros(shape))
    assert mask.sum().item() == 0                                            # zeroed at every step (i2v :810)
    gathers = [e for e in st.log if e[0] == "gather"]
    merges = [e for e in st.log if e[0] == "merge"]
    assert len(gathers) == 4 and len(merges) == 4
    assert all(g[2] is not None and "eps" in g[2] for g in gathers[:3]) and gathers[3][2] is None      # i < total - 1 (:875)
    assert all("prev" in m[2] for m in merges[:3]) and "prev" not in merges[3][2]


def test_plan_cache_is_bounded():
    from dynamicscaler_b200.step import _PlanCache
    c = _PlanCache(capacity=3)
    for k in range(5):
        c[("ring", k)] = k
    assert list(c) == [("ring", 2), ("ring", 3), ("ring", 4)] and ("ring", 0) not in c and c[("ring", 4)] == 4
    c[("ring", 3)] = 33                                   # re-insert keeps the size
    assert len(c) == 3 and c[("ring", 3)] == 33


def test_sharded_steps_cover_every_job_exactly_once():
    """world_size 2 and 3: the ranks' job ids of every group are disjoint and together make up the step's ordered job list
    (sphere: views dealt round-robin inside each frame window; planar: contiguous blocks), idle ranks call gather / merge
    with zero windows and never call the model."""
    views = [(90, 0), (0, 0), (0, 120), (0, 240), (-45, 60)]
    windows, fbs = [(0, 0), (0, 4), (9, 0), (9, 4)], [1, 5]
    lat = torch.zeros(1, 4, 8, 16, 30)
    for world in (2, 3):
        for group_size in (None, 1, 2):
            seen_sphere, seen_ring = {}, {}
            for rank in range(world):
                st = FakeStepper(8, rank, world)
                state = types.SimpleNamespace(shape=(1, 4, 8, 16, 32), dtype=torch.float32, device=torch.device("cpu"),
                                              reset_mask=lambda: None)
                calls = []
                model = lambda x, ts, c_crossattn=None, fps=None, **k: calls.append(x.shape[0]) or x
                cond = {"c_crossattn": [torch.zeros(1, 7, 3)], "fps": 16}
                ids = []
                L.sphere_denoise_step(st, state, views, 12, 8, model=model, t=500, step_i=0, schedule=SCHED, cond=cond,
                                      uncond=None, guidance_scale=1.0, frames=4, group_size=group_size, frame_begins=[1, 3],
                                      on_input=lambda i, j, x: ids.append(list(j)))
                assert all(n > 0 for n in calls)
                n_groups = len([e for e in st.log if e[0] == "plan"])
                assert len([e for e in st.log if e[0] == "merge"]) == n_groups          # idle ranks still take part in the exchange
                seen_sphere[rank] = [j for g in ids for j in g]
                st = FakeStepper(8, rank, world)
                ids = []
                L.ring_denoise_step(st, [lat, lat.clone()], windows, fbs, 12, 8, 4, model=model, t=500, index=2, step_i=0,
                                    schedule=SCHED, cond=cond, uncond=None, guidance_scale=1.0, group_size=group_size,
                                    on_input=lambda i, j, x: ids.append(list(j)))
                seen_ring[rank] = [j for g in ids for j in g]
            for seen, total in ((seen_sphere, 10), (seen_ring, 8)):
                merged = sorted(j for r in seen.values() for j in r)
                assert merged == list(range(total)), (world, group_size, seen)
