"""N > 1 path on CPU: world_size-2 ``gloo`` run of the sharded merge protocol --
contiguous job shards, disjoint-support partial merges into zeroed sum / weight buffers, ONE
all-reduce, final select -- with the kernels replaced by their numpy contracts (tests/_emulate.py).
Checks the host logic the NCCL path uses: shard_range, the view_ptr job slicing of the
pixel-major plans, and that the exchanged result equals the single-rank golden bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, out_path):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from _emulate import pm_merge_bilinear, pm_merge_nearest, pm_unpack
    from conftest import randn
    from dynamicscaler_b200 import maps
    from dynamicscaler_b200.step import shard_range
    from oracle import panorama_oracle as po

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, W, F, h, w, f_len = 16, 32, 6, 6, 10, 4
        views = po.default_sphere_views(3)[:9]
        fbs = [0, 2, 4]
        vms = [maps.view_map(120, th, ph, w, h, H, W) for ph, th in views]
        nj = len(views) * len(fbs)
        pay = [randn(900 + q, nj, 4, f_len, h * w) for q in range(2)]      # same on every rank (seeded)
        ftab = maps.frame_table(F, fbs, f_len)
        plan = maps.pm_nearest_plan(vms) if mode == "nearest" else maps.pm_bilinear_plan(vms)
        merge = pm_merge_nearest if mode == "nearest" else pm_merge_bilinear
        j0, n = shard_range(nj, rank, world)
        n_el = H * W * F * 4
        buf = np.zeros(3 * n_el, np.float32)                                 # [sum0 | sum1 | weight]
        sums = [buf[q * n_el:(q + 1) * n_el].reshape(H * W, F, 4) for q in range(2)]
        weight = buf[2 * n_el:].reshape(H * W, F, 4)
        merge(sums, [p[j0:j0 + n] for p in pay], plan, ftab, fbs, len(views), f_len, h * w, weight, 2, j0, n)
        t = torch.from_numpy(buf)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                             # the one exchange of the step
        assert weight.max() == 1.0                                           # disjoint support: one owner each
        base = randn(5, 4, F, H * W)
        out = [np.where(pm_unpack(weight) > 0, pm_unpack(sums[q]), base) for q in range(2)]
        if rank == 0:
            np.savez(out_path, lat=out[0], den=out[1], weight=pm_unpack(weight))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_sharded_merge_world2_gloo(tmp_path, mode):
    sys.path.insert(0, HERE)
    from _emulate import pm_merge_bilinear, pm_merge_nearest, pm_pack, pm_unpack
    from conftest import randn
    from dynamicscaler_b200 import maps
    from oracle import panorama_oracle as po

    out_path = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), mode, out_path), nprocs=2, join=True)
    got = np.load(out_path)

    # single-rank result of the same plan
    H, W, F, h, w, f_len = 16, 32, 6, 6, 10, 4
    views = po.default_sphere_views(3)[:9]
    fbs = [0, 2, 4]
    vms = [maps.view_map(120, th, ph, w, h, H, W) for ph, th in views]
    nj = len(views) * len(fbs)
    pay = [randn(900 + q, nj, 4, f_len, h * w) for q in range(2)]
    ftab = maps.frame_table(F, fbs, f_len)
    base = randn(5, 4, F, H * W)
    want = [pm_pack(base), pm_pack(base)]
    wmask = np.zeros((H * W, F, 4), np.float32)
    if mode == "nearest":
        pm_merge_nearest(want, pay, maps.pm_nearest_plan(vms), ftab, fbs, len(views), f_len, h * w, wmask, 2)
    else:
        pm_merge_bilinear(want, pay, maps.pm_bilinear_plan(vms), ftab, fbs, len(views), f_len, h * w, wmask, 2)
    assert np.array_equal(got["lat"], pm_unpack(want[0]))
    assert np.array_equal(got["den"], pm_unpack(want[1]))
    assert np.array_equal(got["weight"], pm_unpack(wmask))


# ---------------------------------------------------------------- peer-replica exchange ----
_PEER_CASE = dict(H=16, W=32, F=6, h=6, w=10, f_len=4, fbs=[0, 2, 4], n_views=9)


def _peer_case():
    from conftest import randn
    from oracle import panorama_oracle as po
    c = dict(_PEER_CASE)
    c["views"] = po.default_sphere_views(3)[:c["n_views"]]
    nj = c["n_views"] * len(c["fbs"])
    c["pay"] = [randn(700 + q, nj, 4, c["f_len"], c["h"] * c["w"]) for q in range(2)]
    c["base"] = randn(9, 4, c["F"], c["H"] * c["W"])
    return c


def _peer_worker(rank, world, port, mode, out_dir):
    """What step._merge_state does in peer mode, with the kernels replaced by their numpy contracts
    and the NVLink stores by an all_gather: every rank merges ITS views' surviving rows (plans
    built by the real PanoramaStepper.plan_sphere for this rank) and the rows land in every
    replica.  No reduction is involved -- supports are disjoint -- so replicas end identical."""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from _emulate import pm_merge_bilinear, pm_merge_nearest, pm_pack, pm_unpack
    from dynamicscaler_b200.step import PanoramaStepper

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = _peer_case()
        H, W, F, hw, f_len, fbs = c["H"], c["W"], c["F"], c["h"] * c["w"], c["f_len"], c["fbs"]
        st = PanoramaStepper(H, W, F, torch.device("cpu"), rank, world)
        plan = st.plan_sphere(c["views"], c["w"], c["h"], fov=120, mode=mode, frame_begins=fbs, frames=f_len)
        assert plan.n_local == len(plan.job_ids) == len(range(rank, c["n_views"], world)) * len(fbs)
        pm_plan = [t.numpy() for t in plan.pm_merge]
        ftab = plan.d_ftab.numpy()
        pay = [p[plan.job_ids] for p in c["pay"]]                      # the windows this rank denoised
        rows = [np.zeros((H * W, F, 4), np.float32) for _ in range(2)]  # what this rank would push
        touched = np.zeros((H * W, F, 4), np.float32)
        merge = pm_merge_nearest if mode == "nearest" else pm_merge_bilinear
        j0, nj = st.local_jobs(plan)
        merge(rows, pay, pm_plan, ftab, fbs, plan.vpf, f_len, hw, touched, 2, j0, nj)
        mine = torch.from_numpy(np.stack(rows + [touched]))
        everyone = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)                                # stands in for the peer stores
        replica = [pm_pack(c["base"]), pm_pack(c["base"])]
        owners = np.zeros((H * W, F, 4), np.float32)
        for r in range(world):
            got = everyone[r].numpy()
            for q in range(2):
                replica[q] = np.where(got[2] > 0, got[q], replica[q])
            owners += got[2]
        assert owners.max() == 1.0                                      # one writer per (pixel, frame)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), lat=pm_unpack(replica[0]), den=pm_unpack(replica[1]),
                 owners=pm_unpack(owners), job_ids=np.asarray(plan.job_ids))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_peer_replica_exchange_world2_gloo(tmp_path, mode):
    sys.path.insert(0, HERE)
    from _emulate import pm_merge_bilinear, pm_merge_nearest, pm_pack, pm_unpack
    from dynamicscaler_b200 import maps

    mp.spawn(_peer_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(str(tmp_path / f"r{r}.npz")) for r in range(2)]

    c = _peer_case()
    H, W, F, hw, f_len, fbs = c["H"], c["W"], c["F"], c["h"] * c["w"], c["f_len"], c["fbs"]
    vms = [maps.view_map(120, th, ph, c["w"], c["h"], H, W) for ph, th in c["views"]]
    ftab = maps.frame_table(F, fbs, f_len)
    want = [pm_pack(c["base"]), pm_pack(c["base"])]
    wmask = np.zeros((H * W, F, 4), np.float32)
    if mode == "nearest":
        pm_merge_nearest(want, c["pay"], maps.pm_nearest_plan(vms), ftab, fbs, len(vms), f_len, hw, wmask, 2)
    else:
        pm_merge_bilinear(want, c["pay"], maps.pm_bilinear_plan(vms), ftab, fbs, len(vms), f_len, hw, wmask, 2)
    ids = np.sort(np.concatenate([g["job_ids"] for g in got]))
    assert np.array_equal(ids, np.arange(len(vms) * len(fbs)))         # every job on exactly one rank
    for g in got:                                                       # both replicas == single-rank merge
        assert np.array_equal(g["lat"], pm_unpack(want[0]))
        assert np.array_equal(g["den"], pm_unpack(want[1]))
        assert np.array_equal(g["owners"], pm_unpack(wmask))


# ------------------------------------------------ CFG evaluation balancing of the loop driver ----
def test_plan_eval_balance_schedule():
    """88 evaluations of 44 views on 8 ranks become 11 each; schedules are valid for every split."""
    from dynamicscaler_b200.loop import plan_eval_balance
    for per in ([6, 6, 6, 6, 5, 5, 5, 5], [6, 5], [3, 3, 3, 2], [1, 1, 1, 0, 0, 0, 0, 0], [22, 22], [2, 1, 1, 1, 1, 1, 1, 1],
                [7, 7, 7, 6, 6, 6, 6, 6]):
        moves = plan_eval_balance(per)
        evals = [2 * n for n in per]
        seen = set()
        for d, q, j in moves:
            assert d != q and 0 <= j < per[d] and (d, j) not in seen      # an uncond evaluation moves at most once
            seen.add((d, j))
            evals[d] -= 1
            evals[q] += 1
        assert max(evals) - min(evals) <= 1, (per, evals)
    assert plan_eval_balance([6, 6, 6, 6, 5, 5, 5, 5]) == [(0, 7, 5), (1, 6, 5), (2, 5, 5), (3, 4, 5)]


def _balance_worker(rank, world, port, out_dir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types
    from conftest import randn
    from dynamicscaler_b200 import loop

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_views, F = 7, 3                                                  # rank 0 owns 4 views, rank 1 owns 3: 8 / 6 evals -> 7 / 7
        local = list(range(rank, n_views, world))
        stepper = types.SimpleNamespace(world=world, rank=rank, group=None)
        plan = types.SimpleNamespace(extra={"vms": [None] * n_views}, n_fwin=1)
        xb = torch.stack([torch.from_numpy(randn(300 + v, 4, F, 5, 6)) for v in local])
        calls = []

        def model(x, ts, c_crossattn=None, fps=None, **kw):                # elementwise, depends on the context row
            calls.append(x.shape[0])
            assert ts.numel() == (x.shape[0] * F if x.shape[0] > 1 else 1)
            return torch.sin(x) + c_crossattn[0][:, 0, 0].reshape(-1, 1, 1, 1, 1)

        cond = {"c_crossattn": [torch.full((1, 2, 2), 0.25)], "fps": 16}
        uncond = {"c_crossattn": [torch.full((1, 2, 2), -0.5)], "fps": 16}
        ec, eu = loop._balanced_eval(stepper, plan, model, xb, 500, cond, uncond, F, None, True)
        ok = torch.equal(ec, torch.sin(xb) + 0.25) and torch.equal(eu, torch.sin(xb) - 0.5)
        ok = ok and calls == [7]                                           # 4 + 3 resp. 3 + 4 windows in ONE batch on both ranks
        np.save(os.path.join(out_dir, f"balance{rank}.npy"), np.array([ok]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_balanced_eval_world2_gloo(tmp_path):
    """loop._balanced_eval over gloo: one uncond evaluation travels from rank 0 to rank 1 and its result back;
    every window ends up with both branches at its owner, each rank evaluates 7 windows."""
    mp.spawn(_balance_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / f"balance{r}.npy")[0], f"rank {r}"
