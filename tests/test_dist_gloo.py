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
