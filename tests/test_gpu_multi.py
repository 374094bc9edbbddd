"""Window-sharded multi-GPU step on real GPUs (needs >= 2 CUDA devices; skipped otherwise):
peer-replica push merges (csrc/pm.cu exchange mode over symmetric memory) and the NCCL
all-reduce path must both reproduce the single-GPU result bit for bit on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, exchange, out_dir, multicast=None, dtype_name="float32", frame_ring=False):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from conftest import randn
    from dynamicscaler_b200 import geometry as geo
    from dynamicscaler_b200.step import PanoramaStepper

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        dtype = getattr(torch, dtype_name)
        H, W, F, h, w = 64, 128, (32 if frame_ring else 16), 20, 32
        lat0 = torch.from_numpy(randn(3, 1, 4, F, H, W)).to(dev).to(dtype)
        results = {}
        for mode in ("nearest", "bilinear"):
            ref_st = PanoramaStepper(H, W, F, device=dev)                    # single-GPU reference
            st = PanoramaStepper(H, W, F, device=dev, rank=rank, world=world)
            if multicast is not None:
                st.use_multicast = multicast
            if os.environ.get("DS_TEST_KERNEL_PUBLISH"):                     # in-kernel publication of the merge count
                st.merge_publish_in_kernel = True
            mk_mask = (lambda: torch.zeros_like(lat0)) if frame_ring else (lambda: torch.zeros(1, H, W, device=dev, dtype=dtype))
            ref = ref_st.make_state([lat0, torch.zeros_like(lat0)], mask=mk_mask())
            state = st.make_state([lat0, torch.zeros_like(lat0)], mask=mk_mask(), peer=(exchange == "peer"))
            assert (state.peer_table is not None) == (exchange == "peer")
            if exchange == "peer" and multicast is not None:
                assert state.multicast == multicast, "NVLS multicast alias not available / not honoured"
            fkw = {}
            for i in range(4):
                views = geo.sphere_views(i)[:13]                              # 13 views dealt round-robin
                if frame_ring:                                                # wrapping frame windows (temporal loop)
                    fkw = dict(frame_begins=[(5 * i + 12 * s_) % F for s_ in range(3)], frames=16)
                plan_ref = ref_st.plan_sphere(views, w, h, fov=120, mode=mode, **fkw)
                plan = st.plan_sphere(views, w, h, fov=120, mode=mode, **fkw)
                wins_all = ref_st.gather_all(ref, plan_ref)
                ids = torch.tensor(plan.job_ids, device=dev)                  # this rank's jobs, global numbering
                assert len(plan.job_ids) == plan.n_local == st.local_jobs(plan)[1]
                wins = st.gather_all(state, plan)
                assert torch.equal(wins, wins_all[ids]), (mode, i)
                xp, x0 = (wins_all * 0.5 + (i + 1)).to(dtype), (wins_all * 0.25 - (i + 1)).to(dtype)
                ref_st.merge_all(ref, [xp, x0], plan_ref)
                st.merge_all(state, [xp[ids].contiguous(), x0[ids].contiguous()], plan)
            if exchange == "peer":
                # fused denoise epilogue (CFG + ddim_step inside the merge) across ranks
                from dynamicscaler_b200.step import ddim_coefficients
                coef = ddim_coefficients(0.5, 0.55, 0.1, 0.70710677, 7.5)
                gen = torch.Generator(device=dev).manual_seed(11)
                ec, eu, nz = (torch.randn(wins_all.shape, device=dev, generator=gen).to(dtype) for _ in range(3))
                ref_st.merge_denoised(ref, wins_all, ec, eu, coef, plan_ref, noise=nz)
                st.merge_denoised(state, wins_all[ids].contiguous(), ec[ids].contiguous(), eu[ids].contiguous(), coef,
                                  plan, noise=nz[ids].contiguous())
            a_lat, a_den = state.export()
            b_lat, b_den = ref.export()
            assert not state.peer_error()
            results[mode] = (torch.equal(a_lat, b_lat), torch.equal(a_den, b_den),
                             torch.equal(state.export_mask(), ref.export_mask()))
        ok = all(all(v) for v in results.values())
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([ok]))
        torch.cuda.synchronize()
        dist.barrier()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        ok = False
    finally:
        os._exit(0 if ok else 1)             # no NCCL teardown (see bench.py)


def _run(tmp_path, world, *args):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import time
    ctx = mp.spawn(_worker, args=(world, _free_port(), args[0], str(tmp_path)) + tuple(args[1:]), nprocs=world, join=False)
    deadline = time.time() + 300
    while not ctx.join(timeout=5):               # join() returns after the first exit; loop until all are done
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail("multi-GPU workers did not finish in 300 s")
    for r in range(world):
        assert np.load(tmp_path / f"rank{r}.npy")[0], f"rank {r} mismatch"


@pytest.mark.parametrize("exchange", ["peer", "allreduce"])
def test_sharded_step_two_gpus(tmp_path, exchange):
    _run(tmp_path, 2, exchange)


# Every transport of the fused exchange at every world size the box offers: per-replica TMA row copies
# (multicast off) and NVLS multimem.st stores (multicast on), fp32 and bf16 rows, whole-row and
# frame-ring (per-atom) stores -- each bit-identical to the single-GPU result on every rank.
CASES = [  # (world, multicast, dtype, frame_ring) -- every transport at 2 GPUs, the wide ones where they are the default
    (2, False, "float32", False), (2, True, "float32", False), (2, False, "float32", True), (2, True, "float32", True),
    (2, True, "bfloat16", False), (2, False, "bfloat16", True),
    (4, True, "float32", False), (4, False, "float32", False), (4, True, "bfloat16", True),
    (8, True, "float32", False), (8, False, "float32", True), (8, True, "bfloat16", False),
]


@pytest.mark.parametrize("world,multicast,dtype_name,frame_ring", CASES,
                         ids=[f"w{w}-{'nvls' if m else 'p2p'}-{d}-{'framering' if f else 'rows'}" for w, m, d, f in CASES])
def test_peer_exchange_transports(tmp_path, world, multicast, dtype_name, frame_ring):
    _run(tmp_path, world, "peer", multicast, dtype_name, frame_ring)


@pytest.mark.parametrize("world", [4, 8])
def test_allreduce_fallback_wide(tmp_path, world):
    _run(tmp_path, world, "allreduce")


# ------------------------------------------------ planar ring path, peer-store exchange ----
def _ring_worker(rank, world, port, out_dir, dtype_name="float32"):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from conftest import randn
    from dynamicscaler_b200.step import PanoramaStepper

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        dtype = getattr(torch, dtype_name)
        H, W, F, h, w = 32, 64, 4, 12, 16
        lat0 = torch.from_numpy(randn(5, 1, 4, F, H, W)).to(dev).to(dtype)
        ref_st = PanoramaStepper(H, W, F, device=dev)
        st = PanoramaStepper(H, W, F, device=dev, rank=rank, world=world)
        r_lat, r_den, r_msk = lat0.clone(), torch.zeros_like(lat0), torch.zeros_like(lat0)
        rs = st.make_ring_state([lat0, torch.zeros_like(lat0)], mask=torch.zeros_like(lat0))
        for i in range(4):
            # 4-aligned and unaligned window grids (vector and scalar kernels), wrapping on both axes
            step = 4 if i % 2 == 0 else 3
            windows = [((i * step + 8 * a) % (2 * W - w), (i + 7 * b) % (2 * H - h)) for a in range(8) for b in range(4)]
            if i % 2 == 0:
                windows = [(l // 4 * 4, t) for l, t in windows]
            plan_ref = ref_st.plan_ring(windows, w, h)
            plan = st.plan_ring(windows, w, h)
            wins_all = ref_st.gather_all(r_lat, plan_ref)
            ids = torch.tensor(plan.job_ids, device=dev)
            wins = st.gather_all(rs, plan)
            assert torch.equal(wins, wins_all[ids]), i
            xp, x0 = (wins_all * 0.5 + (i + 1)).to(dtype), (wins_all * 0.25 - (i + 1)).to(dtype)
            ref_st.merge_all([r_lat, r_den], [xp, x0], plan_ref, mask=r_msk)
            st.merge_all(rs, [xp[ids].contiguous(), x0[ids].contiguous()], plan)
        a_lat, a_den = rs.export()
        ok = torch.equal(a_lat, r_lat) and torch.equal(a_den, r_den) and torch.equal(rs.export_mask(), r_msk)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([ok]))
        torch.cuda.synchronize()
        dist.barrier()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        ok = False
    finally:
        os._exit(0 if ok else 1)


def _spawn(tmp_path, world, fn, *args):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    import time
    ctx = mp.spawn(fn, args=(world, _free_port(), str(tmp_path)) + tuple(args), nprocs=world, join=False)
    deadline = time.time() + 300
    try:
        while not ctx.join(timeout=5):
            if time.time() > deadline:
                for p in ctx.processes:
                    p.kill()
                pytest.fail("multi-GPU workers did not finish in 300 s")
    except Exception as e:                       # a worker exited non-zero: show what it recorded
        notes = "".join(open(f).read() for f in sorted(str(x) for x in tmp_path.glob("rank*.txt")))
        pytest.fail(f"{type(e).__name__}: {str(e)[:200]}\n{notes}")
    for r in range(world):
        assert np.load(tmp_path / f"rank{r}.npy")[0], f"rank {r} mismatch"


@pytest.mark.parametrize("world,dtype_name", [(2, "float32"), (2, "bfloat16"), (4, "float32"), (8, "float32")])
def test_ring_peer_exchange(tmp_path, world, dtype_name):
    """Planar ring windows sharded over ranks, merged by peer stores into reference-order replicas
    (ds_merge_nearest_peer): every rank's replica equals the single-GPU result bit for bit."""
    _spawn(tmp_path, world, _ring_worker, dtype_name)


# ------------------------------------------------ denoising-loop driver, sharded ----
def _loop_worker(rank, world, port, out_dir, group_size, balance=False):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from conftest import randn
    from dynamicscaler_b200 import geometry as geo
    from dynamicscaler_b200.loop import sphere_denoise_loop
    from dynamicscaler_b200.step import PanoramaStepper

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        H, W, F, h, w = 64, 128, 4, 16, 24
        lat0 = torch.from_numpy(randn(61, 1, 4, F, H, W)).to(dev)
        ac = np.cumprod(1 - np.linspace(0.00085, 0.012, 1000)).astype(np.float32)
        ts = np.array([1, 251, 501, 751])
        al = ac[ts].astype(np.float64)
        sched = dict(ddim_timesteps=ts, alphas_cumprod=ac, ddim_alphas=al, ddim_alphas_prev=np.append(ac[0], al[:-1]),
                     ddim_sigmas=np.zeros(4), ddim_sqrt_one_minus_alphas=np.sqrt(1 - al))

        def model(x, ts_, c_crossattn=None, fps=None, **kw):          # elementwise: independent of the batch split
            # (no reductions: a mean over the context would depend on the reduction shape, i.e. on the batch size)
            return torch.sin(x * 1.7) * 0.5 + c_crossattn[0][:, 3, 5].reshape(-1, 1, 1, 1, 1) * 0.1

        def eps_fn(i, job_ids, shape, device):
            if not len(job_ids):
                return torch.empty(shape, device=device)
            return torch.stack([torch.from_numpy(randn(7000 + 50 * i + j, *shape[1:])).to(device) for j in job_ids])

        cond = {"c_crossattn": [torch.from_numpy(randn(62, 1, 77, 16)).to(dev)], "fps": 16}
        uncond = {"c_crossattn": [torch.from_numpy(randn(63, 1, 77, 16)).to(dev)], "fps": 16}
        outs = []
        for (r, wsz, peer) in ((0, 1, False), (rank, world, True)):
            st = PanoramaStepper(H, W, F, device=dev, rank=r, world=wsz)
            state = st.make_state([lat0, torch.zeros_like(lat0)], mask=torch.zeros(1, H, W, device=dev), peer=peer)
            sphere_denoise_loop(st, state, model=model, schedule=sched, views_fn=lambda i: geo.sphere_views(i)[:11],
                                width=w, height=h, frames=F, cond=cond, uncond=uncond, guidance_scale=7.5,
                                group_size=group_size, renoise_ratio=0.6, prev_ratio_list=[0.3, 0.4, 0.5, 0.6], eps_fn=eps_fn,
                                balance_evals=balance and peer)
            outs.append(state.export() + [state.export_mask()])
        ok = all(torch.equal(a, b) for a, b in zip(*outs))
        if not ok:
            with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
                for name, a, b in zip(("latent", "denoised", "mask"), *outs):
                    d = (a.float() - b.float()).abs()
                    fh.write(f"rank {rank} {name}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e}\n")
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([ok]))
        torch.cuda.synchronize()
        dist.barrier()
    except BaseException:
        import traceback
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "a") as fh:
            fh.write(traceback.format_exc())
        ok = False
    finally:
        os._exit(0 if ok else 1)


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_denoise_loop_balanced_evals(tmp_path, world):
    """The same with single uncond evaluations moved between ranks (11 views on 2 ranks: 12 / 10 -> 11 / 11
    evaluations; NCCL point-to-point): still bit-identical to the single-GPU loop."""
    _spawn(tmp_path, world, _loop_worker, None, True)


@pytest.mark.parametrize("world,group_size", [(2, None), (2, 3), (4, None), (8, None)])
def test_sharded_denoise_loop(tmp_path, world, group_size):
    """4 DDIM steps of the loop driver (re-noise mix + merge_prev prologue + CFG + ddim_step fused into the
    gather / merge launches), views dealt over the ranks, UNet stand-in evaluated on local windows only:
    every rank's replica equals the single-GPU loop bit for bit."""
    _spawn(tmp_path, world, _loop_worker, group_size)


# ------------------------------------------------ planar shift-window loop, sharded ----
def _ring_loop_worker(rank, world, port, out_dir, exchange, dtype_name="float32", i2v=False):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from conftest import randn
    from dynamicscaler_b200.loop import i2v_planar_schedule, ring_denoise_loop, shift_window_schedule
    from dynamicscaler_b200.step import PanoramaStepper

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    try:
        dtype = getattr(torch, dtype_name)
        frames, h, w, nf, nh, nw = 4, 8, 12, 2, 2, 3
        F, H, W = frames * nf, h * nh, w * nw
        if i2v:                                  # overlapping windows on W, H and F (the I2V planar loop)
            h, w, H, W, nw, nh = 8, 16, 16, 48, 3, 3
            i2v_kw = dict(width=128, height=64, frames=frames, total_w=384, total_h=128, total_f=F, num_windows_w=nw,
                          num_windows_h=nh, loop_step=4, overlap_ratio_f=0.5, loop_step_frame=2, dock_at_f=True,
                          begin_index_offset=1)
        lat0 = torch.from_numpy(randn(71, 1, 4, F, H, W)).to(dev).to(dtype)
        ac = np.cumprod(1 - np.linspace(0.00085, 0.012, 1000)).astype(np.float32)
        ts = np.array([1, 251, 501, 751])
        al = ac[ts].astype(np.float64)
        sched = dict(ddim_timesteps=ts, alphas_cumprod=ac, ddim_alphas=al, ddim_alphas_prev=np.append(ac[0], al[:-1]),
                     ddim_sigmas=np.full(4, 0.02), ddim_sqrt_one_minus_alphas=np.sqrt(1 - al))
        kw = dict(frames=frames, latent_height=h, latent_width=w, num_windows_f=nf, num_windows_h=nh, num_windows_w=nw,
                  loop_step=4, shift_jump_odd_w=True, docking_w=True, docking_f=True, docking_step_range=[1, 3])

        def model(x, ts_, c_crossattn=None, fps=None, **kw_):         # elementwise: independent of the batch split
            return (torch.sin(x.float() * 1.7) * 0.5 + c_crossattn[0][:, 3, 5].reshape(-1, 1, 1, 1, 1) * 0.1).to(x.dtype)

        def noise_fn(i, job_ids, shape, device):
            if not len(job_ids):
                return torch.empty(shape, device=device, dtype=dtype)
            return torch.stack([torch.from_numpy(randn(8000 + 50 * i + j, *shape[1:])).to(device) for j in job_ids]).to(dtype)

        cond = {"c_crossattn": [torch.from_numpy(randn(72, 1, 77, 16)).to(dev)], "fps": 16}
        uncond = {"c_crossattn": [torch.from_numpy(randn(73, 1, 77, 16)).to(dev)], "fps": 16}
        def eps_fn(i, job_ids, shape, device):
            if not len(job_ids):
                return torch.empty(shape, device=device, dtype=dtype)
            return torch.stack([torch.from_numpy(randn(8800 + 50 * i + j, *shape[1:])).to(device) for j in job_ids]).to(dtype)

        def cond_fn(i, job_ids, boxes):          # one context row per window, a function of the window position
            base = cond["c_crossattn"][0]
            rows = [base + 0.01 * (l + 3 * t + 7 * fb) for (l, t, fb) in boxes]
            return {"c_crossattn": [torch.cat(rows, dim=0)], "fps": 16}

        outs = []
        for (r, wsz) in ((0, 1), (rank, world)):
            st = PanoramaStepper(H, W, F, device=dev, rank=r, world=wsz)
            target = [lat0.clone(), torch.zeros_like(lat0)]
            mask = torch.zeros_like(lat0) if i2v else None
            if wsz > 1 and exchange == "peer":
                target = st.make_ring_state(target, mask=mask)
            if i2v:     # whole-step groups and 4-window groups on alternating steps, all mixes on
                ring_denoise_loop(st, target, model=model, schedule=sched, schedule_fn=lambda i: i2v_planar_schedule(i, **i2v_kw),
                                  width=w, height=h, frames=frames, cond=cond_fn, uncond=uncond, guidance_scale=7.5,
                                  group_size=lambda i: 4 if i % 2 else None, noise_fn=noise_fn, mask=mask, renoise_ratio=0.6,
                                  prev_ratio_list=[0.3, 0.4, 0.5, 0.6], eps_fn=eps_fn)
            else:       # whole-step groups on the steps whose windows tile the torus, 2 windows at a time on the docked ones
                ring_denoise_loop(st, target, model=model, schedule=sched, schedule_fn=lambda i: shift_window_schedule(i, **kw),
                                  width=w, height=h, frames=frames, cond=cond, uncond=uncond, guidance_scale=7.5,
                                  group_size=lambda i: 2 if i in (1, 3) else None, noise_fn=noise_fn)
            outs.append(target.export() if hasattr(target, "export") else target)
        ok = all(torch.equal(a, b) for a, b in zip(*outs)) and not torch.equal(outs[0][0], lat0)
        if not ok:
            with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
                for name, a, b in zip(("latent", "denoised"), *outs):
                    d = (a.float() - b.float()).abs()
                    fh.write(f"rank {rank} {name}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e}\n")
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array([ok]))
        torch.cuda.synchronize()
        dist.barrier()
    except BaseException:
        import traceback
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "a") as fh:
            fh.write(traceback.format_exc())
        ok = False
    finally:
        os._exit(0 if ok else 1)


@pytest.mark.parametrize("world,exchange,dtype_name", [(2, "peer", "float32"), (2, "allreduce", "float32"),
                                                        (2, "peer", "bfloat16"), (4, "peer", "float32"), (8, "peer", "float32")])
def test_sharded_ring_denoise_loop(tmp_path, world, exchange, dtype_name):
    """4 DDIM steps of the planar shift-window loop (ring gather -> stand-in UNet on the local windows -> nearest merge
    fused with CFG + ddim_step incl. eta > 0 noise, ds_merge_nearest_ddim), windows split over the ranks, docked steps
    with overlapping windows: every rank's replica equals the single-GPU loop bit for bit -- through peer stores
    into reference-order replicas and through the zero-buffer all-reduce."""
    _spawn(tmp_path, world, _ring_loop_worker, exchange, dtype_name)


@pytest.mark.parametrize("world,exchange,dtype_name", [(2, "peer", "float32"), (2, "allreduce", "float32"),
                                                        (2, "peer", "bfloat16"), (4, "peer", "float32"), (8, "peer", "float32")])
def test_sharded_i2v_ring_denoise_loop(tmp_path, world, exchange, dtype_name):
    """The planar I2V loop sharded: overlapping windows on W, H and F with frame docking, latent-shaped mask panorama in
    the replicas, re-noise mix on the read side (mask windows gathered before the cross-GPU barrier), merge_prev mix
    inside the fused merge, one context row per window: every rank's replica equals the single-GPU loop bit for bit."""
    _spawn(tmp_path, world, _ring_loop_worker, exchange, dtype_name, True)
