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


def _worker(rank, world, port, exchange, out_dir):
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
        H, W, F, h, w = 64, 128, 16, 20, 32
        lat0 = torch.from_numpy(randn(3, 1, 4, F, H, W)).to(dev)
        results = {}
        for mode in ("nearest", "bilinear"):
            ref_st = PanoramaStepper(H, W, F, device=dev)                    # single-GPU reference
            st = PanoramaStepper(H, W, F, device=dev, rank=rank, world=world)
            ref = ref_st.make_state([lat0, torch.zeros_like(lat0)], mask=torch.zeros(1, H, W, device=dev))
            state = st.make_state([lat0, torch.zeros_like(lat0)], mask=torch.zeros(1, H, W, device=dev),
                                  peer=(exchange == "peer"))
            assert (state.peer_table is not None) == (exchange == "peer")
            for i in range(4):
                views = geo.sphere_views(i)[:13]                              # 13 views dealt round-robin: 7 + 6
                plan_ref = ref_st.plan_sphere(views, w, h, fov=120, mode=mode)
                plan = st.plan_sphere(views, w, h, fov=120, mode=mode)
                wins_all = ref_st.gather_all(ref, plan_ref)
                ids = torch.tensor(plan.job_ids, device=dev)                  # this rank's jobs, global numbering
                assert len(plan.job_ids) == plan.n_local == st.local_jobs(plan)[1]
                wins = st.gather_all(state, plan)
                assert torch.equal(wins, wins_all[ids]), (mode, i)
                xp, x0 = wins_all * 0.5 + (i + 1), wins_all * 0.25 - (i + 1)
                ref_st.merge_all(ref, [xp, x0], plan_ref)
                st.merge_all(state, [xp[ids].contiguous(), x0[ids].contiguous()], plan)
            if exchange == "peer":
                # fused denoise epilogue (CFG + ddim_step inside the merge) across ranks
                from dynamicscaler_b200.step import ddim_coefficients
                coef = ddim_coefficients(0.5, 0.55, 0.1, 0.70710677, 7.5)
                gen = torch.Generator(device=dev).manual_seed(11)
                ec, eu, nz = (torch.randn(wins_all.shape, device=dev, generator=gen) for _ in range(3))
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


@pytest.mark.parametrize("exchange", ["peer", "allreduce"])
def test_sharded_step_two_gpus(tmp_path, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import time
    ctx = mp.spawn(_worker, args=(2, _free_port(), exchange, str(tmp_path)), nprocs=2, join=False)
    deadline = time.time() + 240
    while not ctx.join(timeout=5):               # join() returns after the first exit; loop until all are done
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail("multi-GPU workers did not finish in 240 s")
    for r in range(2):
        assert np.load(tmp_path / f"rank{r}.npy")[0], f"rank {r} mismatch"
