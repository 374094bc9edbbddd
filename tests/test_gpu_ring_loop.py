"""The planar shift-window denoising loop (dynamicscaler_b200/loop.py ring_denoise_loop) and the nearest
merge fused with CFG + ddim_step on reference-order tensors (ds_merge_nearest_ddim) on a B200
(``pytest -m gpu``):

* replay of the REAL reference run (tests/golden/ring_loop.npz, made by tests/golden/make_golden_ring_loop.py
  from ``VC2_Pipeline_T2V.basic_sample_shift_multi_windows``): the driver must feed the UNet bit-identical
  windows and land on bit-identical panoramas after EVERY step -- one window at a time, and with one batched
  UNet call per step wherever the step's windows tile the torus;
* the same run through the DROP-IN ``RingLatent`` class, call for call like the reference's loop body;
* the fused merge against ddim_step on the host (oracle) followed by the plain merge, for ring and
  nearest-sphere plans, vector and scalar paths, frame rings, masks, eta > 0, 16-bit dtypes.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, randn
from oracle import panorama_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dynamicscaler_b200 import _lib, lib
    assert lib.ds_abi_version() == _lib.ABI_VERSION
    return torch.device("cuda:0")


def G(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def load_ring_loop():
    g = np.load(os.path.join(GOLDEN, "ring_loop.npz"))
    frames, h, w, nw, nh, nf, loop_step, steps = [int(v) for v in g["cfg"]]
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    dock = [int(v) for v in g["dock_steps"]]
    kw = dict(frames=frames, latent_height=h, latent_width=w, num_windows_f=nf, num_windows_h=nh, num_windows_w=nw,
              loop_step=loop_step, shift_jump_odd_w=True, docking_w=True, docking_step_range=dock)
    counts = [len(po.shift_window_jobs(i, **kw)) for i in range(steps)]
    base = [0] + [int(v) for v in np.cumsum(counts)]
    return g, (frames, h, w, nw, nh, nf, steps), sched, kw, dock, base


@pytest.mark.parametrize("order", ["sequential", "batched"])
def test_ring_loop_replays_the_real_reference_pipeline(dev, order):
    """3 DDIM steps of the reference's own planar loop (12, 12 and -- docked, overlapping -- 16 windows): every UNet
    input (sha256) and the latent / denoised panoramas after every step equal the real pipeline's, bit for bit.
    ``batched``: one UNet call for the whole step on the steps whose windows tile the torus, one window at a time
    on the docked step."""
    from dynamicscaler_b200.loop import ring_denoise_loop, shift_window_schedule
    from dynamicscaler_b200.step import PanoramaStepper
    g, (frames, h, w, nw, nh, nf, steps), sched, kw, dock, base = load_ring_loop()
    F, H, W = frames * nf, h * nh, w * nw
    st = PanoramaStepper(H, W, F, device=dev)
    lat = G(randn(51, 1, 4, F, H, W), dev)
    den = torch.zeros_like(lat)
    cur = {"calls": 0}

    def on_input(i, job_ids, xb):
        cur.update(i=i, jobs=list(job_ids), branch=0)
        for q, job in enumerate(job_ids):
            assert np.array_equal(sha(N(xb[q:q + 1])), g["unet_in_sha"][base[i] + job]), (i, job)

    def replay_model(x, ts, c_crossattn=None, fps=None, **kw_):
        n = len(cur["jobs"])
        assert x.shape[0] == n and ts.numel() == (n * frames if n > 1 else 1)
        assert kw_["temporal_length"] == frames and kw_["clean_cond"] is True and fps == 16
        want_ctx = g["text_emb"] if cur["branch"] == 0 else g["uncond_emb"]
        assert np.array_equal(N(c_crossattn[0][:1]), want_ctx)
        src = g["eps_cond"] if cur["branch"] == 0 else g["eps_uncond"]
        cur["branch"] += 1
        cur["calls"] += n
        return torch.cat([G(src[base[cur["i"]] + job], dev) for job in cur["jobs"]], dim=0)

    def check_step(i, target):
        if i > 0:
            assert np.array_equal(sha(N(target[0])), g["step_latent_sha"][i - 1]), i - 1
            assert np.array_equal(sha(N(target[1])), g["step_denoised_sha"][i - 1]), i - 1

    cond = {"c_crossattn": [G(g["text_emb"], dev)], "fps": 16}
    uncond = {"c_crossattn": [G(g["uncond_emb"], dev)], "fps": 16}
    group = 1 if order == "sequential" else (lambda i: 1 if i in dock else None)
    ring_denoise_loop(st, [lat, den], model=replay_model, schedule=sched,
                      schedule_fn=lambda i: shift_window_schedule(i, **kw), width=w, height=h, frames=frames, cond=cond,
                      uncond=uncond, guidance_scale=float(g["guidance"]), group_size=group, cfg_batch=False,
                      on_input=on_input, before_step=check_step)
    assert cur["calls"] == 2 * base[-1]
    check_step(steps, [lat, den])
    assert np.array_equal(N(den), g["final_denoised"])


def test_ring_loop_dropin_class_replays_the_real_reference_pipeline(dev):
    """The reference's loop body (pipeline/t2v_normal_pipeline.py:526-552) call for call through the DROP-IN
    ``RingLatent``: get_window_latent, CFG + ddim_step as plain torch ops on the recorded UNet outputs, two
    set_window_latent calls -- same digests after every step."""
    from dynamicscaler_b200.shift_window_utils import RingLatent
    g, (frames, h, w, nw, nh, nf, steps), sched, kw, dock, base = load_ring_loop()
    F, H, W = frames * nf, h * nh, w * nw
    lat0 = G(randn(51, 1, 4, F, H, W), dev)
    h_lat = RingLatent(init_latent=lat0)
    h_den = RingLatent(init_latent=torch.zeros_like(lat0))
    gscale = float(g["guidance"])
    for i in range(steps):
        idx = steps - i - 1
        f32 = lambda v: torch.full((1, 1, 1, 1, 1), float(v), device=dev)                   # scheduler.py:78-81
        a_t, a_prev, sig, s1m = (f32(sched[k][idx]) for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                                                              "ddim_sqrt_one_minus_alphas"))
        for n, (left, top, fb) in enumerate(po.shift_window_jobs(i, **kw)):
            box = dict(pos_left=left, pos_right=left + w, pos_top=top, pos_down=top + h, frame_begin=fb, frame_end=fb + frames)
            x = h_lat.get_window_latent(**box)
            j = base[i] + n
            assert np.array_equal(sha(N(x)), g["unet_in_sha"][j]), (i, n)
            ec, eu = G(g["eps_cond"][j], dev), G(g["eps_uncond"][j], dev)
            e = eu + gscale * (ec - eu)                                                   # :603
            pred_x0 = (x - s1m * e) / a_t.sqrt()                                          # scheduler.py:84-89
            x_prev = a_prev.sqrt() * pred_x0 + (1. - a_prev - sig ** 2).sqrt() * e
            h_lat.set_window_latent(x_prev, **box)
            h_den.set_window_latent(pred_x0, **box)
        assert np.array_equal(sha(N(h_lat.torch_latent)), g["step_latent_sha"][i]), i
        assert np.array_equal(sha(N(h_den.torch_latent)), g["step_denoised_sha"][i]), i
    assert np.array_equal(N(h_den.torch_latent), g["final_denoised"])


CASES = [  # (kind, dtype, eta case of ddim.npz, frame ring, mask)
    ("ring_vec", torch.float32, 0, False, "plane"),
    ("ring_scalar", torch.float32, 3, False, "full"),
    ("ring_vec", torch.float32, 2, True, "full"),
    ("ring_scalar", torch.bfloat16, 0, True, None),
    ("ring_vec", torch.float16, 3, False, "plane"),
    ("sphere", torch.float32, 0, False, "plane"),
    ("sphere", torch.float32, 3, True, "full"),
]


@pytest.mark.parametrize("kind,dtype,case,fring,mask_kind", CASES)
def test_fused_reference_order_merge_vs_reference_chain(dev, kind, dtype, case, fring, mask_kind):
    """merge_denoised_ref (CFG + ddim_step inside ds_merge_nearest_ddim) == ddim_step on the host values (oracle,
    pinned to the reference scheduler by tests/golden/ddim.npz) followed by merge_all, bit for bit.  16-bit
    dtypes: the fused kernel keeps the chain in float32 and rounds ONCE (like the pixel-major fused merges), so
    the comparison is against the float32 oracle chain of the 16-bit inputs, rounded once."""
    from dynamicscaler_b200.step import PanoramaStepper, ddim_coefficients
    g = np.load(os.path.join(GOLDEN, "ddim.npz"))
    eta, index, gs, a_t, a_prev, sig, s1m = g["cases"][case]
    H, W, F, h, w = 32, 64, (12 if fring else 6), 12, 24
    st = PanoramaStepper(H, W, F, device=dev)
    fkw = dict(frame_begins=[9, 2], frames=6) if fring else {}
    if kind == "sphere":
        plan = st.plan_sphere([(0, 0), (30, 50), (-60, 200), (90, 0), (0, 20)], w, h, fov=100, mode="nearest", **fkw)
    else:
        off = 0 if kind == "ring_vec" else 3
        wins = [(off + 8 * a, 5 * b + off) for a in range(6) for b in range(4)]           # overlapping, wrapping
        plan = st.plan_ring(wins, w, h, **fkw)
        assert plan.vec == (4 if kind == "ring_vec" else 1)
    lat0 = G(randn(7, 1, 4, F, H, W), dev).to(dtype)
    mk = lambda: None if mask_kind is None else torch.zeros((1, H, W) if mask_kind == "plane" else (1, 4, F, H, W),
                                                            device=dev, dtype=dtype)
    if mask_kind == "plane" and fring:
        pytest.skip("a [1, H, W] mask has no frame axis")
    a_lat, a_den, a_m = lat0.clone(), torch.zeros_like(lat0), mk()
    b_lat, b_den, b_m = lat0.clone(), torch.zeros_like(lat0), mk()
    x = st.gather_all(lat0, plan)
    gen = torch.Generator(device=dev).manual_seed(11 + case)
    ec = torch.randn(x.shape, device=dev, generator=gen).to(dtype)
    eu = torch.randn(x.shape, device=dev, generator=gen).to(dtype) if gs != 1.0 else None
    nz = torch.randn(x.shape, device=dev, generator=gen).to(dtype) if sig > 0 else None
    f = lambda t: None if t is None else N(t.float())
    xp, x0 = po.ddim_step_cfg(f(x), f(ec), f(eu), gs, a_t, a_prev, sig, s1m, noise=f(nz))
    st.merge_all([a_lat, a_den], [G(xp, dev).to(dtype), G(x0, dev).to(dtype)], plan, mask=a_m)
    st.merge_denoised_ref([b_lat, b_den], x, ec, eu, ddim_coefficients(a_t, a_prev, sig, s1m, gs), plan, noise=nz, mask=b_m)
    if mask_kind is not None:
        assert torch.equal(a_m, b_m) and a_m.sum().item() > 0
    for ta, tb in ((a_lat, b_lat), (a_den, b_den)):
        assert torch.equal(ta, tb), (kind, case, float((ta.float() - tb.float()).abs().max()))
    assert not torch.equal(b_lat, lat0)
