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


CASES = [  # (kind, dtype, eta case of ddim.npz, frame ring, mask, merge_prev mix)
    ("ring_vec", torch.float32, 0, False, "plane", False),
    ("ring_scalar", torch.float32, 3, False, "full", False),
    ("ring_vec", torch.float32, 2, True, "full", False),
    ("ring_scalar", torch.bfloat16, 0, True, None, False),
    ("ring_vec", torch.float16, 3, False, "plane", False),
    ("sphere", torch.float32, 0, False, "plane", False),
    ("sphere", torch.float32, 3, True, "full", False),
    ("ring_vec", torch.float32, 0, True, "full", True),
    ("ring_scalar", torch.float32, 3, False, "full", True),
    ("ring_vec", torch.bfloat16, 0, False, "full", True),
    ("ring_scalar", torch.float16, 2, True, "full", True),
]


@pytest.mark.parametrize("kind,dtype,case,fring,mask_kind,prev_mix", CASES)
def test_fused_reference_order_merge_vs_reference_chain(dev, kind, dtype, case, fring, mask_kind, prev_mix):
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
    xp_t, pkw = G(xp, dev).to(dtype), {}
    if prev_mix:     # i2v :938-943 on the ROUNDED x_prev, torch ops in the tensor dtype (the drop-in kernel is pinned by its own test)
        mv = (torch.rand(x.shape, device=dev, generator=gen) > 0.4).to(dtype)
        mixed = xp_t * (1 - 0.35) + x * 0.35
        xp_t = xp_t * (1 - mv) + mixed * mv
        pkw = dict(prev=x, prev_mask=mv, prev_ratio=0.35)
    st.merge_all([a_lat, a_den], [xp_t, G(x0, dev).to(dtype)], plan, mask=a_m)
    st.merge_denoised_ref([b_lat, b_den], x, ec, eu, ddim_coefficients(a_t, a_prev, sig, s1m, gs), plan, noise=nz, mask=b_m,
                          **pkw)
    if mask_kind is not None:
        assert torch.equal(a_m, b_m) and a_m.sum().item() > 0
    for ta, tb in ((a_lat, b_lat), (a_den, b_den)):
        assert torch.equal(ta, tb), (kind, case, float((ta.float() - tb.float()).abs().max()))
    assert not torch.equal(b_lat, lat0)


# ---------------------------------------------------------------- the planar I2V loop ----
def load_ring_loop_i2v():
    g = np.load(os.path.join(GOLDEN, "ring_loop_i2v.npz"))
    names = ("frames", "height", "width", "total_h", "total_w", "total_f", "nw", "nh", "loop_step", "steps",
             "loop_step_frame", "begin_index_offset")
    c = {k: int(v) for k, v in zip(names, g["cfg"])}
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    kw = dict(width=c["width"], height=c["height"], frames=c["frames"], total_w=c["total_w"], total_h=c["total_h"],
              total_f=c["total_f"], num_windows_w=c["nw"], num_windows_h=c["nh"], loop_step=c["loop_step"],
              overlap_ratio_f=float(g["overlap_f"]), loop_step_frame=c["loop_step_frame"], dock_at_f=True,
              begin_index_offset=c["begin_index_offset"])
    counts = [len(po.i2v_planar_jobs(i, **kw)) for i in range(c["steps"])]
    base = [0] + [int(v) for v in np.cumsum(counts)]
    image = ((g["image_u8"].astype(np.float32).transpose(2, 0, 1) / np.float32(255.)) - np.float32(0.5)) * np.float32(2)
    assert np.array_equal(sha(image), g["image_f32_sha"])
    return g, c, sched, kw, base, image


def test_i2v_ring_loop_replays_the_real_reference_pipeline(dev):
    """3 DDIM steps of the reference's planar I2V loop (27 + 36 + 27 overlapping windows, frame docking, latent-shaped
    mask panorama, re-noise mix, merge_prev mix, one condition-image window per latent window): every UNet input, every
    condition-image window (sha256) and the latent / denoised panoramas after every step equal the real pipeline's,
    bit for bit."""
    from dynamicscaler_b200.loop import i2v_planar_schedule, ring_denoise_loop
    from dynamicscaler_b200.step import PanoramaStepper
    g, c, sched, kw, base, image = load_ring_loop_i2v()
    frames, h, w = c["frames"], c["height"] // 8, c["width"] // 8
    F, H, W = c["total_f"], c["total_h"] // 8, c["total_w"] // 8
    st = PanoramaStepper(H, W, F, device=dev)
    lat = G(randn(81, 1, 4, F, H, W), dev)
    den, mask = torch.zeros_like(lat), torch.ones_like(lat)           # the loop must zero the mask itself
    d_image = G(image, dev)
    text, uncond_emb = G(g["text_emb"], dev), G(g["uncond_emb"], dev)
    cur = {"calls": 0, "draws": 0}

    def cond_fn(i, job_ids, boxes):
        wins = st.gather_image_windows(d_image, [(l * 8, t * 8) for (l, t, _) in boxes], c["width"], c["height"])
        for q, job in enumerate(job_ids):
            assert np.array_equal(sha(N(wins[q])), g["img_win_sha"][base[i] + job]), (i, job)
        emb = torch.cat([G(g["img_emb"][base[i] + job], dev) for job in job_ids], dim=0)
        return {"c_crossattn": [torch.cat([text.expand(len(job_ids), -1, -1), emb], dim=1)], "fps": 16}

    def on_input(i, job_ids, xb):
        cur.update(i=i, jobs=list(job_ids), branch=0)
        for q, job in enumerate(job_ids):
            assert np.array_equal(sha(N(xb[q:q + 1])), g["unet_in_sha"][base[i] + job]), (i, job)

    def replay_model(x, ts, c_crossattn=None, fps=None, **kw_):
        src = g["eps_cond"] if cur["branch"] == 0 else g["eps_uncond"]
        assert c_crossattn[0].shape[1] == (77 + 4 if cur["branch"] == 0 else 77)
        cur["branch"] += 1
        cur["calls"] += x.shape[0]
        return torch.cat([G(src[base[cur["i"]] + job], dev) for job in cur["jobs"]], dim=0)

    def eps_fn(i, job_ids, shape, device):
        cur["draws"] += 1
        return G(randn(6000 + cur["draws"], *shape[1:]), device).view(shape)

    def check_step(i, target):
        if i > 0:
            assert np.array_equal(sha(N(target[0])), g["step_latent_sha"][i - 1]), i - 1
            assert np.array_equal(sha(N(target[1])), g["step_denoised_sha"][i - 1]), i - 1

    ring_denoise_loop(st, [lat, den], model=replay_model, schedule=sched,
                      schedule_fn=lambda i: i2v_planar_schedule(i, **kw), width=w, height=h, frames=frames, cond=cond_fn,
                      uncond={"c_crossattn": [uncond_emb], "fps": 16}, guidance_scale=float(g["guidance"]), group_size=1,
                      cfg_batch=False, on_input=on_input, before_step=check_step, mask=mask,
                      renoise_ratio=float(g["ratio"]), prev_ratio_list=[float(v) for v in g["prev_ratios"]], eps_fn=eps_fn)
    assert cur["calls"] == 2 * base[-1] and cur["draws"] == int(g["renoise_draws"])
    check_step(c["steps"], [lat, den])
    assert np.array_equal(N(den), g["final_denoised"])


def test_i2v_ring_loop_dropin_classes_replay_the_real_reference_pipeline(dev):
    """The reference's loop body (pipeline/i2v_sphere_panorama_pipeline.py:859-968) call for call through the DROP-IN
    ``RingLatent`` / ``RingImageTensor`` / ``mix_latents_with_mask``: same digests after every step."""
    from dynamicscaler_b200.shift_window_utils import RingImageTensor, RingLatent
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    g, c, sched, kw, base, image = load_ring_loop_i2v()
    frames, h, w = c["frames"], c["height"] // 8, c["width"] // 8
    lat0 = G(randn(81, 1, 4, c["total_f"], c["total_h"] // 8, c["total_w"] // 8), dev)
    h_lat = RingLatent(init_latent=lat0)
    h_den = RingLatent(init_latent=torch.zeros_like(lat0))
    h_img = RingImageTensor(image_path=None, image_tensor=G(image, dev), height=c["total_h"], width=c["total_w"])
    ac = torch.from_numpy(np.asarray(sched["alphas_cumprod"], dtype=np.float32)).to(dev)
    ddim_ts = [int(v) for v in sched["ddim_timesteps"]]
    total, ratio, gscale = c["steps"], float(g["ratio"]), float(g["guidance"])
    draws = 0
    for i in range(total):
        idx = total - i - 1
        f32 = lambda v: torch.full((1, 1, 1, 1, 1), float(v), device=dev)                   # scheduler.py:78-81
        a_t, a_prev, sig, s1m = (f32(sched[k][idx]) for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                                                              "ddim_sqrt_one_minus_alphas"))
        h_msk = RingLatent(init_latent=torch.zeros_like(lat0))                            # :810
        for n, (left, top, fb) in enumerate(po.i2v_planar_jobs(i, **kw)):
            box = dict(pos_left=left, pos_right=left + w, pos_top=top, pos_down=top + h, frame_begin=fb, frame_end=fb + frames)
            x = h_lat.get_window_latent(**box)
            prev = x.clone()
            m = h_msk.get_window_latent(**box)
            if i < total - 1:                                                             # :875-884
                a_a, a_b = ac[ddim_ts[total - i - 2]], ac[ddim_ts[total - i - 1]]
                cc, ss = torch.sqrt(a_b / a_a), torch.sqrt(1 - a_b / a_a)
                draws += 1
                noised = cc * x.clone() + ss * G(randn(6000 + draws, *x.shape), dev)
                x = mix_latents_with_mask(latent_1=x, latent_to_add=noised, mask=m, mix_ratio=ratio)
            j = base[i] + n
            assert np.array_equal(sha(N(x)), g["unet_in_sha"][j]), (i, n)
            img = h_img.get_window_tensor(pos_left=left * 8, pos_right=left * 8 + c["width"], pos_top=top * 8,
                                          pos_down=top * 8 + c["height"])
            assert np.array_equal(sha(N(img)), g["img_win_sha"][j]), (i, n)
            ec, eu = G(g["eps_cond"][j], dev), G(g["eps_uncond"][j], dev)
            e = eu + gscale * (ec - eu)
            pred_x0 = (x - s1m * e) / a_t.sqrt()
            x_prev = a_prev.sqrt() * pred_x0 + (1. - a_prev - sig ** 2).sqrt() * e
            if i < total - 1:                                                             # :938-943
                x_prev = mix_latents_with_mask(latent_1=x_prev, latent_to_add=prev, mask=m, mix_ratio=float(g["prev_ratios"][i]))
            h_lat.set_window_latent(x_prev, **box)
            h_den.set_window_latent(pred_x0, **box)
            h_msk.set_window_latent(torch.ones_like(x_prev), **box)
        assert np.array_equal(sha(N(h_lat.torch_latent)), g["step_latent_sha"][i]), i
        assert np.array_equal(sha(N(h_den.torch_latent)), g["step_denoised_sha"][i]), i
    assert draws == int(g["renoise_draws"])
    assert np.array_equal(N(h_den.torch_latent), g["final_denoised"])


@pytest.mark.parametrize("group_size", [None, 4])
def test_i2v_ring_loop_batched_order_vs_oracle(dev, group_size):
    """Whole-step (and 4-window) groups on OVERLAPPING windows: one stand-in UNet call per group; its outputs are
    recorded and replayed into the oracle composition with the same grouping, panoramas must agree bit for bit."""
    from dynamicscaler_b200.loop import i2v_planar_schedule, ring_denoise_loop
    from dynamicscaler_b200.step import PanoramaStepper
    g, c, sched, kw, base, image = load_ring_loop_i2v()
    frames, h, w = c["frames"], c["height"] // 8, c["width"] // 8
    F, H, W = c["total_f"], c["total_h"] // 8, c["total_w"] // 8
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = randn(83, 1, 4, F, H, W)
    lat, den = G(lat0, dev), torch.zeros(1, 4, F, H, W, device=dev)
    mask = torch.zeros_like(lat)
    ctx_c, ctx_u = G(randn(84, 1, 77, 16), dev), G(randn(85, 1, 77, 16), dev)
    rec, cur = {}, {}

    def on_input(i, job_ids, xb):
        cur.update(i=i, jobs=list(job_ids))

    def model(x, ts, c_crossattn=None, fps=None, **kw_):              # elementwise stand-in, both CFG branches in one batch
        n = len(cur["jobs"])
        assert x.shape[0] == 2 * n and ts.numel() == 2 * n * frames
        y = torch.sin(x * 1.3) * 0.5 + c_crossattn[0][:, 3, 5].reshape(-1, 1, 1, 1, 1) * 0.1
        for q, job in enumerate(cur["jobs"]):
            rec[(cur["i"], job)] = (N(x[q:q + 1]), N(y[q:q + 1]), N(y[n + q:n + q + 1]))
        return y

    def eps_fn(i, job_ids, shape, device):
        return torch.stack([G(randn(9500 + 100 * i + j, *shape[1:]), device) for j in job_ids])

    prev = [0.3, 0.5, 0.7]
    ring_denoise_loop(st, [lat, den], model=model, schedule=sched, schedule_fn=lambda i: i2v_planar_schedule(i, **kw),
                      width=w, height=h, frames=frames, cond={"c_crossattn": [ctx_c], "fps": 16},
                      uncond={"c_crossattn": [ctx_u], "fps": 16}, guidance_scale=7.5, group_size=group_size,
                      on_input=on_input, mask=mask, renoise_ratio=0.6, prev_ratio_list=prev, eps_fn=eps_fn)
    nw = c["nw"] * c["nh"]

    def groups_fn(i, n_jobs):
        if group_size is None:
            return [list(range(n_jobs))]
        return [list(range(s0 + g0, min(s0 + g0 + group_size, s0 + nw))) for s0 in range(0, n_jobs, nw)
                for g0 in range(0, nw, group_size)]

    def model_fn(i, n, x):
        xin, ec, eu = rec[(i, n)]
        assert np.array_equal(x, xin), (i, n)                           # ring gather + re-noise mix, bit for bit
        return ec, eu

    ref_lat, ref_den, _ = po.ring_denoise_loop(lat0, sched, lambda i: po.i2v_planar_jobs(i, **kw), model_fn, 7.5, w, h,
                                               frames, groups_fn=groups_fn, renoise_ratio=0.6, prev_ratio_list=prev,
                                               eps_fn=lambda i, n, shape: randn(9500 + 100 * i + n, *shape), use_mask=True)
    assert np.array_equal(N(lat), ref_lat) and np.array_equal(N(den), ref_den)
    assert mask.sum().item() > 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_renoise_mix_vs_reference_chain(dev, dtype):
    """ds_renoise_mix == re_noise (pipeline/scheduler.py:98-110, epsilon injected, 0-dim float32 c and s) followed by
    mix_latents_with_mask (utils/tensor_utils.py:30-39) as torch CPU ops in the tensor dtype, bit for bit; vector and
    scalar paths; window-shaped mask (I2V planar loop) and the [0, 0, [0]] plane of every mask window (T2V planar loop)."""
    from dynamicscaler_b200 import _lib, lib
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    for shape in ((3, 1, 4, 4, 8, 12), (2, 1, 4, 3, 5, 7)):
        x, eps = (torch.from_numpy(randn(s, *shape)).to(dtype) for s in (91, 92))
        m = (torch.from_numpy(randn(93, *shape)) > 0.3).to(dtype)
        c_t, s_t, r = torch.tensor(0.9721), torch.tensor(0.2346), 0.6
        noised = c_t * x + s_t * eps
        mixed = x * (1 - r) + noised * r
        dx, de, dm = x.to(dev), eps.to(dev), m.to(dev)
        for plane in (False, True):
            if plane:                       # per window n: mask[n][0, 0, [0]] -> [1, h, w], repeated over B, C, F (tensor_utils.py:21-23)
                me = torch.stack([m[n][0, 0, [0]].unsqueeze(0).unsqueeze(0).repeat(*shape[1:4], 1, 1) for n in range(shape[0])])
            else:
                me = m
            want = x * (1 - me) + mixed * me
            assert want.dtype == dtype
            out = torch.empty_like(dx)
            hw = shape[-1] * shape[-2]
            lib.check(lib.ds_renoise_mix(dx.data_ptr(), de.data_ptr(), dm.data_ptr(), out.data_ptr(), float(c_t), float(s_t),
                                         r, 1 - r, dx.numel(), dx.numel() // shape[0], hw if plane else 0,
                                         _lib.dtype_code(dtype), _lib.stream_ptr(dev)), "ds_renoise_mix")
            assert torch.equal(out.cpu(), want), (dtype, shape, plane, float((out.cpu().float() - want.float()).abs().max()))


# ------------------------------------------ the planar loop of the T2V sphere pipeline ----
def load_ring_loop_t2v():
    g = np.load(os.path.join(GOLDEN, "ring_loop_t2v.npz"))
    names = ("frames", "height", "width", "total_h", "total_w", "nw", "nh", "nf", "loop_step", "steps")
    c = {k: int(v) for k, v in zip(names, g["cfg"])}
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    kw = dict(width=c["width"], height=c["height"], frames=c["frames"], total_w=c["total_w"], total_h=c["total_h"],
              num_windows_w=c["nw"], num_windows_h=c["nh"], num_windows_f=c["nf"], loop_step=c["loop_step"], dock_at_h=True)
    counts = [len(po.t2v_planar_jobs(i, **kw)) for i in range(c["steps"])]
    return g, c, sched, kw, [0] + [int(v) for v in np.cumsum(counts)]


def test_t2v_ring_loop_replays_the_real_reference_pipeline(dev):
    """3 DDIM steps of the T2V sphere pipeline's planar loop (18 + 24 + 24 overlapping windows, windows docked at the
    top / bottom edge, latent-shaped mask panorama whose [0, 0, [0]] plane drives the re-noise mix): every UNet input
    and the latent / denoised panoramas after every step equal the real pipeline's, bit for bit."""
    from dynamicscaler_b200.loop import ring_denoise_loop, t2v_planar_schedule
    from dynamicscaler_b200.step import PanoramaStepper
    g, c, sched, kw, base = load_ring_loop_t2v()
    frames, h, w = c["frames"], c["height"] // 8, c["width"] // 8
    F, H, W = frames * c["nf"], c["total_h"] // 8, c["total_w"] // 8
    st = PanoramaStepper(H, W, F, device=dev)
    lat = G(randn(95, 1, 4, F, H, W), dev)
    den, mask = torch.zeros_like(lat), torch.ones_like(lat)
    cur = {"calls": 0, "draws": 0}

    def on_input(i, job_ids, xb):
        cur.update(i=i, jobs=list(job_ids), branch=0)
        for q, job in enumerate(job_ids):
            assert np.array_equal(sha(N(xb[q:q + 1])), g["unet_in_sha"][base[i] + job]), (i, job)

    def replay_model(x, ts, c_crossattn=None, fps=None, **kw_):
        src = g["eps_cond"] if cur["branch"] == 0 else g["eps_uncond"]
        cur["branch"] += 1
        cur["calls"] += x.shape[0]
        return torch.cat([G(src[base[cur["i"]] + job], dev) for job in cur["jobs"]], dim=0)

    def eps_fn(i, job_ids, shape, device):
        cur["draws"] += 1
        return G(randn(7000 + cur["draws"], *shape[1:]), device).view(shape)

    def check_step(i, target):
        if i > 0:
            assert np.array_equal(sha(N(target[0])), g["step_latent_sha"][i - 1]), i - 1
            assert np.array_equal(sha(N(target[1])), g["step_denoised_sha"][i - 1]), i - 1

    ring_denoise_loop(st, [lat, den], model=replay_model, schedule=sched,
                      schedule_fn=lambda i: t2v_planar_schedule(i, **kw), width=w, height=h, frames=frames,
                      cond={"c_crossattn": [G(g["text_emb"], dev)], "fps": 16},
                      uncond={"c_crossattn": [G(g["uncond_emb"], dev)], "fps": 16}, guidance_scale=float(g["guidance"]),
                      group_size=1, cfg_batch=False, on_input=on_input, before_step=check_step, mask=mask,
                      renoise_ratio=float(g["ratio"]), eps_fn=eps_fn, mask_frame0=True)
    assert cur["calls"] == 2 * base[-1] and cur["draws"] == int(g["renoise_draws"])
    check_step(c["steps"], [lat, den])
    assert np.array_equal(N(den), g["final_denoised"])


def test_t2v_ring_loop_dropin_classes_replay_the_real_reference_pipeline(dev):
    """The same run call for call through the DROP-IN ``RingLatent`` / ``mix_latents_with_mask`` (the loop body of
    pipeline/t2v_sphere_panorama_pipeline.py:537-632, incl. the [1, h, w] mask plane of :553)."""
    from dynamicscaler_b200.shift_window_utils import RingLatent
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    g, c, sched, kw, base = load_ring_loop_t2v()
    frames, h, w = c["frames"], c["height"] // 8, c["width"] // 8
    lat0 = G(randn(95, 1, 4, frames * c["nf"], c["total_h"] // 8, c["total_w"] // 8), dev)
    h_lat = RingLatent(init_latent=lat0)
    h_den = RingLatent(init_latent=torch.zeros_like(lat0))
    ac = torch.from_numpy(np.asarray(sched["alphas_cumprod"], dtype=np.float32)).to(dev)
    ddim_ts = [int(v) for v in sched["ddim_timesteps"]]
    total, ratio, gscale = c["steps"], float(g["ratio"]), float(g["guidance"])
    draws = 0
    for i in range(total):
        idx = total - i - 1
        f32 = lambda v: torch.full((1, 1, 1, 1, 1), float(v), device=dev)
        a_t, a_prev, sig, s1m = (f32(sched[k][idx]) for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                                                              "ddim_sqrt_one_minus_alphas"))
        h_msk = RingLatent(init_latent=torch.zeros_like(lat0))                            # :494
        for n, (left, top, fb) in enumerate(po.t2v_planar_jobs(i, **kw)):
            box = dict(pos_left=left, pos_right=left + w, pos_top=top, pos_down=top + h, frame_begin=fb, frame_end=fb + frames)
            x = h_lat.get_window_latent(**box)
            m = h_msk.get_window_latent(**box)
            if i < total - 1:                                                             # :548-559
                a_a, a_b = ac[ddim_ts[total - i - 2]], ac[ddim_ts[total - i - 1]]
                cc, ss = torch.sqrt(a_b / a_a), torch.sqrt(1 - a_b / a_a)
                draws += 1
                noised = cc * x.clone() + ss * G(randn(7000 + draws, *x.shape), dev)
                x = mix_latents_with_mask(latent_1=x, latent_to_add=noised, mask=m[0, 0, [0]], mix_ratio=ratio)
            j = base[i] + n
            assert np.array_equal(sha(N(x)), g["unet_in_sha"][j]), (i, n)
            ec, eu = G(g["eps_cond"][j], dev), G(g["eps_uncond"][j], dev)
            e = eu + gscale * (ec - eu)
            pred_x0 = (x - s1m * e) / a_t.sqrt()
            x_prev = a_prev.sqrt() * pred_x0 + (1. - a_prev - sig ** 2).sqrt() * e
            h_lat.set_window_latent(x_prev, **box)
            h_den.set_window_latent(pred_x0, **box)
            h_msk.set_window_latent(torch.ones_like(x_prev), **box)
        assert np.array_equal(sha(N(h_lat.torch_latent)), g["step_latent_sha"][i]), i
        assert np.array_equal(sha(N(h_den.torch_latent)), g["step_denoised_sha"][i]), i
    assert draws == int(g["renoise_draws"])
    assert np.array_equal(N(h_den.torch_latent), g["final_denoised"])
