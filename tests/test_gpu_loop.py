"""Denoising-loop driver (dynamicscaler_b200/loop.py) and the fused gather / merge prologues and
epilogues of SURVEY §8f-1 on a B200 (``pytest -m gpu``):

* replay of the REAL reference pipeline run (tests/golden/loop.npz, made by
  tests/golden/make_golden_loop.py from ``basic_sample_shift_shpere_panorama``): with
  ``group_size=1`` the driver must feed the UNet bit-identical windows at every call and end on
  bit-identical panoramas;
* the batched order (whole step = one group, one UNet call) against the oracle composition, with a
  tiny stand-in module that has the reference model's call signature (tests/standin_unet.py);
* the merge_prev_denoised_ratio prologue and the latent-shaped (I2V) mask against the oracle.
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


def load_loop():
    g = np.load(os.path.join(GOLDEN, "loop.npz"))
    F, H, W, h, w, steps, fov = [int(v) for v in g["cfg"]]
    views0 = [(float(p), float(t)) for p, t in g["views"]]
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    views_fn = lambda i: [(p, t + (i % 10) * (fov // 10)) for p, t in views0]
    return g, (F, H, W, h, w, steps, fov), views0, sched, views_fn


def test_loop_replays_the_real_reference_pipeline(dev):
    """3 DDIM steps x 6 views of the reference's own sequential loop, re-noise mix on: every UNet input
    (sha256) and the final latent / denoised panoramas equal the real pipeline's, bit for bit."""
    from dynamicscaler_b200.loop import sphere_denoise_loop
    from dynamicscaler_b200.step import PanoramaStepper
    g, (F, H, W, h, w, steps, fov), views0, sched, views_fn = load_loop()
    nv = len(views0)
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = G(randn(41, 1, 4, F, H, W), dev)
    state = st.make_state([lat0, torch.zeros_like(lat0)], mask=torch.zeros(1, H, W, device=dev))
    calls = {"n": 0, "draws": 0}

    def replay_model(x, ts, c_crossattn=None, fps=None, **kw):
        k = calls["n"]
        calls["n"] += 1
        j, branch = divmod(k, 2)                                   # cond call, then uncond call, per window
        assert ts.numel() == 1 and kw["temporal_length"] == F and kw["clean_cond"] is True
        assert np.array_equal(sha(N(x)), g["unet_in_sha"][j]), f"UNet input of call {k} differs from the reference's"
        want_ctx = g["text_emb"] if branch == 0 else g["uncond_emb"]
        assert np.array_equal(N(c_crossattn[0]), want_ctx)
        return G(g["eps_cond"][j] if branch == 0 else g["eps_uncond"][j], dev)

    def eps_fn(i, job_ids, shape, device):
        calls["draws"] += 1
        return G(randn(5000 + calls["draws"], *shape[1:]), device).view(shape)

    cond = {"c_crossattn": [G(g["text_emb"], dev)], "fps": 16}
    uncond = {"c_crossattn": [G(g["uncond_emb"], dev)], "fps": 16}
    sphere_denoise_loop(st, state, model=replay_model, schedule=sched, views_fn=views_fn, width=w, height=h, frames=F,
                        cond=cond, uncond=uncond, guidance_scale=float(g["guidance"]), fov=fov, group_size=1,
                        renoise_ratio=float(g["ratio"]), eps_fn=eps_fn, cfg_batch=False)
    assert calls["n"] == steps * nv * 2 and calls["draws"] == int(g["renoise_draws"])
    out_lat, out_den = state.export()
    assert np.array_equal(N(out_lat), g["final_latent"])
    assert np.array_equal(N(out_den), g["final_denoised"])


@pytest.mark.parametrize("group_size", [None, 2])
def test_loop_batched_order_vs_oracle(dev, group_size):
    """Whole-step (and 2-window) groups: ONE UNet call per group with timesteps of length n*f (P9),
    both CFG branches in one batch.  The stand-in's outputs are recorded and replayed into the oracle
    composition, so panoramas must agree bit for bit; the batched call itself must agree with
    per-window calls to 1e-4 (fp32 convolutions, TF32 off)."""
    from dynamicscaler_b200.loop import sphere_denoise_loop
    from dynamicscaler_b200.step import PanoramaStepper
    from standin_unet import StandInUNet
    g, (F, H, W, h, w, steps, fov), views0, sched, views_fn = load_loop()
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = randn(43, 1, 4, F, H, W)
    state = st.make_state([G(lat0, dev), torch.zeros(1, 4, F, H, W, device=dev)], mask=torch.zeros(1, H, W, device=dev))
    net = StandInUNet(ctx_dim=g["text_emb"].shape[-1]).to(dev).eval()
    torch.backends.cudnn.allow_tf32 = False                          # the P9 check below compares fp32 convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    rec = {}                                                         # (step, job) -> (x, eps_cond, eps_uncond)
    cur = {}

    def on_input(i, job_ids, xb):
        cur["i"], cur["jobs"] = i, list(job_ids)

    def model(x, ts, c_crossattn=None, fps=None, **kw):
        n = x.shape[0] // 2
        assert ts.numel() == (x.shape[0] * F if x.shape[0] > 1 else 1)
        with torch.no_grad():
            y = net(x, ts, c_crossattn=c_crossattn, fps=fps, **kw)
            # P9 on the stand-in: the batched call equals per-window calls
            y1 = net(x[:1], ts[:1], c_crossattn=[c_crossattn[0][:1]], fps=fps, **kw)
        assert (y[:1] - y1).abs().max().item() <= 1e-4
        for q, job in enumerate(cur["jobs"]):
            rec[(cur["i"], job)] = (N(x[q:q + 1]), N(y[q:q + 1]), N(y[n + q:n + q + 1]))
        return y

    def eps_fn(i, job_ids, shape, device):
        return torch.stack([G(randn(9000 + 100 * i + j, *shape[1:]), device) for j in job_ids])

    cond = {"c_crossattn": [G(g["text_emb"], dev)], "fps": 16}
    uncond = {"c_crossattn": [G(g["uncond_emb"], dev)], "fps": 16}
    sphere_denoise_loop(st, state, model=model, schedule=sched, views_fn=views_fn, width=w, height=h, frames=F, cond=cond,
                        uncond=uncond, guidance_scale=7.5, fov=fov, group_size=group_size, renoise_ratio=0.6,
                        prev_ratio_list=[0.3, 0.5, 0.7], eps_fn=eps_fn, on_input=on_input)

    def model_fn(i, n, x):
        xin, ec, eu = rec[(i, n)]
        assert np.array_equal(x, xin), (i, n)                       # gather + re-noise mix, bit for bit
        return ec, eu

    ref_lat, ref_den, ref_mask, _ = po.sphere_denoise_loop(
        lat0, sched, views_fn, model_fn, lambda i, n, shape: randn(9000 + 100 * i + n, *shape), 7.5, fov=fov, width=w,
        height=h, renoise_ratio=0.6, prev_ratio_list=[0.3, 0.5, 0.7], group_size=group_size)
    out_lat, out_den = state.export()
    assert np.array_equal(N(out_lat), ref_lat) and np.array_equal(N(out_den), ref_den)
    assert np.array_equal(N(state.export_mask())[0], ref_mask[0, 0, 0])


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_merge_prev_mix_full_mask_vs_oracle(dev, mode):
    """The I2V flavour: latent-shaped mask panorama (i2v_sphere_panorama_pipeline.py:220), re-noise
    mix with a windows-shaped mask view and the merge_prev prologue inside both fused merges."""
    from dynamicscaler_b200.step import PanoramaStepper, ddim_coefficients, renoise_coefficients
    F, H, W, h, w = 16, 32, 64, 10, 16
    views = po.default_sphere_views(3)[:8]
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = randn(51, 1, 4, F, H, W)
    mask0 = (randn(52, 1, 4, F, H, W) > 0.3).astype(np.float32)      # some pixels already denoised this step
    state = st.make_state([G(lat0, dev), torch.zeros(1, 4, F, H, W, device=dev)], mask=G(mask0, dev))
    plan_n = st.plan_sphere(views, w, h, fov=120, mode="nearest")
    plan = st.plan_sphere(views, w, h, fov=120, mode=mode)
    n = len(views)
    eps = randn(53, n, 1, 4, F, h, w)
    ac = np.linspace(0.999, 0.01, 1000).astype(np.float32)
    ts = np.array([1, 334, 667])
    c, s = renoise_coefficients(ac, ts, 0, 1)
    x, raw, mview = st.gather_all(state, plan_n, renoise=dict(eps=G(eps, dev), c=c, s=s, ratio=0.6, want_raw=True,
                                                             want_mask=True))
    oc, os_ = po.renoise_scalars(ac, ts, 0, 1)
    assert (np.float32(c), np.float32(s)) == (oc, os_)
    for j, (ph, th) in enumerate(views):
        r_raw, _ = po.proxy_get_nearest(lat0, 120, th, ph, w, h)
        r_m, _ = po.proxy_get_nearest(mask0, 120, th, ph, w, h)
        r_x = po.mix_mask(r_raw, po.re_noise(r_raw, eps[j], oc, os_), r_m, 0.6)
        assert np.array_equal(N(raw[j]), r_raw) and np.array_equal(N(mview[j]), r_m) and np.array_equal(N(x[j]), r_x), j
    ec, eu, nz = randn(54, *eps.shape), randn(55, *eps.shape), randn(56, *eps.shape)
    coef = ddim_coefficients(0.5, 0.55, 0.1, 0.70710677, 7.5)
    st.merge_denoised(state, x, G(ec, dev), G(eu, dev), coef, plan, noise=G(nz, dev), prev=raw, prev_mask=mview,
                      prev_ratio=0.4)
    ref_lat, ref_den, ref_mask = lat0.copy(), np.zeros_like(lat0), mask0.copy()
    put = po.proxy_set_nearest if mode == "nearest" else po.proxy_set_bilinear
    for j, (ph, th) in enumerate(views):
        xp, x0 = po.ddim_step_cfg(N(x[j]), ec[j], eu[j], 7.5, 0.5, 0.55, 0.1, 0.70710677, noise=nz[j])
        xp = po.mix_mask(xp, N(raw[j]), N(mview[j]), 0.4)
        put(ref_lat, xp, 120, th, ph)
        put(ref_den, x0, 120, th, ph)
        put(ref_mask, np.ones_like(xp), 120, th, ph)
    out_lat, out_den = state.export()
    tol = 0.0 if mode == "nearest" else 1e-5
    assert np.abs(N(out_lat) - ref_lat).max() <= tol and np.abs(N(out_den) - ref_den).max() <= tol
    assert np.array_equal(N(state.export_mask()), ref_mask)


def test_dropin_classes_replay_the_real_reference_pipeline(dev):
    """The reference's own call sites, one view at a time, through the DROP-IN classes (regime (i)): the body of
    the real loop (pipeline/t2v_sphere_panorama_pipeline.py:166-291) mirrored call for call -- PanoramaLatentProxy /
    PanoramaTensor gets, re_noise with the recorded epsilon, mix_latents_with_mask, the recorded UNet outputs,
    CFG and ddim_step as plain torch ops, the three sets -- must land on the REAL pipeline's final panoramas
    (tests/golden/loop.npz) bit for bit, and feed the UNet bit-identical windows on the way."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy, PanoramaTensor
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    g, (F, H, W, h, w, steps, fov), views0, sched, views_fn = load_loop()
    nv = len(views0)
    lat0 = G(randn(41, 1, 4, F, H, W), dev)
    h_lat = PanoramaLatentProxy(equirect_tensor=lat0)
    h_den = PanoramaLatentProxy(equirect_tensor=torch.zeros_like(lat0))
    ac = torch.from_numpy(np.asarray(sched["alphas_cumprod"], dtype=np.float32)).to(dev)
    ddim_ts = [int(v) for v in sched["ddim_timesteps"]]
    total, ratio, gscale = len(ddim_ts), float(g["ratio"]), float(g["guidance"])
    draws = 0
    for i in range(steps):
        h_msk = PanoramaTensor(equirect_tensor=torch.zeros_like(lat0[0, 0, [0]]))          # :173
        for n, (ph, th) in enumerate(views_fn(i)):
            view, _ = h_lat.get_view_tensor_no_interpolate(fov=fov, theta=th, phi=ph, width=w, height=h)
            m, _ = h_msk.get_view_tensor_no_interpolate(fov=fov, theta=th, phi=ph, width=w, height=h)
            if i < total - 1:                                                             # :212-221
                a_a, a_b = ac[ddim_ts[total - i - 2]], ac[ddim_ts[total - i - 1]]         # scheduler.py:99-105
                c, s = torch.sqrt(a_b / a_a), torch.sqrt(1 - a_b / a_a)
                draws += 1
                noised = c * view.clone() + s * G(randn(5000 + draws, *view.shape), dev)
                view = mix_latents_with_mask(latent_1=view, latent_to_add=noised, mask=m, mix_ratio=ratio)
            j = i * nv + n
            assert np.array_equal(sha(N(view.contiguous())), g["unet_in_sha"][j]), (i, n)
            ec, eu = G(g["eps_cond"][j], dev), G(g["eps_uncond"][j], dev)
            e = eu + gscale * (ec - eu)                                                   # :258
            idx = total - i - 1
            f32 = lambda v: torch.full((1, 1, 1, 1, 1), float(v), device=dev)            # scheduler.py:78-81
            a_t, a_prev, sig, s1m = (f32(sched[k][idx]) for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                                                                  "ddim_sqrt_one_minus_alphas"))
            pred_x0 = (view - s1m * e) / a_t.sqrt()                                       # :84-89 (sigma = 0)
            x_prev = a_prev.sqrt() * pred_x0 + (1. - a_prev - sig ** 2).sqrt() * e
            h_lat.set_view_tensor_no_interpolation(view_tensor=x_prev, fov=fov, theta=th, phi=ph)
            h_den.set_view_tensor_no_interpolation(view_tensor=pred_x0, fov=fov, theta=th, phi=ph)
            h_msk.set_view_tensor_no_interpolation(view_tensor=torch.ones_like(x_prev[0, 0, [0]]), fov=fov, theta=th, phi=ph)
    assert draws == int(g["renoise_draws"])
    assert np.array_equal(N(h_lat.get_equirect_tensor()), g["final_latent"])
    assert np.array_equal(N(h_den.get_equirect_tensor()), g["final_denoised"])
