"""The I2V sphere loop on a B200 (``pytest -m gpu``): replay of the REAL
``VC2_Pipeline_I2V_SpherePano.basic_sample_shift_shpere_panorama`` run (tests/golden/loop_i2v.npz, made by
tests/golden/make_golden_loop_i2v.py) -- sphere views over wrapping frame windows, latent-shaped mask panorama,
re-noise mix, merge_prev mix, one condition-image view per latent view -- through the loop driver and through the
drop-in classes."""
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


def load():
    g = np.load(os.path.join(GOLDEN, "loop_i2v.npz"))
    names = ("frames", "total_f", "H", "W", "h", "w", "steps", "fov", "loop_step_frame", "loop_step_theta")
    c = {k: int(v) for k, v in zip(names, g["cfg"])}
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    views0 = [(float(p), float(t)) for p, t in g["views"]]
    views_fn = lambda i: [(p, t + (i % c["loop_step_theta"]) * (c["fov"] // c["loop_step_theta"])) for p, t in views0]
    fb_fn = lambda i: po.frame_window_begins(i, frames=c["frames"], total_f=c["total_f"], overlap=float(g["overlap_f"]),
                                             loop_step_frame=c["loop_step_frame"])
    image = ((g["image_u8"].astype(np.float32).transpose(2, 0, 1) / np.float32(255.)) - np.float32(0.5)) * np.float32(2)
    assert np.array_equal(sha(image), g["image_f32_sha"])
    nv = len(views0)
    base = [0] + [int(v) for v in np.cumsum([len(fb_fn(i)) * nv for i in range(c["steps"])])]
    return g, c, sched, views_fn, fb_fn, image, base


def test_i2v_sphere_loop_replays_the_real_reference_pipeline(dev):
    """3 DDIM steps x 3 wrapping frame windows x 6 views of the reference's own sequential I2V sphere loop: every UNet
    input, every condition-image view (sha256) and both returned panoramas equal the real pipeline's, bit for bit."""
    from dynamicscaler_b200.geometry import frame_windows
    from dynamicscaler_b200.loop import sphere_denoise_loop
    from dynamicscaler_b200.step import PanoramaStepper
    g, c, sched, views_fn, fb_fn, image, base = load()
    F, H, W, h, w, frames = c["total_f"], c["H"], c["W"], c["h"], c["w"], c["frames"]
    for i in range(c["steps"]):
        assert frame_windows(i, frames=frames, total_f=F, overlap=float(g["overlap_f"]),
                             loop_step_frame=c["loop_step_frame"]) == fb_fn(i)
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = G(randn(101, 1, 4, F, H, W), dev)
    state = st.make_state([lat0, torch.zeros_like(lat0)], mask=torch.zeros_like(lat0))      # latent-shaped mask
    d_image, text = G(image, dev), G(g["text_emb"], dev)
    cur = {"calls": 0, "draws": 0}

    def cond_fn(i, job_ids, boxes):
        views = st.gather_image_views(d_image, [(phi, th) for (phi, th, _) in boxes], w * 8, h * 8, fov=c["fov"])
        for q, job in enumerate(job_ids):
            assert np.array_equal(sha(N(views[q])), g["img_view_sha"][base[i] + job]), (i, job)
        emb = torch.cat([G(g["img_emb"][base[i] + job], dev) for job in job_ids], dim=0)
        return {"c_crossattn": [torch.cat([text.expand(len(job_ids), -1, -1), emb], dim=1)], "fps": 16}

    def on_input(i, job_ids, xb):
        cur.update(i=i, jobs=list(job_ids), branch=0)
        for q, job in enumerate(job_ids):
            assert np.array_equal(sha(N(xb[q:q + 1])), g["unet_in_sha"][base[i] + job]), (i, job)

    def replay_model(x, ts, c_crossattn=None, fps=None, **kw_):
        src = g["eps_cond"] if cur["branch"] == 0 else g["eps_uncond"]
        assert c_crossattn[0].shape[1] == (77 + 4 if cur["branch"] == 0 else 77) and kw_["temporal_length"] == frames
        cur["branch"] += 1
        cur["calls"] += x.shape[0]
        return torch.cat([G(src[base[cur["i"]] + job], dev) for job in cur["jobs"]], dim=0)

    def eps_fn(i, job_ids, shape, device):
        cur["draws"] += 1
        return G(randn(8000 + cur["draws"], *shape[1:]), device).view(shape)

    sphere_denoise_loop(st, state, model=replay_model, schedule=sched, views_fn=views_fn, width=w, height=h, frames=frames,
                        cond=cond_fn, uncond={"c_crossattn": [G(g["uncond_emb"], dev)], "fps": 16},
                        guidance_scale=float(g["guidance"]), fov=c["fov"], group_size=1, renoise_ratio=float(g["ratio"]),
                        prev_ratio_list=[float(v) for v in g["prev_ratios"]], eps_fn=eps_fn, cfg_batch=False,
                        frame_begins_fn=fb_fn, on_input=on_input)
    assert cur["calls"] == 2 * base[-1] and cur["draws"] == int(g["renoise_draws"])
    out_lat, out_den = state.export()
    assert np.array_equal(N(out_lat), g["final_latent"])
    assert np.array_equal(N(out_den), g["final_denoised"])


def test_i2v_sphere_loop_dropin_classes_replay_the_real_reference_pipeline(dev):
    """The reference's loop body (pipeline/i2v_sphere_panorama_pipeline.py:284-470) call for call through the DROP-IN
    ``RingPanoramaLatentProxy`` / ``PanoramaTensor`` / ``mix_latents_with_mask``."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaTensor
    from dynamicscaler_b200.ring_panorama_tensor_utils import RingPanoramaLatentProxy
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    g, c, sched, views_fn, fb_fn, image, base = load()
    F, H, W, h, w, frames, fov = c["total_f"], c["H"], c["W"], c["h"], c["w"], c["frames"], c["fov"]
    nv = len(g["views"])
    lat0 = G(randn(101, 1, 4, F, H, W), dev)
    h_lat = RingPanoramaLatentProxy(equirect_tensor=lat0)
    h_den = RingPanoramaLatentProxy(equirect_tensor=torch.zeros_like(lat0))
    h_img = PanoramaTensor(equirect_tensor=torch.from_numpy(image))                 # a CPU tensor, as in the reference (:223-224)
    ac = torch.from_numpy(np.asarray(sched["alphas_cumprod"], dtype=np.float32)).to(dev)
    ddim_ts = [int(v) for v in sched["ddim_timesteps"]]
    total, ratio, gscale = c["steps"], float(g["ratio"]), float(g["guidance"])
    draws = 0
    for i in range(total):
        idx = total - i - 1
        f32 = lambda v: torch.full((1, 1, 1, 1, 1), float(v), device=dev)
        a_t, a_prev, sig, s1m = (f32(sched[k][idx]) for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                                                              "ddim_sqrt_one_minus_alphas"))
        h_msk = RingPanoramaLatentProxy(equirect_tensor=torch.zeros_like(lat0))          # :242
        n = 0
        for fb in fb_fn(i):
            fkw = dict(frame_begin=fb, frame_end=fb + frames)
            for (phi, th) in views_fn(i):
                x, _ = h_lat.get_view_tensor_no_interpolate(fov=fov, theta=th, phi=phi, width=w, height=h, **fkw)
                prev = x.clone()
                m, _ = h_msk.get_view_tensor_no_interpolate(fov=fov, theta=th, phi=phi, width=w, height=h, **fkw)
                if i < total - 1:                                                         # :355-365
                    a_a, a_b = ac[ddim_ts[total - i - 2]], ac[ddim_ts[total - i - 1]]
                    cc, ss = torch.sqrt(a_b / a_a), torch.sqrt(1 - a_b / a_a)
                    draws += 1
                    noised = cc * x.clone() + ss * G(randn(8000 + draws, *x.shape), dev)
                    x = mix_latents_with_mask(latent_1=x, latent_to_add=noised, mask=m, mix_ratio=ratio)
                j = base[i] + n
                assert np.array_equal(sha(N(x.contiguous())), g["unet_in_sha"][j]), (i, n)
                img, _ = h_img.get_view_tensor_no_interpolate(fov=fov, theta=th, phi=phi, width=w * 8, height=h * 8)
                assert np.array_equal(sha(N(img.contiguous())), g["img_view_sha"][j]), (i, n)
                ec, eu = G(g["eps_cond"][j], dev), G(g["eps_uncond"][j], dev)
                e = eu + gscale * (ec - eu)
                pred_x0 = (x - s1m * e) / a_t.sqrt()
                x_prev = a_prev.sqrt() * pred_x0 + (1. - a_prev - sig ** 2).sqrt() * e
                if i < total - 1:                                                         # :430-435
                    x_prev = mix_latents_with_mask(latent_1=x_prev, latent_to_add=prev, mask=m,
                                                   mix_ratio=float(g["prev_ratios"][i]))
                h_lat.set_view_tensor_no_interpolation(view_tensor=x_prev, fov=fov, theta=th, phi=phi, **fkw)
                h_den.set_view_tensor_no_interpolation(view_tensor=pred_x0, fov=fov, theta=th, phi=phi, **fkw)
                h_msk.set_view_tensor_no_interpolation(view_tensor=torch.ones_like(x_prev), fov=fov, theta=th, phi=phi, **fkw)
                n += 1
    assert draws == int(g["renoise_draws"])
    assert np.array_equal(N(h_lat.get_equirect_tensor()), g["final_latent"])
    assert np.array_equal(N(h_den.get_equirect_tensor()), g["final_denoised"])
