"""The oracle (oracle/panorama_oracle.py) against outputs of the imported reference
(tests/golden/*.npz, made by tests/golden/make_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

import torch

from conftest import GOLDEN, none_if_neg, randn
from oracle import panorama_oracle as po


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_ring_segments_and_index(small):
    for i, (b, e, s) in enumerate(small["slices/cases"]):
        want = small[f"slices/{i}"]
        assert np.array_equal(po.ring_index(b, e, s), want)
        segs = po.ring_segments(int(b), int(e), int(s))
        assert np.array_equal(np.concatenate([np.arange(st, st + ln) for st, ln in segs]), want)


def test_ring_get_set(small):
    lat0 = randn(11, *small["ring/shape"])
    for i, c in enumerate(small["ring/get_cases"]):
        assert np.array_equal(po.ring_get(lat0, *none_if_neg(c)), small[f"ring/get/{i}"])
    for i, c in enumerate(small["ring/set_cases"]):
        lat = lat0.copy()
        po.ring_set(lat, randn(100 + i, *po.ring_get(lat0, *none_if_neg(c)).shape), *none_if_neg(c))
        assert np.array_equal(lat, small[f"ring/set/{i}"])


def test_ring_illegal_windows():
    lat = np.zeros((1, 1, 4, 8, 16), np.float32)
    with pytest.raises(AssertionError):
        po.ring_get(lat, 20, 20)            # left < right
    with pytest.raises(AssertionError):
        po.ring_get(lat, 0, 33)             # right <= 2W
    with pytest.raises(AssertionError):
        po.ring_set(lat, np.zeros((1, 1, 4, 8, 17), np.float32), 0, 17)   # wrap onto itself


def test_image_ring(small):
    img0 = randn(12, 3, 8, 12)
    for i, c in enumerate(small["img/cases"]):
        assert np.array_equal(po.image_ring_get(img0, *none_if_neg(c)), small[f"img/get/{i}"])


def test_uv_map_bit_exact(small):
    for i, (fov, th, ph, w, h, H, W) in enumerate(small["uv/cases"]):
        u, v = po.uv_map(fov, th, ph, int(w), int(h), int(H), int(W))
        if f"uv/{i}/u" in small:
            assert np.array_equal(u, small[f"uv/{i}/u"]) and np.array_equal(v, small[f"uv/{i}/v"])
        else:
            assert np.array_equal(sha(u), small[f"uv/{i}/u_sha"]) and np.array_equal(sha(v), small[f"uv/{i}/v_sha"])


def test_sphere_small(small):
    pano5 = randn(21, 1, 2, 3, 16, 32)
    for i, (fov, th, ph) in enumerate(small["sph/views"]):
        view_in = randn(300 + i, 1, 2, 3, 6, 10)
        vw, m = po.proxy_get_nearest(pano5, fov, th, ph, 10, 6)
        assert np.array_equal(vw, small[f"sph/get_nearest/{i}"])
        assert np.array_equal(m, small[f"sph/get_nearest_mask/{i}"])
        got = po.proxy_get_bilinear(pano5, fov, th, ph, 10, 6)
        assert np.abs(got - small[f"sph/get_bilinear/{i}"]).max() <= 1e-6
        assert np.abs(po.sphere_get_bilinear_torch(pano5, fov, th, ph, 10, 6) - small[f"sph/get_bilinear/{i}"]).max() == 0
        for name, fn in (("set_nearest", po.proxy_set_nearest), ("set_bilinear", po.proxy_set_bilinear),
                         ("set_round", po.proxy_set_round)):
            lat = pano5.copy()
            fn(lat, view_in, fov, th, ph)
            assert np.array_equal(lat, small[f"sph/{name}/{i}"]), (name, i)


def test_sphere_image_and_mask(small):
    img = randn(22, 3, 16, 32)
    for i, (fov, th, ph) in enumerate(small["sph/views"][:3]):
        assert np.array_equal(po.sphere_get_nearest(img, fov, th, ph, 10, 6)[0], small[f"sph/img_get/{i}"])
        assert np.abs(po.sphere_get_bilinear(img, fov, th, ph, 10, 6) - small[f"sph/img_get_bilinear/{i}"]).max() <= 1e-6
        for name, fn in (("img_set", po.sphere_set_nearest), ("img_set_round", po.sphere_set_round),
                         ("img_set_bilinear", po.sphere_set_bilinear)):
            a = img.copy()
            fn(a, randn(400 + i, 3, 6, 10), fov, th, ph)
            assert np.array_equal(a, small[f"sph/{name}/{i}"]), (name, i)
        m = np.zeros((1, 16, 32), np.float32)
        po.sphere_set_nearest(m, np.ones((1, 6, 10), np.float32), fov, th, ph)
        assert np.array_equal(m, small[f"sph/mask_set/{i}"])


def test_frame_ring_sphere(small):
    rp5 = randn(31, 1, 2, 7, 16, 32)
    for i, (fov, th, ph, fb, fe) in enumerate(small["rsph/cases"]):
        fb, fe = int(fb), int(fe)
        view_in = randn(500 + i, 1, 2, fe - fb, 6, 10)
        assert np.array_equal(po.proxy_get_nearest(rp5, fov, th, ph, 10, 6, fb, fe)[0], small[f"rsph/get_nearest/{i}"])
        assert np.abs(po.proxy_get_bilinear(rp5, fov, th, ph, 10, 6, fb, fe) - small[f"rsph/get_bilinear/{i}"]).max() <= 1e-6
        for name, fn in (("set_nearest", po.proxy_set_nearest), ("set_bilinear", po.proxy_set_bilinear),
                         ("set_round", po.proxy_set_round)):
            if f"rsph/{name}/{i}" not in small:
                continue
            lat = rp5.copy()
            fn(lat, view_in, fov, th, ph, fb, fe)
            assert np.array_equal(lat, small[f"rsph/{name}/{i}"]), (name, i)


def test_mix_mask(small):
    l1, add = randn(41, 1, 4, 5, 6, 10), randn(42, 1, 4, 5, 6, 10)
    m3 = (randn(43, 1, 6, 10) > 0).astype(np.float32)
    m5 = (randn(44, 1, 4, 5, 6, 10) > 0).astype(np.float32)
    for i, (m, r) in enumerate([(m3, 0.5), (m3, 0.3), (m5, 0.8), (m5, 1.0), (m3, 0.0)]):
        assert np.array_equal(po.mix_mask(l1, add, m, r), small[f"mix/{i}"])
    with pytest.raises(NotImplementedError):
        po.mix_mask(l1, add, m3[0], 0.5)


def test_last_writer_flags():
    p = np.array([5, 3, 5, 7, 3, 3, 9])
    assert po.last_writer_flags(p).tolist() == [False, False, True, True, False, True, True]


def _payload(n, win):
    return (win * np.float32(0.5) + randn(700 + n, *win.shape), win * np.float32(0.25) - randn(800 + n, *win.shape))


def test_step_sequential_small(small):
    """The reference's own window-after-window order (t2v_sphere_panorama_pipeline.py:176-291)."""
    lat = randn(51, 1, 2, 3, 16, 32)
    den = np.zeros_like(lat)
    msk = np.zeros((1, 16, 32), np.float32)
    for n, (ph, th) in enumerate(small["step/views"]):
        vw, _ = po.proxy_get_nearest(lat, 120, th, ph, 10, 6)
        mk, _ = po.sphere_get_nearest(msk, 120, th, ph, 10, 6)
        noised = vw * np.float32(0.9) + randn(600 + n, 1, 2, 3, 6, 10) * np.float32(0.4359)
        vw = po.mix_mask(vw, noised, mk, 0.6)
        x_prev, x0 = _payload(n, vw)
        po.proxy_set_nearest(lat, x_prev, 120, th, ph)
        po.proxy_set_nearest(den, x0, 120, th, ph)
        po.sphere_set_nearest(msk, np.ones((1, 6, 10), np.float32), 120, th, ph)
    assert np.array_equal(lat, small["step/seq/latent"])
    assert np.array_equal(den, small["step/seq/denoised"])
    assert np.array_equal(msk, small["step/seq/mask"])


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_step_batched_small(small, mode):
    lat = randn(51, 1, 2, 3, 16, 32)
    den = np.zeros_like(lat)
    views = [(ph, th) for ph, th in small["step/views"]]
    wins, mask = po.sphere_step_batched(lat, den, views, _payload, fov=120, width=10, height=6, mode=mode)
    tol = 0 if mode == "nearest" else 1e-6
    assert np.abs(np.stack(wins) - small[f"step/batched_{mode}/windows"]).max() <= tol
    assert np.abs(lat - small[f"step/batched_{mode}/latent"]).max() <= tol * 10
    assert np.abs(den - small[f"step/batched_{mode}/denoised"]).max() <= tol * 10
    assert np.array_equal(mask[0, 0, 0][None] if mode == "nearest" else mask[0, 0, 0][None],
                          small[f"step/batched_{mode}/mask"])


def test_config1_ring_digest(configs):
    lat = randn(0, 1, 4, 16, 64, 128)
    den = np.zeros_like(lat)
    windows = [(int(l), int(t), 0) for l, t in configs["cfg1/windows"]]
    assert windows[:9] == [(l, t, 0) for l, t in
                           [(12 + 8 * a, 3 + 8 * b) for a in range(16) for b in range(8)][:9]]
    pay = lambda n, w: (randn(1000 + n, *w.shape), randn(2000 + n, *w.shape))
    wins, mask = po.ring_step_batched(lat, den, windows, pay)
    assert np.array_equal(sha(np.stack(wins)), configs["cfg1/windows_sha"])
    assert np.array_equal(sha(lat), configs["cfg1/latent_sha"])
    assert np.array_equal(sha(den), configs["cfg1/denoised_sha"])
    assert np.array_equal(sha(np.broadcast_to(mask, lat.shape)), configs["cfg1/mask_sha"])


def test_config2_views_and_index_maps(configs):
    views = po.default_sphere_views(1)
    assert len(views) == 44
    assert np.array_equal(np.array(views, dtype=np.float64), configs["cfg2/views"])
    for n, (ph, th) in enumerate(views):
        u, v = po.uv_map(120, th, ph, 64, 40, 128, 256)
        assert np.array_equal(po.nearest_read_index(u, v, 128, 256), configs["cfg2/idx_read"][n])
        assert np.array_equal(po.nearest_write_index(u, v, 128, 256)[0], configs["cfg2/idx_write"][n])


def test_config2_nearest_digest(configs):
    lat = randn(0, 1, 4, 16, 128, 256)
    den = np.zeros_like(lat)
    pay = lambda n, w: (randn(1000 + n, *w.shape), randn(2000 + n, *w.shape))
    wins, mask = po.sphere_step_batched(lat, den, po.default_sphere_views(1), pay, mode="nearest")
    assert np.array_equal(sha(np.stack(wins)), configs["cfg2/nearest/windows_sha"])
    assert np.array_equal(sha(lat), configs["cfg2/nearest/latent_sha"])
    assert np.array_equal(sha(den), configs["cfg2/nearest/denoised_sha"])
    assert np.array_equal(sha(mask[0, 0, 0][None]), configs["cfg2/nearest/mask_sha"])


def test_config2_bilinear(configs):
    lat = randn(0, 1, 4, 16, 128, 256)
    den = np.zeros_like(lat)
    pay = lambda n, w: (randn(1000 + n, *w.shape), randn(2000 + n, *w.shape))
    wins, mask = po.sphere_step_batched(lat, den, po.default_sphere_views(1), pay, mode="bilinear")
    s = lambda a: a.reshape(-1)[::4099]
    assert np.abs(s(np.stack(wins)) - configs["cfg2/bilinear/windows_sample"]).max() <= 1e-6
    assert np.array_equal(sha(lat), configs["cfg2/bilinear/latent_sha"])
    assert np.array_equal(sha(den), configs["cfg2/bilinear/denoised_sha"])


def test_config4_frame_ring_digest(configs):
    lat = randn(4, 1, 4, 64, 128, 256)
    jobs = [(int(fb), ph, th) for fb, ph, th in configs["cfg4/jobs"]]
    wins = [po.proxy_get_nearest(lat, 120, th, ph, 64, 40, fb, fb + 16)[0] for fb, ph, th in jobs]
    for n, (fb, ph, th) in enumerate(jobs):
        po.proxy_set_nearest(lat, randn(1000 + n, 1, 4, 16, 40, 64), 120, th, ph, fb, fb + 16)
    assert np.array_equal(sha(np.stack(wins)), configs["cfg4/windows_sha"])
    assert np.array_equal(sha(lat), configs["cfg4/latent_sha"])


def test_geometry_helpers():
    # SURVEY A.9
    assert po.ring_window_grid(0, 2, 2, 1024, 512) == [(0, 0), (0, 24), (64, 0), (64, 24)]
    g3 = po.ring_window_grid(0, 4, 4, 2048, 1024)
    assert sorted({l for l, _ in g3}) == [0, 64, 128, 192] and sorted({t for _, t in g3}) == [0, 29, 59, 88]
    g5 = po.ring_window_grid(1, 16, 16, 8192, 4096)
    assert g5[0] == (4, 1) and len(g5) == 256
    assert po.frame_window_begins(0) == [(4 * s) % 64 for s in range(13)]
    assert po.frame_window_begins(3)[0] == 3
    assert len(po.frame_window_begins(0, overlap=0.5)) == 7


def test_resize_video_latent_golden():
    """oracle.resize_video_latent vs the reference's utils/diffusion_utils.py outputs
    (tests/golden/resize.npz, generated by tests/golden/make_golden_resize.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize.npz"))
    modes = ["nearest", "bilinear", "bicubic"]
    for n, row in enumerate(g["cases"]):
        seed, shape, th, tw, mode, align = int(row[0]), tuple(int(v) for v in row[1:6]), int(row[6]), int(row[7]), modes[row[8]], bool(row[9])
        got = po.resize_video_latent(randn(seed, *shape), th, tw, mode=mode, align_corners=align)
        want = g[f"out/{n}"]
        assert got.shape == want.shape
        if mode == "nearest":
            assert np.array_equal(got, want), n
        else:
            assert np.abs(got - want).max() <= 1e-5, (n, np.abs(got - want).max())


def test_ddim_step_cfg_golden():
    """oracle.ddim_step_cfg vs the reference's scheduler.ddim_step + CFG combine (tests/golden/ddim.npz)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddim.npz"))
    shape = tuple(int(v) for v in g["shape"])
    for n, (eta, index, gs, a_t, a_prev, sig, s1m) in enumerate(g["cases"]):
        x, ec, eu, nz = (randn(s + n, *shape) for s in (100, 200, 300, 400))
        x_prev, pred_x0 = po.ddim_step_cfg(x, ec, eu if gs != 1.0 else None, gs, a_t, a_prev, sig, s1m, noise=nz)
        assert np.array_equal(pred_x0, g[f"pred_x0/{n}"]), n
        assert np.array_equal(x_prev, g[f"x_prev/{n}"]), n


def test_frame_ring_port_matches_reference(small):
    """oracle/torch_port.FrameRingSpherePano (the timed CPU baseline of cfg4) == the reference's
    RingPanoramaLatentProxy on every frame-ring golden, bit for bit."""
    from oracle import torch_port as tp
    rp5 = randn(31, 1, 2, 7, 16, 32)
    for i, (fov, th, ph, fb, fe) in enumerate(small["rsph/cases"]):
        fb, fe = int(fb), int(fe)
        p = tp.FrameRingSpherePano(torch.from_numpy(rp5))
        assert np.array_equal(p.get_nearest(fov, th, ph, 10, 6, fb, fe)[0].numpy(), small[f"rsph/get_nearest/{i}"])
        assert np.array_equal(p.get_bilinear(fov, th, ph, 10, 6, fb, fe).numpy(), small[f"rsph/get_bilinear/{i}"])
        vin = torch.from_numpy(randn(500 + i, 1, 2, fe - fb, 6, 10))
        p = tp.FrameRingSpherePano(torch.from_numpy(rp5))
        p.set_nearest(vin, fov, th, ph, fb, fe)
        assert np.array_equal(p.latent().numpy(), small[f"rsph/set_nearest/{i}"])
        p = tp.FrameRingSpherePano(torch.from_numpy(rp5))
        p.set_bilinear(vin, fov, th, ph, fb, fe)
        assert np.array_equal(p.latent().numpy(), small[f"rsph/set_bilinear/{i}"])


def loop_golden():
    return np.load(os.path.join(GOLDEN, "loop.npz"))


def loop_inputs(g):
    """(schedule dict, step-0 views, views_fn, eps_fn, model_fn factory) of tests/golden/loop.npz."""
    F, H, W, h, w, steps, fov = [int(v) for v in g["cfg"]]
    views0 = [(float(p), float(t)) for p, t in g["views"]]
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    views_fn = lambda i: [(p, t + (i % 10) * (fov // 10)) for p, t in views0]
    return sched, views0, views_fn


def test_denoise_loop_oracle_matches_real_pipeline():
    """oracle.sphere_denoise_loop in the reference's sequential order == the REAL
    VC2_Pipeline_T2V_SpherePano.basic_sample_shift_shpere_panorama run (tests/golden/loop.npz:
    3 DDIM steps x 6 views, re-noise mix on, reduced-width reference UNet): every UNet input and both
    final panoramas bit for bit, with the recorded UNet outputs replayed."""
    import hashlib
    g = loop_golden()
    F, H, W, h, w, steps, fov = [int(v) for v in g["cfg"]]
    sched, views0, views_fn = loop_inputs(g)
    nv = len(views0)
    draws = {"n": 0}

    def eps_fn(i, n, shape):
        draws["n"] += 1
        return randn(5000 + draws["n"], *shape)

    def model_fn(i, n, x):
        j = i * nv + n
        got = np.frombuffer(hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(got, g["unet_in_sha"][j]), (i, n)
        return g["eps_cond"][j], g["eps_uncond"][j]

    lat, den, _, _ = po.sphere_denoise_loop(randn(41, 1, 4, F, H, W), sched, views_fn, model_fn, eps_fn,
                                            float(g["guidance"]), fov=fov, width=w, height=h,
                                            renoise_ratio=float(g["ratio"]), group_size=1)
    assert draws["n"] == int(g["renoise_draws"])
    assert np.array_equal(lat, g["final_latent"]) and np.array_equal(den, g["final_denoised"])
    assert float(g["p9_batched_vs_looped_maxabs"]) < 1e-4      # P9: one UNet call over all windows == looped calls


def ring_loop_golden():
    g = np.load(os.path.join(GOLDEN, "ring_loop.npz"))
    frames, h, w, nw, nh, nf, loop_step, steps = [int(v) for v in g["cfg"]]
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    kw = dict(frames=frames, latent_height=h, latent_width=w, num_windows_f=nf, num_windows_h=nh, num_windows_w=nw,
              loop_step=loop_step, shift_jump_odd_w=True, docking_w=True, docking_step_range=[int(v) for v in g["dock_steps"]])
    return g, (frames, h, w, nw, nh, nf, steps), sched, kw


def test_ring_loop_oracle_matches_real_pipeline():
    """oracle.ring_denoise_loop + oracle.shift_window_jobs == the REAL ``VC2_Pipeline_T2V.
    basic_sample_shift_multi_windows`` run (tests/golden/ring_loop.npz: 3 DDIM steps, 3 x 2 x 2 windows,
    odd-step jump, docking with overlapping windows at the last step): every UNet input, the latent and
    denoised panoramas after EVERY step and the returned denoised panorama, bit for bit -- in the
    reference's one-window-at-a-time order, and with whole-step groups on the steps whose windows tile the
    torus (the batched order the driver runs)."""
    import hashlib
    sha = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    g, (frames, h, w, nw, nh, nf, steps), sched, kw = ring_loop_golden()
    jobs_fn = lambda i: po.shift_window_jobs(i, **kw)
    counts = [len(jobs_fn(i)) for i in range(steps)]
    assert counts == [12, 12, 16] and sum(counts) == len(g["unet_in_sha"])
    base = np.concatenate([[0], np.cumsum(counts)])
    dock = [int(v) for v in g["dock_steps"]]

    def model_fn(i, n, x):
        j = int(base[i]) + n
        assert np.array_equal(sha(x), g["unet_in_sha"][j]), (i, n)
        return g["eps_cond"][j], g["eps_uncond"][j]

    def on_step(i, lat, den):
        assert np.array_equal(sha(lat), g["step_latent_sha"][i]), i
        assert np.array_equal(sha(den), g["step_denoised_sha"][i]), i

    lat0 = randn(51, 1, 4, frames * nf, h * nh, w * nw)
    for groups_fn in (None, lambda i, n: [[k] for k in range(n)] if i in dock else [list(range(n))]):
        _, den, _ = po.ring_denoise_loop(lat0, sched, jobs_fn, model_fn, float(g["guidance"]), w, h, frames,
                                         groups_fn=groups_fn, on_step=on_step)
        assert np.array_equal(den, g["final_denoised"])
    # ... and whole-step groups on the docked step do NOT reproduce it (overlapping windows read stale data)
    with pytest.raises(AssertionError):
        po.ring_denoise_loop(lat0, sched, jobs_fn, model_fn, float(g["guidance"]), w, h, frames,
                             groups_fn=lambda i, n: [list(range(n))], on_step=on_step)


def ring_loop_i2v_golden():
    g = np.load(os.path.join(GOLDEN, "ring_loop_i2v.npz"))
    names = ("frames", "height", "width", "total_h", "total_w", "total_f", "nw", "nh", "loop_step", "steps",
             "loop_step_frame", "begin_index_offset")
    c = {k: int(v) for k, v in zip(names, g["cfg"])}
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    jobs_kw = dict(width=c["width"], height=c["height"], frames=c["frames"], total_w=c["total_w"], total_h=c["total_h"],
                   total_f=c["total_f"], num_windows_w=c["nw"], num_windows_h=c["nh"], loop_step=c["loop_step"],
                   overlap_ratio_f=float(g["overlap_f"]), loop_step_frame=c["loop_step_frame"], dock_at_f=True,
                   begin_index_offset=c["begin_index_offset"])
    return g, c, sched, jobs_kw


def test_i2v_ring_loop_oracle_matches_real_pipeline():
    """oracle.ring_denoise_loop (mask panorama + re_noise mix + merge_prev mix) over oracle.i2v_planar_jobs == the REAL
    ``VC2_Pipeline_I2V_SpherePano.basic_sample_shift_multi_windows`` run (tests/golden/ring_loop_i2v.npz: 3 DDIM steps,
    3 x 3 spatial windows overlapping on H, 3-4 frame windows with docking, 90 windows): every UNet input, every
    condition-image window, the panoramas after every step and the returned denoised panorama, bit for bit."""
    import hashlib
    sha = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    g, c, sched, jobs_kw = ring_loop_i2v_golden()
    jobs_fn = lambda i: po.i2v_planar_jobs(i, **jobs_kw)
    counts = [len(jobs_fn(i)) for i in range(c["steps"])]
    assert counts == [27, 36, 27] and sum(counts) == len(g["unet_in_sha"])
    base = np.concatenate([[0], np.cumsum(counts)])
    h, w, frames = c["height"] // 8, c["width"] // 8, c["frames"]
    image = ((g["image_u8"].astype(np.float32).transpose(2, 0, 1) / np.float32(255.)) - np.float32(0.5)) * np.float32(2)
    assert np.array_equal(sha(image), g["image_f32_sha"])
    draws = {"n": 0}

    def eps_fn(i, n, shape):
        draws["n"] += 1
        return randn(6000 + draws["n"], *shape)

    def model_fn(i, n, x):
        j = int(base[i]) + n
        assert np.array_equal(sha(x), g["unet_in_sha"][j]), (i, n)
        left, top, _ = jobs_fn(i)[n]
        win = po.image_ring_get(image, left * 8, left * 8 + c["width"], top * 8, top * 8 + c["height"])   # :814-817
        assert np.array_equal(sha(win), g["img_win_sha"][j]), (i, n)
        return g["eps_cond"][j], g["eps_uncond"][j]

    def on_step(i, lat, den):
        assert np.array_equal(sha(lat), g["step_latent_sha"][i]), i
        assert np.array_equal(sha(den), g["step_denoised_sha"][i]), i

    lat0 = randn(81, 1, 4, c["total_f"], c["total_h"] // 8, c["total_w"] // 8)
    _, den, _ = po.ring_denoise_loop(lat0, sched, jobs_fn, model_fn, float(g["guidance"]), w, h, frames, on_step=on_step,
                                     renoise_ratio=float(g["ratio"]), prev_ratio_list=[float(v) for v in g["prev_ratios"]],
                                     eps_fn=eps_fn, use_mask=True)
    assert draws["n"] == int(g["renoise_draws"])
    assert np.array_equal(den, g["final_denoised"])


def ring_loop_t2v_golden():
    g = np.load(os.path.join(GOLDEN, "ring_loop_t2v.npz"))
    names = ("frames", "height", "width", "total_h", "total_w", "nw", "nh", "nf", "loop_step", "steps")
    c = {k: int(v) for k, v in zip(names, g["cfg"])}
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    jobs_kw = dict(width=c["width"], height=c["height"], frames=c["frames"], total_w=c["total_w"], total_h=c["total_h"],
                   num_windows_w=c["nw"], num_windows_h=c["nh"], num_windows_f=c["nf"], loop_step=c["loop_step"],
                   dock_at_h=True)
    return g, c, sched, jobs_kw


def test_t2v_ring_loop_oracle_matches_real_pipeline():
    """oracle.ring_denoise_loop (mask panorama, re-noise mix with the mask window's [0, 0, [0]] plane) over
    oracle.t2v_planar_jobs == the REAL ``VC2_Pipeline_T2V_SpherePano.basic_sample_shift_multi_windows`` run
    (tests/golden/ring_loop_t2v.npz: 3 DDIM steps, 3 x 3 overlapping windows + windows docked at the top / bottom
    edge, 2 frame windows; 66 windows): every UNet input, the panoramas after every step and the returned denoised
    panorama, bit for bit."""
    import hashlib
    sha = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    g, c, sched, jobs_kw = ring_loop_t2v_golden()
    jobs_fn = lambda i: po.t2v_planar_jobs(i, **jobs_kw)
    counts = [len(jobs_fn(i)) for i in range(c["steps"])]
    assert counts == [18, 24, 24] and sum(counts) == len(g["unet_in_sha"])
    base = np.concatenate([[0], np.cumsum(counts)])
    draws = {"n": 0}

    def eps_fn(i, n, shape):
        draws["n"] += 1
        return randn(7000 + draws["n"], *shape)

    def model_fn(i, n, x):
        j = int(base[i]) + n
        assert np.array_equal(sha(x), g["unet_in_sha"][j]), (i, n)
        return g["eps_cond"][j], g["eps_uncond"][j]

    def on_step(i, lat, den):
        assert np.array_equal(sha(lat), g["step_latent_sha"][i]), i
        assert np.array_equal(sha(den), g["step_denoised_sha"][i]), i

    lat0 = randn(95, 1, 4, c["frames"] * c["nf"], c["total_h"] // 8, c["total_w"] // 8)
    _, den, _ = po.ring_denoise_loop(lat0, sched, jobs_fn, model_fn, float(g["guidance"]), c["width"] // 8, c["height"] // 8,
                                     c["frames"], on_step=on_step, renoise_ratio=float(g["ratio"]), eps_fn=eps_fn,
                                     use_mask=True, mask_frame0=True)
    assert draws["n"] == int(g["renoise_draws"])
    assert np.array_equal(den, g["final_denoised"])


def loop_i2v_golden():
    g = np.load(os.path.join(GOLDEN, "loop_i2v.npz"))
    names = ("frames", "total_f", "H", "W", "h", "w", "steps", "fov", "loop_step_frame", "loop_step_theta")
    c = {k: int(v) for k, v in zip(names, g["cfg"])}
    sched = {k: g[k] for k in ("ddim_timesteps", "alphas_cumprod", "ddim_alphas", "ddim_alphas_prev", "ddim_sigmas",
                               "ddim_sqrt_one_minus_alphas")}
    views0 = [(float(p), float(t)) for p, t in g["views"]]
    views_fn = lambda i: [(p, t + (i % c["loop_step_theta"]) * (c["fov"] // c["loop_step_theta"])) for p, t in views0]
    fb_fn = lambda i: po.frame_window_begins(i, frames=c["frames"], total_f=c["total_f"], overlap=float(g["overlap_f"]),
                                             loop_step_frame=c["loop_step_frame"])
    return g, c, sched, views_fn, fb_fn


def test_i2v_sphere_loop_oracle_matches_real_pipeline():
    """oracle.sphere_denoise_loop with wrapping frame windows, the latent-shaped mask panorama, the re-noise mix and the
    merge_prev mix == the REAL ``VC2_Pipeline_I2V_SpherePano.basic_sample_shift_shpere_panorama`` run
    (tests/golden/loop_i2v.npz: 3 DDIM steps x 3 frame windows x 6 views): every UNet input, every condition-image view
    (``PanoramaTensor`` nearest read of the [3, 128, 256] image) and both returned panoramas, bit for bit."""
    import hashlib
    sha = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)
    g, c, sched, views_fn, fb_fn = loop_i2v_golden()
    nv = len(g["views"])
    counts = [len(fb_fn(i)) * nv for i in range(c["steps"])]
    assert counts == [18, 18, 18] and sum(counts) == len(g["unet_in_sha"])
    base = np.concatenate([[0], np.cumsum(counts)])
    image = ((g["image_u8"].astype(np.float32).transpose(2, 0, 1) / np.float32(255.)) - np.float32(0.5)) * np.float32(2)
    assert np.array_equal(sha(image), g["image_f32_sha"])
    draws = {"n": 0}

    def eps_fn(i, n, shape):
        draws["n"] += 1
        return randn(8000 + draws["n"], *shape)

    def model_fn(i, n, x):
        j = int(base[i]) + n
        assert np.array_equal(sha(x), g["unet_in_sha"][j]), (i, n)
        phi, th = views_fn(i)[n % nv]
        view, _ = po.sphere_get_nearest(image, c["fov"], th, phi, c["w"] * 8, c["h"] * 8)          # i2v :367-372
        assert np.array_equal(sha(view), g["img_view_sha"][j]), (i, n)
        return g["eps_cond"][j], g["eps_uncond"][j]

    lat, den, _, _ = po.sphere_denoise_loop(randn(101, 1, 4, c["total_f"], c["H"], c["W"]), sched, views_fn, model_fn, eps_fn,
                                            float(g["guidance"]), fov=c["fov"], width=c["w"], height=c["h"],
                                            renoise_ratio=float(g["ratio"]), prev_ratio_list=[float(v) for v in g["prev_ratios"]],
                                            group_size=1, frame_begins_fn=fb_fn, frames=c["frames"], full_mask=True)
    assert draws["n"] == int(g["renoise_draws"])
    assert np.array_equal(lat, g["final_latent"]) and np.array_equal(den, g["final_denoised"])


def test_vae_tiling_oracle_matches_reference():
    """oracle.tiled_encode / wrap_pad / crop_wrap_padded == the reference's tiled_vae_encode_tensor_simple
    and the inline wrap-padded decode of i2v_sphere_panorama_pipeline.py:974-991 (tests/golden/vae.npz)."""
    import torch.nn.functional as F
    v = np.load(os.path.join(GOLDEN, "vae.npz"))
    mix = torch.tensor([[0.5, -0.25, 0.125], [0.1, 0.2, 0.3], [-0.4, 0.3, 0.2], [0.05, 0.6, -0.35]])

    def enc(a):
        x = torch.from_numpy(np.ascontiguousarray(a))
        B, C, Fr, H, W = x.shape
        p = F.avg_pool2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), 8)
        y = torch.einsum("oc,nchw->nohw", mix, p) + 0.01 * p.mean(dim=(1, 2, 3), keepdim=True)
        return y.view(B, Fr, 4, H // 8, W // 8).permute(0, 2, 1, 3, 4).numpy()

    for n in range(3):
        shape = tuple(int(x) for x in v[f"encode/{n}/shape"])
        ht, wt, oh, ow = (int(x) for x in v[f"encode/{n}/kw"])
        got = po.tiled_encode(randn(600 + n, *shape), enc, h_tile_num=ht, w_tile_num=wt, overlap_h=oh, overlap_w=ow)
        assert np.array_equal(got, v[f"encode/{n}/latent"]), n
        shape = tuple(int(x) for x in v[f"decode/{n}/shape"])
        cn = int(v[f"decode/{n}/chunk_num"])
        padded = po.wrap_pad(randn(700 + n, *shape), cn)
        assert np.array_equal(padded, v[f"decode/{n}/padded"]), n
        up = np.repeat(np.repeat(padded, 2, axis=-1), 2, axis=-2) * np.arange(1, shape[2] + 1, dtype=np.float32)[None, None, :, None, None]
        assert np.array_equal(po.crop_wrap_padded(up, cn), v[f"decode/{n}/videos"]), n
