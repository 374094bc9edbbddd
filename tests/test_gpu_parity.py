"""CUDA path (through the C ABI) vs the oracle and the reference's golden vectors.
Everything here needs a B200: run with ``pytest -m gpu``."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import none_if_neg, randn
from oracle import panorama_oracle as po

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5          # BASELINE.json north_star: blended latents within 1e-5 max-abs in fp32
BF16_TOL = 1e-2          # ... 1e-2 in bf16


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dynamicscaler_b200 import lib
    from dynamicscaler_b200 import _lib
    assert lib.ds_abi_version() == _lib.ABI_VERSION          # the CUDA extension must be the thing that runs
    return torch.device("cuda:0")


def G(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


# ------------------------------------------------------------------ ring windows ----
def test_ring_get_set_golden(small, dev):
    from dynamicscaler_b200.shift_window_utils import RingLatent
    lat0 = randn(11, *small["ring/shape"])
    for i, c in enumerate(small["ring/get_cases"]):
        got = RingLatent(G(lat0, dev)).get_window_latent(*none_if_neg(c))
        assert got.is_contiguous() and np.array_equal(N(got), small[f"ring/get/{i}"]), i
    for i, c in enumerate(small["ring/set_cases"]):
        r = RingLatent(G(lat0, dev))
        shp = po.ring_get(lat0, *none_if_neg(c)).shape
        r.set_window_latent(G(randn(100 + i, *shp), dev), *none_if_neg(c))
        assert np.array_equal(N(r.torch_latent), small[f"ring/set/{i}"]), i


def test_ring_image_golden(small, dev):
    from dynamicscaler_b200.shift_window_utils import RingImageTensor
    img0 = randn(12, 3, 8, 12)
    for src in (torch.from_numpy(img0), G(img0, dev)):                # host image (uploaded once) or device image
        rit = RingImageTensor(None, src, height=8, width=12)
        for i, c in enumerate(small["img/cases"]):
            assert np.array_equal(N(rit.get_window_tensor(*none_if_neg(c))), small[f"img/get/{i}"])


@pytest.mark.parametrize("shape,win", [
    ((1, 4, 16, 64, 128), (12, 76, 3, 43, 0, 16)),        # config-1 window, vector path
    ((1, 4, 16, 64, 128), (13, 77, 60, 100, 9, 25)),      # unaligned left -> scalar loads, wraps on all axes
    ((2, 3, 5, 9, 18), (17, 30, 8, 15, 4, 9)),            # odd sizes -> scalar kernel
    ((1, 2, 3, 8, 16), (4, 32, 2, 16, 1, 6)),             # read longer than the axis (pixels repeat)
])
def test_ring_get_vs_oracle(dev, shape, win):
    from dynamicscaler_b200.shift_window_utils import RingLatent
    lat0 = randn(5, *shape)
    r = RingLatent(G(lat0, dev))
    assert np.array_equal(N(r.get_window_latent(*win)), po.ring_get(lat0, *win))
    F, H, W = shape[2:]
    l, rr, t, d, fb, fe = win
    if rr - l <= W and d - t <= H and fe - fb <= F:
        x = randn(6, shape[0], shape[1], fe - fb, d - t, rr - l)
        r.set_window_latent(G(x, dev), *win)
        ref = lat0.copy()
        po.ring_set(ref, x, *win)
        assert np.array_equal(N(r.torch_latent), ref)


def test_ring_write_asserts(dev):
    from dynamicscaler_b200.shift_window_utils import RingLatent
    r = RingLatent(torch.zeros(1, 1, 4, 8, 16, device=dev))
    with pytest.raises(AssertionError, match="warp should not occur"):
        r.set_window_latent(torch.zeros(1, 1, 4, 8, 17, device=dev), 0, 17)
    with pytest.raises(AssertionError, match="does not match target window shape"):
        r.set_window_latent(torch.zeros(1, 1, 4, 8, 5, device=dev), 0, 4)
    r.torch_latent = torch.ones(1, 1, 4, 16, 8, device=dev).permute(0, 1, 2, 4, 3)     # callers assign it
    assert N(r.get_window_latent(0, 16, 0, 8)).sum() == 4 * 8 * 16


# ---------------------------------------------------------------------- sphere ----
def test_sphere_proxy_golden(small, dev):
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy
    pano5 = randn(21, 1, 2, 3, 16, 32)
    for i, (fov, th, ph) in enumerate(small["sph/views"]):
        view_in = G(randn(300 + i, 1, 2, 3, 6, 10), dev)
        px = PanoramaLatentProxy(G(pano5, dev))
        vw, m = px.get_view_tensor_no_interpolate(fov, th, ph, 10, 6)
        assert np.array_equal(N(vw), small[f"sph/get_nearest/{i}"])
        assert np.array_equal(N(m), small[f"sph/get_nearest_mask/{i}"])
        got = px.get_view_tensor_interpolate(fov, th, ph, 10, 6)
        assert np.abs(N(got) - small[f"sph/get_bilinear/{i}"]).max() <= FP32_TOL
        assert np.array_equal(N(px.get_equirect_tensor()), pano5)               # reads do not mutate
        for name in ("set_view_tensor_no_interpolation", "set_view_tensor_bilinear", "set_view_tensor"):
            px = PanoramaLatentProxy(G(pano5, dev))
            getattr(px, name)(view_in, fov, th, ph)
            key = {"set_view_tensor_no_interpolation": "set_nearest", "set_view_tensor_bilinear": "set_bilinear",
                   "set_view_tensor": "set_round"}[name]
            want = small[f"sph/{key}/{i}"]
            got = N(px.get_equirect_tensor())
            assert got.shape == want.shape
            if key == "set_bilinear":
                assert np.abs(got - want).max() <= FP32_TOL, (name, i)
            else:
                assert np.array_equal(got, want), (name, i)


def test_sphere_bilinear_write_is_bit_exact(small, dev):
    """Stronger than the 1e-5 bar: the gather-form splat reproduces index_add_'s order."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy
    pano5 = randn(21, 1, 2, 3, 16, 32)
    for i, (fov, th, ph) in enumerate(small["sph/views"]):
        px = PanoramaLatentProxy(G(pano5, dev))
        px.set_view_tensor_bilinear(G(randn(300 + i, 1, 2, 3, 6, 10), dev), fov, th, ph)
        assert np.array_equal(N(px.get_equirect_tensor()), small[f"sph/set_bilinear/{i}"])


def test_panorama_tensor_image_and_mask_golden(small, dev):
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaTensor
    img = randn(22, 3, 16, 32)
    for i, (fov, th, ph) in enumerate(small["sph/views"][:3]):
        pt = PanoramaTensor(G(img, dev))
        assert np.array_equal(N(pt.get_view_tensor_no_interpolate(fov, th, ph, 10, 6)[0]), small[f"sph/img_get/{i}"])
        assert np.abs(N(pt.get_view_tensor_interpolate(fov, th, ph, 10, 6)) - small[f"sph/img_get_bilinear/{i}"]).max() <= FP32_TOL
        v = G(randn(400 + i, 3, 6, 10), dev)
        pt.set_view_tensor_no_interpolation(v, fov, th, ph)
        assert np.array_equal(N(pt.equirect_tensor), small[f"sph/img_set/{i}"])
        pt = PanoramaTensor(G(img, dev)); pt.set_view_tensor(v, fov, th, ph)
        assert np.array_equal(N(pt.equirect_tensor), small[f"sph/img_set_round/{i}"])
        pt = PanoramaTensor(G(img, dev)); pt.set_view_tensor_bilinear(v, fov, th, ph)
        assert np.array_equal(N(pt.equirect_tensor), small[f"sph/img_set_bilinear/{i}"])
        pm = PanoramaTensor(torch.zeros(1, 16, 32, device=dev))
        pm.set_view_tensor_no_interpolation(torch.ones(1, 6, 10, device=dev), fov, th, ph)
        assert np.array_equal(N(pm.equirect_tensor), small[f"sph/mask_set/{i}"])


def test_cpu_condition_image_is_uploaded_once(small, dev):
    """The I2V pipeline wraps the CPU condition image (torch.from_numpy, i2v_sphere_panorama_pipeline.py:
    223-224) in a PanoramaTensor and reads a view from it per window (:367): the image is uploaded once,
    views equal the reference's (computed from the same CPU tensor) bit for bit."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaTensor
    img = torch.from_numpy(randn(22, 3, 16, 32))                       # stays on the CPU, like the reference's
    pt = PanoramaTensor(img)
    stores = set()
    for i, (fov, th, ph) in enumerate(small["sph/views"][:3]):
        view, mask = pt.get_view_tensor_no_interpolate(fov, th, ph, 10, 6)
        assert view.is_cuda and np.array_equal(N(view), small[f"sph/img_get/{i}"])
        assert mask.shape == (6, 10) and bool((mask == 1).all())
        stores.add(pt.equirect_tensor.data_ptr())
    assert len(stores) == 1 and pt.equirect_tensor.is_cuda             # one upload, then resident


def test_proxy_single_frame_and_single_channel(dev):
    """N == 1 or C == 1 latents: the permuted [B,N,C,H,W] alias is then itself contiguous, so proxy layout
    must come from the explicit flag, not from strides.  Shapes and values vs the oracle."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy
    for (C, Fn) in ((4, 1), (1, 5), (1, 1)):
        lat0 = randn(70 + C + Fn, 1, C, Fn, 16, 32)
        px = PanoramaLatentProxy(G(lat0, dev))
        view, _ = px.get_view_tensor_no_interpolate(fov=100, theta=30, phi=20, width=10, height=6)
        ref, _ = po.proxy_get_nearest(lat0, 100, 30, 20, 10, 6)
        assert tuple(view.shape) == (1, C, Fn, 6, 10) and np.array_equal(N(view), ref), (C, Fn)
        new = randn(80 + C + Fn, 1, C, Fn, 6, 10)
        px.set_view_tensor_no_interpolation(G(new, dev), fov=100, theta=30, phi=20)
        want = lat0.copy()
        po.proxy_set_nearest(want, new, 100, 30, 20)
        out = px.get_equirect_tensor()
        assert tuple(out.shape) == (1, C, Fn, 16, 32) and np.array_equal(N(out), want), (C, Fn)
        got_b = px.get_view_tensor_interpolate(fov=100, theta=30, phi=20, width=10, height=6)
        assert tuple(got_b.shape) == (1, C, Fn, 6, 10)
        assert np.abs(N(got_b) - po.proxy_get_bilinear(want, 100, 30, 20, 10, 6)).max() <= FP32_TOL


def test_frame_ring_sphere_golden(small, dev):
    from dynamicscaler_b200.ring_panorama_tensor_utils import RingPanoramaLatentProxy
    rp5 = randn(31, 1, 2, 7, 16, 32)
    for i, (fov, th, ph, fb, fe) in enumerate(small["rsph/cases"]):
        fb, fe = int(fb), int(fe)
        view_in = G(randn(500 + i, 1, 2, fe - fb, 6, 10), dev)
        px = RingPanoramaLatentProxy(G(rp5, dev))
        got = px.get_view_tensor_no_interpolate(fov, th, ph, 10, 6, frame_begin=fb, frame_end=fe)[0]
        assert np.array_equal(N(got), small[f"rsph/get_nearest/{i}"])
        got = px.get_view_tensor_interpolate(fov, th, ph, 10, 6, frame_begin=fb, frame_end=fe)
        assert np.abs(N(got) - small[f"rsph/get_bilinear/{i}"]).max() <= FP32_TOL
        for name, key in (("set_view_tensor_no_interpolation", "set_nearest"),
                          ("set_view_tensor_bilinear", "set_bilinear"), ("set_view_tensor", "set_round")):
            if f"rsph/{key}/{i}" not in small:
                continue
            px = RingPanoramaLatentProxy(G(rp5, dev))
            getattr(px, name)(view_in, fov, th, ph, frame_begin=fb, frame_end=fe)
            assert np.array_equal(N(px.get_equirect_tensor()), small[f"rsph/{key}/{i}"]), (name, i)


def test_grid_sample_modes(dev):
    """interpolate_mode='nearest' / align_corners=False against torch's own grid_sample."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy
    import torch.nn.functional as Fn
    pano = randn(3, 1, 2, 2, 16, 32)
    u, v = po.uv_map(100, 30, 20, 10, 6, 16, 32)
    grid = torch.stack(((torch.from_numpy(u) / 31) * 2 - 1, (torch.from_numpy(v) / 15) * 2 - 1), -1)[None]
    px = PanoramaLatentProxy(G(pano, dev))
    for mode, ac in (("nearest", True), ("bilinear", False), ("nearest", False)):
        want = Fn.grid_sample(torch.from_numpy(pano).permute(0, 2, 1, 3, 4).reshape(2, 2, 16, 32),
                              grid.repeat(2, 1, 1, 1), mode=mode, padding_mode="border", align_corners=ac)
        want = want.view(1, 2, 2, 6, 10).permute(0, 2, 1, 3, 4).numpy()
        got = N(px.get_view_tensor_interpolate(100, 30, 20, 10, 6, interpolate_mode=mode, interpolate_align_corners=ac))
        assert np.abs(got - want).max() <= FP32_TOL, (mode, ac)
    with pytest.raises(NotImplementedError):
        px.get_view_tensor_interpolate(100, 30, 20, 10, 6, interpolate_mode="bicubic")


# --------------------------------------------------------------------- mask mix ----
def test_mix_mask_golden(small, dev):
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    l1, add = randn(41, 1, 4, 5, 6, 10), randn(42, 1, 4, 5, 6, 10)
    m3 = (randn(43, 1, 6, 10) > 0).astype(np.float32)
    m5 = (randn(44, 1, 4, 5, 6, 10) > 0).astype(np.float32)
    for i, (m, r) in enumerate([(m3, 0.5), (m3, 0.3), (m5, 0.8), (m5, 1.0), (m3, 0.0)]):
        got = mix_latents_with_mask(G(l1, dev), G(add, dev), G(m, dev), r)
        assert np.array_equal(N(got), small[f"mix/{i}"]), i
    # non-contiguous inputs (the proxies hand out permuted views) and soft masks
    soft = np.abs(randn(45, 1, 7, 9)) % 1
    a, b = randn(46, 2, 5, 3, 7, 9), randn(47, 2, 5, 3, 7, 9)
    got = mix_latents_with_mask(G(a.transpose(0, 2, 1, 3, 4), dev).permute(0, 2, 1, 3, 4), G(b, dev), G(soft, dev), 0.37)
    assert np.array_equal(N(got), po.mix_mask(a, b, soft, 0.37))


# ----------------------------------------------------------- step compositions ----
def _payload(n, win):
    return (win * np.float32(0.5) + randn(700 + n, *win.shape), win * np.float32(0.25) - randn(800 + n, *win.shape))


def test_sequential_step_dropin_golden(small, dev):
    """The reference's own window-after-window loop (t2v_sphere_panorama_pipeline.py:176-291)
    with the drop-in classes: regime (i), unmodified call sites."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy, PanoramaTensor
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    lat = G(randn(51, 1, 2, 3, 16, 32), dev)
    h_lat = PanoramaLatentProxy(lat)
    h_den = PanoramaLatentProxy(torch.zeros_like(lat))
    h_msk = PanoramaTensor(torch.zeros_like(lat[0, 0, [0]]))
    for n, (ph, th) in enumerate(small["step/views"]):
        vw, _ = h_lat.get_view_tensor_no_interpolate(fov=120, theta=th, phi=ph, width=10, height=6)
        mk, _ = h_msk.get_view_tensor_no_interpolate(fov=120, theta=th, phi=ph, width=10, height=6)
        noised = vw * 0.9 + G(randn(600 + n, 1, 2, 3, 6, 10), dev) * 0.4359
        vw = mix_latents_with_mask(latent_1=vw, latent_to_add=noised, mask=mk, mix_ratio=0.6)
        x_prev = vw * 0.5 + G(randn(700 + n, 1, 2, 3, 6, 10), dev)
        x0 = vw * 0.25 - G(randn(800 + n, 1, 2, 3, 6, 10), dev)
        h_lat.set_view_tensor_no_interpolation(view_tensor=x_prev, fov=120, theta=th, phi=ph)
        h_den.set_view_tensor_no_interpolation(view_tensor=x0, fov=120, theta=th, phi=ph)
        h_msk.set_view_tensor_no_interpolation(view_tensor=torch.ones_like(x_prev[0, 0, [0]]), fov=120, theta=th, phi=ph)
    # torch's GPU elementwise ops may contract a*b+c differently from the CPU ones used for the
    # golden run, so the payloads differ in the last ulp; the tiling path itself adds no error.
    assert np.abs(N(h_lat.get_equirect_tensor()) - small["step/seq/latent"]).max() <= FP32_TOL
    assert np.abs(N(h_den.get_equirect_tensor()) - small["step/seq/denoised"]).max() <= FP32_TOL
    assert np.array_equal(N(h_msk.equirect_tensor), small["step/seq/mask"])


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_batched_step_small_golden(small, dev, mode):
    from dynamicscaler_b200.step import PanoramaStepper
    views = [(ph, th) for ph, th in small["step/views"]]
    st = PanoramaStepper(16, 32, 3, device=dev)
    plan = st.plan_sphere(views, 10, 6, fov=120, mode=mode)
    lat = G(randn(51, 1, 2, 3, 16, 32), dev)
    den = torch.zeros_like(lat)
    mask = torch.zeros(1, 16, 32, device=dev)
    wins = st.gather_all(lat, plan)
    tol = 0 if mode == "nearest" else FP32_TOL
    assert np.abs(N(wins) - small[f"step/batched_{mode}/windows"]).max() <= tol
    gold = small[f"step/batched_{mode}/windows"]
    pays = [_payload(n, gold[n]) for n in range(len(views))]
    st.merge_all([lat, den], [G(np.stack([p[0] for p in pays]), dev), G(np.stack([p[1] for p in pays]), dev)], plan, mask=mask)
    assert np.abs(N(lat) - small[f"step/batched_{mode}/latent"]).max() <= tol
    assert np.abs(N(den) - small[f"step/batched_{mode}/denoised"]).max() <= tol
    assert np.array_equal(N(mask), small[f"step/batched_{mode}/mask"])


def test_config1_ring_step_digest(configs, dev):
    """BASELINE.json config 1 at full size, bit-exact against the reference's digests."""
    from dynamicscaler_b200.step import PanoramaStepper
    lat = G(randn(0, 1, 4, 16, 64, 128), dev)
    den, msk = torch.zeros_like(lat), torch.zeros_like(lat)
    windows = [(int(l), int(t)) for l, t in configs["cfg1/windows"]]
    st = PanoramaStepper(64, 128, 16, device=dev)
    plan = st.plan_ring(windows, 64, 40)
    assert plan.vec == 4 and plan.coverage == 1.0
    wins = st.gather_all(lat, plan)
    assert np.array_equal(sha(N(wins)), configs["cfg1/windows_sha"])
    x_prev = G(np.stack([randn(1000 + n, 1, 4, 16, 40, 64) for n in range(128)]), dev)
    x0 = G(np.stack([randn(2000 + n, 1, 4, 16, 40, 64) for n in range(128)]), dev)
    st.merge_all([lat, den], [x_prev, x0], plan, mask=msk)
    assert np.array_equal(sha(N(lat)), configs["cfg1/latent_sha"])
    assert np.array_equal(sha(N(den)), configs["cfg1/denoised_sha"])
    assert np.array_equal(sha(N(msk)), configs["cfg1/mask_sha"])


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_config2_sphere_step_digest(configs, dev, mode):
    """BASELINE.json config 2 (the bench workload) at full size against the reference."""
    from dynamicscaler_b200.step import PanoramaStepper
    views = [(ph, th) for ph, th in configs["cfg2/views"]]
    st = PanoramaStepper(128, 256, 16, device=dev)
    plan = st.plan_sphere(views, 64, 40, fov=120, mode=mode)
    lat = G(randn(0, 1, 4, 16, 128, 256), dev)
    den = torch.zeros_like(lat)
    msk = torch.zeros(1, 128, 256, device=dev)
    wins = st.gather_all(lat, plan)
    if mode == "nearest":
        assert np.array_equal(sha(N(wins)), configs["cfg2/nearest/windows_sha"])
    else:
        assert np.abs(N(wins).reshape(-1)[::4099] - configs["cfg2/bilinear/windows_sample"]).max() <= FP32_TOL
    x_prev = G(np.stack([randn(1000 + n, 1, 4, 16, 40, 64) for n in range(44)]), dev)
    x0 = G(np.stack([randn(2000 + n, 1, 4, 16, 40, 64) for n in range(44)]), dev)
    st.merge_all([lat, den], [x_prev, x0], plan, mask=msk)
    for name, t in (("latent", lat), ("denoised", den), ("mask", msk)):
        a = N(t)
        step = 4099 if a.size > 100000 else 7
        assert np.abs(a.reshape(-1)[::step] - configs[f"cfg2/{mode}/{name}_sample"]).max() <= FP32_TOL, name
        assert np.array_equal(sha(a), configs[f"cfg2/{mode}/{name}_sha"]), name     # in fact bit-exact


def test_config4_frame_ring_digest(configs, dev):
    from dynamicscaler_b200.ring_panorama_tensor_utils import RingPanoramaLatentProxy
    lat = G(randn(4, 1, 4, 64, 128, 256), dev)
    jobs = [(int(fb), ph, th) for fb, ph, th in configs["cfg4/jobs"]]
    h_lat = RingPanoramaLatentProxy(lat)
    wins = [h_lat.get_view_tensor_no_interpolate(120, th, ph, 64, 40, frame_begin=fb, frame_end=fb + 16)[0].clone()
            for fb, ph, th in jobs]
    for n, (fb, ph, th) in enumerate(jobs):
        h_lat.set_view_tensor_no_interpolation(G(randn(1000 + n, 1, 4, 16, 40, 64), dev), 120, th, ph,
                                               frame_begin=fb, frame_end=fb + 16)
    assert np.array_equal(sha(N(torch.stack(wins))), configs["cfg4/windows_sha"])
    assert np.array_equal(sha(N(h_lat.get_equirect_tensor())), configs["cfg4/latent_sha"])


def test_config4_batched_frame_windows_vs_oracle(dev):
    """13 wrapping frame windows x 6 views (temporal loop), batched, vs the oracle composition."""
    from dynamicscaler_b200.step import PanoramaStepper
    F, H, W = 64, 32, 64
    views = po.default_sphere_views(2)[:6]
    fbs = po.frame_window_begins(3)
    lat0 = randn(7, 1, 4, F, H, W)
    st = PanoramaStepper(H, W, F, device=dev)
    plan = st.plan_sphere(views, 16, 10, fov=120, mode="nearest", frame_begins=fbs, frames=16)
    lat = G(lat0, dev)
    den, msk = torch.zeros_like(lat), torch.zeros_like(lat)
    wins = st.gather_all(lat, plan)
    pays = {}

    def pay(n, win):
        pays[n] = (randn(3000 + n, *win.shape), randn(4000 + n, *win.shape))
        return pays[n]

    ref_lat, ref_den = lat0.copy(), np.zeros_like(lat0)
    ref_wins, ref_mask = po.sphere_step_batched(ref_lat, ref_den, views, pay, fov=120, width=16, height=10,
                                                mode="nearest", frame_begins=fbs, frames=16)
    assert np.array_equal(N(wins), np.stack(ref_wins))
    st.merge_all([lat, den], [G(np.stack([pays[n][0] for n in range(plan.n_jobs)]), dev),
                              G(np.stack([pays[n][1] for n in range(plan.n_jobs)]), dev)], plan, mask=msk)
    assert np.array_equal(N(lat), ref_lat) and np.array_equal(N(den), ref_den)
    assert np.array_equal(N(msk), np.broadcast_to(ref_mask, lat0.shape))


# ------------------------------------------------ full-size, size-independent checks ----
def test_config5_ring_properties(dev):
    """4K-scale panorama [1,4,16,512,1024], 16x16 windows (BASELINE config 5): round trip,
    idempotence and coverage -- properties that need no CPU-side recomputation."""
    from dynamicscaler_b200.step import PanoramaStepper
    H, W, F = 512, 1024, 16
    windows = po.ring_window_grid(5, 16, 16, 8192, 4096)
    st = PanoramaStepper(H, W, F, device=dev)
    plan = st.plan_ring(windows, 64, 40)
    assert plan.n_jobs == 256 and plan.vec == 4 and plan.coverage == 1.0
    gen = torch.Generator(device=dev).manual_seed(0)
    lat = torch.randn(1, 4, F, H, W, device=dev, generator=gen)
    before = lat.clone()
    wins = st.gather_all(lat, plan)
    # merge(gather(x)) == x : every pixel is covered and receives its own value back
    st.merge_all([lat], [wins], plan)
    assert torch.equal(lat, before)
    # spot-check windows against plain modular indexing done by torch on the GPU
    for j in (0, 17, 255):
        l, t = windows[j]
        ys = (t + torch.arange(40, device=dev)) % H
        xs = (l + torch.arange(64, device=dev)) % W
        assert torch.equal(wins[j], before[:, :, :, ys][:, :, :, :, xs])
    # last writer wins: merging window-index payloads leaves, per pixel, the highest covering index
    pay = torch.arange(256, device=dev, dtype=torch.float32).view(256, 1, 1, 1, 1, 1).expand(256, 1, 4, F, 40, 64).contiguous()
    st.merge_all([lat], [pay], plan)
    owner = torch.full((H, W), -1.0, device=dev)
    for j, (l, t) in enumerate(windows):
        ys = (t + torch.arange(40, device=dev)) % H
        xs = (l + torch.arange(64, device=dev)) % W
        owner[ys[:, None], xs[None, :]] = float(j)
    assert torch.equal(lat[0, 0, 0], owner) and torch.equal(lat[0, 3, 15], owner)
    # idempotence
    again = lat.clone()
    st.merge_all([lat], [pay], plan)
    assert torch.equal(lat, again)


def test_sphere_full_size_vs_oracle_all_offsets(dev):
    """[1,4,16,256,512] (north-star shape), nearest, three longitude offsets vs the oracle."""
    from dynamicscaler_b200.step import PanoramaStepper
    H, W, F = 256, 512, 16
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = randn(9, 1, 4, F, H, W)
    for i_step in (0, 4, 9):
        views = po.default_sphere_views(i_step)
        plan = st.plan_sphere(views, 64, 40, fov=120, mode="nearest")
        lat = G(lat0, dev)
        wins = st.gather_all(lat, plan)
        x_prev = wins * 2
        st.merge_all([lat], [x_prev], plan)
        ref = lat0.copy()
        ref_wins, _ = po.sphere_step_batched(ref, np.zeros_like(ref), views, lambda n, w: (w * np.float32(2), w), mode="nearest")
        assert np.array_equal(N(wins), np.stack(ref_wins)), i_step
        assert np.array_equal(N(lat), ref), i_step
        assert 0.5 < plan.coverage < 0.6


# ---------------------------------------------------------------- bf16 / fp16 ----
@pytest.mark.parametrize("half_dtype", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_bf16_within_tolerance(dev, half_dtype):
    """bf16 / fp16 latents (the reference's use_fp16 mode, gen_pano_360.py:191-213), fp32 maps: vs the fp32
    oracle on the rounded inputs at 1e-2 (nearest paths are pure copies and stay bit-exact)."""
    from dynamicscaler_b200.step import PanoramaStepper
    from dynamicscaler_b200.tensor_utils import mix_latents_with_mask
    H, W, F = 32, 64, 4
    views = po.default_sphere_views(1)[:10]
    lat_bf = G(randn(13, 1, 4, F, H, W), dev).to(half_dtype)
    lat0 = lat_bf.float().cpu().numpy()
    st = PanoramaStepper(H, W, F, device=dev)
    for mode in ("nearest", "bilinear"):
        plan = st.plan_sphere(views, 16, 10, fov=120, mode=mode)
        lat = lat_bf.clone()
        wins = st.gather_all(lat, plan)
        pays = G(np.stack([randn(50 + n, 1, 4, F, 10, 16) for n in range(10)]), dev).to(half_dtype)
        st.merge_all([lat], [pays], plan)
        ref = lat0.copy()
        pay_np = pays.float().cpu().numpy()
        ref_wins, _ = po.sphere_step_batched(ref, np.zeros_like(ref), views, lambda n, w: (pay_np[n], pay_np[n]),
                                             fov=120, width=16, height=10, mode=mode)
        assert np.abs(wins.float().cpu().numpy() - np.stack(ref_wins)).max() <= (0 if mode == "nearest" else BF16_TOL * 4)
        assert np.abs(lat.float().cpu().numpy() - ref).max() <= (0 if mode == "nearest" else BF16_TOL * 4)
    a, b = lat_bf, (lat_bf * 0.5)
    m = (torch.rand(1, H, W, device=dev) > 0.5).to(half_dtype)
    got = mix_latents_with_mask(a, b, m, 0.25).float().cpu().numpy()
    want = po.mix_mask(a.float().cpu().numpy(), b.float().cpu().numpy(), m.float().cpu().numpy(), 0.25)
    assert np.abs(got - want).max() <= BF16_TOL * 4


# ----------------------------------------------------------- raw C-ABI calls ----
def test_raw_cabi_ring_gather_with_device_descriptors(dev):
    """Straight through the C ABI with plain pointers, as a non-Python host would."""
    from dynamicscaler_b200._lib import lib
    lat0 = randn(21, 1, 3, 4, 16, 32)
    lat = G(lat0, dev)
    wins = torch.tensor([[4, 2, 0], [28, 12, 3], [0, 0, 1]], dtype=torch.int32, device=dev)
    out = torch.empty(3, 3, 4, 8, 12, device=dev)
    rc = lib.ds_ring_gather(lat.data_ptr(), out.data_ptr(), wins.data_ptr(), 3, 0, 0, 0, 3, 4, 16, 32, 4, 8, 12, 0,
                            torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    for j, (l, t, f0) in enumerate(wins.tolist()):
        assert np.array_equal(N(out[j]), po.ring_get(lat0, l, l + 12, t, t + 8, f0, f0 + 4)[0])
    assert lib.ds_ring_gather(lat.data_ptr(), out.data_ptr(), None, 1, 70, 0, 0, 3, 4, 16, 32, 4, 8, 12, 0, None) == -4


# ------------------------------------------------------------ resize_video_latent ----
def test_resize_video_latent_golden(dev):
    """ds_resize_planes vs the reference's outputs (tests/golden/resize.npz): nearest bit-exact,
    bilinear / bicubic within 1e-5; plus the oracle on a stage-transition-sized latent."""
    import os
    from dynamicscaler_b200.diffusion_utils import padding_latents_at_end, padding_latents_at_front, resize_video_latent
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resize.npz"))
    modes = ["nearest", "bilinear", "bicubic"]
    for n, row in enumerate(g["cases"]):
        seed, shape, th, tw = int(row[0]), tuple(int(v) for v in row[1:6]), int(row[6]), int(row[7])
        mode, align = modes[row[8]], bool(row[9])
        got = N(resize_video_latent(G(randn(seed, *shape), dev), th, tw, mode=mode, align_corners=align))
        want = g[f"out/{n}"]
        assert got.shape == want.shape
        if mode == "nearest":
            assert np.array_equal(got, want), n
        else:
            assert np.abs(got - want).max() <= FP32_TOL, (n, np.abs(got - want).max())
    # stage transitions of gen_pano_360.py at full size: 128x256 -> 64x128 nearest, 64x128 -> 128x256 bicubic
    x = randn(77, 1, 4, 16, 128, 256)
    assert np.array_equal(N(resize_video_latent(G(x, dev), 64, 128, mode="nearest")),
                          po.resize_video_latent(x, 64, 128, mode="nearest"))
    y = randn(78, 1, 4, 16, 64, 128)
    assert np.abs(N(resize_video_latent(G(y, dev), 128, 256, mode="bicubic")) -
                  po.resize_video_latent(y, 128, 256, mode="bicubic")).max() <= FP32_TOL
    # odd widths (scalar tail), bf16
    z = G(randn(79, 1, 4, 2, 7, 9), dev).to(torch.bfloat16)
    ref = po.resize_video_latent(z.float().cpu().numpy(), 15, 19, mode="bilinear")
    assert np.abs(resize_video_latent(z, 15, 19).float().cpu().numpy() - ref).max() <= BF16_TOL * 4
    with pytest.raises(RuntimeError):
        resize_video_latent(torch.zeros(2, 4, 3, 8, 8, device=dev), 4, 4)       # the reference's .view fails too
    t = G(randn(80, 1, 4, 3, 4, 4), dev)
    assert padding_latents_at_front(t, 2).shape[2] == 5 and torch.equal(padding_latents_at_front(t, 2)[:, :, 1], t[:, :, 0])
    assert padding_latents_at_end(t, 1).shape[2] == 4 and torch.equal(padding_latents_at_end(t, 1)[:, :, 3], t[:, :, 2])


def test_batched_condition_image_windows(dev):
    """§8f-3: all condition-image windows / views of a step in one launch each, vs the per-call
    drop-in classes (themselves pinned to the reference's golden vectors above) and the oracle."""
    from dynamicscaler_b200.panorama_tensor_utils import PanoramaTensor
    from dynamicscaler_b200.shift_window_utils import RingImageTensor
    from dynamicscaler_b200.step import PanoramaStepper
    img0 = randn(91, 3, 64, 128)
    img = G(img0, dev)
    st = PanoramaStepper(8, 16, 4, device=dev)
    windows = [(0, 0), (100, 40), (127, 63), (64, 10)]                      # wrapping on both axes
    got = st.gather_image_windows(img, windows, 40, 32)
    ring = RingImageTensor(image_path=None, image_tensor=img, height=64, width=128)
    for n, (l, t) in enumerate(windows):
        assert torch.equal(got[n], ring.get_window_tensor(l, l + 40, t, t + 32))
        assert np.array_equal(N(got[n]), po.ring_get(img0[None, :, None], l, l + 40, t, t + 32, 0, 1)[0, :, 0])
    with pytest.raises(AssertionError):
        st.gather_image_windows(img, [(200, 0)], 80, 32)
    views = po.default_sphere_views(2)[:9]
    gv = st.gather_image_views(img, views, 20, 12, fov=120)
    pt = PanoramaTensor(img)
    for n, (ph, th) in enumerate(views):
        assert torch.equal(gv[n], pt.get_view_tensor_no_interpolate(120, th, ph, 20, 12)[0])
