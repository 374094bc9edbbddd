"""Pixel-major resident path (csrc/pm.cu, PanoramaState) vs the oracle, the reference's
golden digests and the plane-major kernels.  Needs a B200: ``pytest -m gpu``."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import randn
from oracle import panorama_oracle as po

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 1e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dynamicscaler_b200 import lib
    assert lib.ds_abi_version() == 3
    return torch.device("cuda:0")


def G(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().float().cpu().numpy() if t.dtype == torch.bfloat16 else t.detach().cpu().numpy()


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("F,H,W", [(16, 128, 256), (3, 6, 12), (5, 10, 20), (64, 16, 32)])
def test_pack_unpack_round_trip(dev, dtype, F, H, W):
    """pack = ds_pm_merge_nearest(identity plan), unpack = ds_pm_gather(identity map); ragged
    tile tails (H*W not a multiple of 64/32) and frame counts that are not multiples of 8."""
    from dynamicscaler_b200.step import PanoramaStepper
    if (H * W) % (16 // torch.empty(0, dtype=dtype).element_size()) != 0:
        pytest.skip("unpack needs H*W to be a multiple of the 16-byte vector")
    st = PanoramaStepper(H, W, F, device=dev)
    lat = G(randn(3, 1, 4, F, H, W), dev).to(dtype)
    state = st.make_state([lat])
    pm = state.pms[0]
    assert pm.shape == (H * W, F, 4)
    # pm[p][f][pl] == lat[0][pl][f][p]
    want = lat[0].permute(2, 3, 1, 0).reshape(H * W, F, 4)
    assert torch.equal(pm, want)
    assert torch.equal(state.export(0), lat)


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_config2_state_step_digest(configs, dev, mode):
    """BASELINE.json config 2 (the bench workload) through the resident pixel-major state,
    against the reference's digests: bit-exact for both modes."""
    from dynamicscaler_b200.step import PanoramaStepper
    views = [(ph, th) for ph, th in configs["cfg2/views"]]
    st = PanoramaStepper(128, 256, 16, device=dev)
    plan = st.plan_sphere(views, 64, 40, fov=120, mode=mode)
    lat = G(randn(0, 1, 4, 16, 128, 256), dev)
    state = st.make_state([lat, torch.zeros_like(lat)], mask=torch.zeros(1, 128, 256, device=dev))
    wins = st.gather_all(state, plan)
    if mode == "nearest":
        assert np.array_equal(sha(N(wins)), configs["cfg2/nearest/windows_sha"])
    else:
        assert np.abs(N(wins).reshape(-1)[::4099] - configs["cfg2/bilinear/windows_sample"]).max() <= FP32_TOL
    # identical to the plane-major kernel, bit for bit
    assert torch.equal(wins, st.gather_all(lat, plan))
    x_prev = G(np.stack([randn(1000 + n, 1, 4, 16, 40, 64) for n in range(44)]), dev)
    x0 = G(np.stack([randn(2000 + n, 1, 4, 16, 40, 64) for n in range(44)]), dev)
    st.merge_all(state, [x_prev, x0], plan)
    out_lat, out_den = state.export()
    for name, t in (("latent", out_lat), ("denoised", out_den), ("mask", state.export_mask())):
        assert np.array_equal(sha(N(t)), configs[f"cfg2/{mode}/{name}_sha"]), name


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_state_frame_windows_vs_oracle(dev, mode):
    """Wrapping frame windows (temporal loop) x views on the resident state, latent-shaped
    (I2V) mask, vs the oracle composition."""
    from dynamicscaler_b200.step import PanoramaStepper
    F, H, W = 64, 32, 64
    views = po.default_sphere_views(2)[:6]
    fbs = po.frame_window_begins(3)
    lat0 = randn(7, 1, 4, F, H, W)
    st = PanoramaStepper(H, W, F, device=dev)
    plan = st.plan_sphere(views, 16, 10, fov=120, mode=mode, frame_begins=fbs, frames=16)
    lat = G(lat0, dev)
    state = st.make_state([lat, torch.zeros_like(lat)], mask=torch.zeros_like(lat))
    wins = st.gather_all(state, plan)
    pays = {}

    def pay(n, win):
        pays[n] = (randn(3000 + n, *win.shape), randn(4000 + n, *win.shape))
        return pays[n]

    ref_lat, ref_den = lat0.copy(), np.zeros_like(lat0)
    ref_wins, ref_mask = po.sphere_step_batched(ref_lat, ref_den, views, pay, fov=120, width=16, height=10,
                                                mode=mode, frame_begins=fbs, frames=16)
    tol = 0 if mode == "nearest" else FP32_TOL
    assert np.abs(N(wins) - np.stack(ref_wins)).max() <= tol
    st.merge_all(state, [G(np.stack([pays[n][0] for n in range(plan.n_jobs)]), dev),
                         G(np.stack([pays[n][1] for n in range(plan.n_jobs)]), dev)], plan)
    out_lat, out_den = state.export()
    assert np.abs(N(out_lat) - ref_lat).max() <= tol * 10 and np.abs(N(out_den) - ref_den).max() <= tol * 10
    assert np.array_equal(N(state.export_mask()), np.broadcast_to(ref_mask, lat0.shape))


def test_state_full_size_nearest_vs_oracle(dev):
    """[1,4,16,256,512] (north-star shape), nearest, two longitude offsets, vs the oracle."""
    from dynamicscaler_b200.step import PanoramaStepper
    H, W, F = 256, 512, 16
    st = PanoramaStepper(H, W, F, device=dev)
    lat0 = randn(9, 1, 4, F, H, W)
    for i_step in (0, 7):
        views = po.default_sphere_views(i_step)
        plan = st.plan_sphere(views, 64, 40, fov=120, mode="nearest")
        state = st.make_state([G(lat0, dev)])
        wins = st.gather_all(state, plan)
        st.merge_all(state, [wins * 2], plan)
        ref = lat0.copy()
        ref_wins, _ = po.sphere_step_batched(ref, np.zeros_like(ref), views, lambda n, w: (w * np.float32(2), w), mode="nearest")
        assert np.array_equal(N(wins), np.stack(ref_wins)), i_step
        assert np.array_equal(N(state.export(0)), ref), i_step


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_state_equals_plane_major_ragged(dev, mode):
    """Odd geometry (hw = 12*7... not a multiple of 64, F = 5): the state path and the
    plane-major kernels of sphere.cu must agree bit for bit, 3 consecutive steps."""
    from dynamicscaler_b200.step import PanoramaStepper
    H, W, F, h, w = 24, 48, 5, 9, 12
    st = PanoramaStepper(H, W, F, device=dev)
    lat = G(randn(31, 1, 4, F, H, W), dev)
    den = torch.zeros_like(lat)
    msk = torch.zeros(1, H, W, device=dev)
    state = st.make_state([lat, den], mask=msk)
    for i_step in range(3):
        views = po.default_sphere_views(i_step)[:20]
        plan = st.plan_sphere(views, w, h, fov=100, mode=mode)
        a = st.gather_all(lat, plan)
        b = st.gather_all(state, plan)
        assert torch.equal(a, b)
        xp, x0 = a * 0.5 + 1, a * 0.25 - 1
        st.merge_all([lat, den], [xp, x0], plan, mask=msk)
        st.merge_all(state, [xp, x0], plan)
        e_lat, e_den = state.export()
        assert torch.equal(e_lat, lat) and torch.equal(e_den, den)
        assert torch.equal(state.export_mask(), msk)


def test_state_bf16_within_tolerance(dev):
    from dynamicscaler_b200.step import PanoramaStepper
    H, W, F = 32, 64, 4
    views = po.default_sphere_views(1)[:10]
    lat_bf = G(randn(13, 1, 4, F, H, W), dev).to(torch.bfloat16)
    lat0 = lat_bf.float().cpu().numpy()
    st = PanoramaStepper(H, W, F, device=dev)
    for mode in ("nearest", "bilinear"):
        plan = st.plan_sphere(views, 16, 10, fov=120, mode=mode)
        state = st.make_state([lat_bf])
        wins = st.gather_all(state, plan)
        pays = G(np.stack([randn(50 + n, 1, 4, F, 10, 16) for n in range(10)]), dev).to(torch.bfloat16)
        st.merge_all(state, [pays], plan)
        ref = lat0.copy()
        pay_np = pays.float().cpu().numpy()
        ref_wins, _ = po.sphere_step_batched(ref, np.zeros_like(ref), views, lambda n, w: (pay_np[n], pay_np[n]),
                                             fov=120, width=16, height=10, mode=mode)
        tol = 0 if mode == "nearest" else BF16_TOL * 4
        assert np.abs(N(wins) - np.stack(ref_wins)).max() <= tol
        assert np.abs(N(state.export(0)) - ref).max() <= tol


def test_state_rejects_other_plane_counts(dev):
    from dynamicscaler_b200.step import PanoramaStepper
    st = PanoramaStepper(16, 32, 3, device=dev)
    with pytest.raises(AssertionError):
        st.make_state([torch.zeros(1, 2, 3, 16, 32, device=dev)])
