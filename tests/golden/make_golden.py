"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, CPU, fp32) on seeded inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed here as
fixtures; inputs are regenerated from the recorded seeds with
``numpy.random.default_rng(seed).standard_normal(dtype=float32)``.
Large outputs are stored as sha256 digests plus a strided sample.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("DS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
_pl = types.ModuleType("pytorch_lightning")
_pl.LightningModule = torch.nn.Module
_pl.seed_everything = lambda *a, **k: None
_plu = types.ModuleType("pytorch_lightning.utilities")
_plu.rank_zero_only = lambda f: f
sys.modules.setdefault("pytorch_lightning", _pl)
sys.modules.setdefault("pytorch_lightning.utilities", _plu)

from utils.shift_window_utils import RingLatent, RingImageTensor, get_dimension_slices_and_sizes  # noqa: E402
from utils.panorama_tensor_utils import PanoramaTensor, PanoramaLatentProxy  # noqa: E402
from utils.ring_panorama_tensor_utils import RingPanoramaLatentProxy  # noqa: E402
from utils.tensor_utils import mix_latents_with_mask  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import panorama_oracle as po  # noqa: E402  (geometry helpers only)


def randn(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def N(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


out = {}

# ---- 1. ring slices ------------------------------------------------------
SLICE_CASES = [(0, 8, 8), (3, 8, 8), (3, 11, 8), (7, 16, 8), (8, 16, 8), (5, 6, 8), (0, 16, 8),
               (120, 256, 128), (60, 124, 128), (100, 164, 128), (15, 31, 16), (63, 79, 64)]
out["slices/cases"] = np.array(SLICE_CASES, dtype=np.int64)
for i, (b, e, s) in enumerate(SLICE_CASES):
    sl, sz = get_dimension_slices_and_sizes(b, e, s)
    assert [x.stop - x.start for x in sl] == sz
    out[f"slices/{i}"] = np.concatenate([np.arange(x.start, x.stop) for x in sl]).astype(np.int64)

# ---- 2. RingLatent get / set ---------------------------------------------
RING_SHAPE = (2, 3, 5, 8, 12)
RING_GET = [  # left,right,top,down,fb,fe  (None = default)
    (0, 12, 0, 8, 0, 5), (2, 9, 1, 6, 1, 4), (8, 16, 5, 11, 3, 7), (11, 23, 7, 15, 4, 9),
    (None, None, None, None, 2, 7), (4, 8, None, None, None, None), (5, 24, 3, 16, 0, 10),
    (0, 24, 0, 16, 0, 10), (9, 10, 7, 8, 4, 5)]
RING_SET = [(0, 12, 0, 8, 0, 5), (2, 9, 1, 6, 1, 4), (8, 16, 5, 11, 3, 7), (11, 23, 7, 15, 4, 9),
            (None, None, None, None, 2, 7), (4, 8, None, None, None, None), (9, 10, 7, 8, 4, 5)]
enc = lambda cases: np.array([[-1 if v is None else v for v in c] for c in cases], dtype=np.int64)
out["ring/shape"] = np.array(RING_SHAPE)
out["ring/get_cases"] = enc(RING_GET)
out["ring/set_cases"] = enc(RING_SET)
lat0 = randn(11, *RING_SHAPE)
for i, c in enumerate(RING_GET):
    rl = RingLatent(T(lat0))
    out[f"ring/get/{i}"] = N(rl.get_window_latent(*c))
for i, c in enumerate(RING_SET):
    rl = RingLatent(T(lat0))
    shp = rl.get_window_latent(*c).shape
    rl.set_window_latent(T(randn(100 + i, *shp)), *c)
    out[f"ring/set/{i}"] = N(rl.torch_latent)

# ---- 3. RingImageTensor ---------------------------------------------------
img0 = randn(12, 3, 8, 12)
IMG_GET = [(0, 12, None, None), (5, 11, 2, 7), (9, 17, 6, 12), (11, 24, 0, 16)]
out["img/cases"] = enc(IMG_GET)
for i, c in enumerate(IMG_GET):
    rit = RingImageTensor(image_path=None, image_tensor=T(img0), height=8, width=12)
    out[f"img/get/{i}"] = N(rit.get_window_tensor(*c))

# ---- 4. uv maps -----------------------------------------------------------
UV_CASES = [  # fov, theta, phi, w, h, H, W
    (120, 0, 90, 64, 40, 128, 256), (120, 0, -90, 64, 40, 128, 256), (120, 72, 75, 64, 40, 128, 256),
    (120, 408, -45, 64, 40, 128, 256), (120, 300, 0, 64, 40, 128, 256), (120, 60, 60, 64, 40, 64, 128),
    (120, 132, -60, 64, 40, 256, 512), (90, 45, 30, 10, 6, 16, 32), (120, 350, -80, 10, 6, 16, 32),
    (100, 12.5, 10, 12, 8, 16, 32), (120, 24, 45, 512, 320, 1024, 2048)]
out["uv/cases"] = np.array(UV_CASES, dtype=np.float64)
for i, (fov, th, ph, w, h, H, W) in enumerate(UV_CASES):
    pt = PanoramaTensor(torch.zeros(1, H, W))
    u, v = pt._get_uv(fov=fov, theta=th, phi=ph, width=w, height=h)
    if w * h <= 4096:
        out[f"uv/{i}/u"], out[f"uv/{i}/v"] = N(u), N(v)
    else:  # big image-cond map: digest + floor indices digest
        out[f"uv/{i}/u_sha"], out[f"uv/{i}/v_sha"] = digest(N(u)), digest(N(v))

# ---- 5. sphere ops, small geometry ----------------------------------------
SM = dict(B=1, C=2, F=3, H=16, W=32, h=6, w=10)
SVIEWS = [(90, 0, 90), (90, 0, -90), (90, 45, 30), (120, 350, -80), (120, 408, 45), (100, 12.5, 10)]
out["sph/views"] = np.array(SVIEWS, dtype=np.float64)
pano5 = randn(21, SM["B"], SM["C"], SM["F"], SM["H"], SM["W"])
for i, (fov, th, ph) in enumerate(SVIEWS):
    view_in = randn(300 + i, SM["B"], SM["C"], SM["F"], SM["h"], SM["w"])
    px = PanoramaLatentProxy(T(pano5))
    vw, m = px.get_view_tensor_no_interpolate(fov, th, ph, SM["w"], SM["h"])
    out[f"sph/get_nearest/{i}"] = N(vw)
    out[f"sph/get_nearest_mask/{i}"] = N(m)
    out[f"sph/get_bilinear/{i}"] = N(px.get_view_tensor_interpolate(fov, th, ph, SM["w"], SM["h"]))
    px = PanoramaLatentProxy(T(pano5)); px.set_view_tensor_no_interpolation(T(view_in), fov, th, ph)
    out[f"sph/set_nearest/{i}"] = N(px.get_equirect_tensor())
    px = PanoramaLatentProxy(T(pano5)); px.set_view_tensor_bilinear(T(view_in), fov, th, ph)
    out[f"sph/set_bilinear/{i}"] = N(px.get_equirect_tensor())
    px = PanoramaLatentProxy(T(pano5)); px.set_view_tensor(T(view_in), fov, th, ph)
    out[f"sph/set_round/{i}"] = N(px.get_equirect_tensor())
# PanoramaTensor on a 3-D image and a [1,H,W] mask
img = randn(22, 3, 16, 32)
for i, (fov, th, ph) in enumerate(SVIEWS[:3]):
    pt = PanoramaTensor(T(img))
    out[f"sph/img_get/{i}"] = N(pt.get_view_tensor_no_interpolate(fov, th, ph, 10, 6)[0])
    out[f"sph/img_get_bilinear/{i}"] = N(pt.get_view_tensor_interpolate(fov, th, ph, 10, 6))
    pt.set_view_tensor_no_interpolation(T(randn(400 + i, 3, 6, 10)), fov, th, ph)
    out[f"sph/img_set/{i}"] = N(pt.equirect_tensor)
    pt = PanoramaTensor(T(img)); pt.set_view_tensor(T(randn(400 + i, 3, 6, 10)), fov, th, ph)
    out[f"sph/img_set_round/{i}"] = N(pt.equirect_tensor)
    pt = PanoramaTensor(T(img)); pt.set_view_tensor_bilinear(T(randn(400 + i, 3, 6, 10)), fov, th, ph)
    out[f"sph/img_set_bilinear/{i}"] = N(pt.equirect_tensor)
    pm = PanoramaTensor(torch.zeros(1, 16, 32)); pm.set_view_tensor_no_interpolation(torch.ones(1, 6, 10), fov, th, ph)
    out[f"sph/mask_set/{i}"] = N(pm.equirect_tensor)

# ---- 6. frame-ring sphere proxies ------------------------------------------
RS = dict(B=1, C=2, F=7, H=16, W=32, h=6, w=10)
RCASES = [(90, 45, 30, 0, 7), (120, 350, -80, 2, 5), (120, 408, 45, 5, 9), (100, 12.5, 10, 6, 13), (90, 0, 90, 3, 10)]
out["rsph/cases"] = np.array(RCASES, dtype=np.float64)
rp5 = randn(31, RS["B"], RS["C"], RS["F"], RS["H"], RS["W"])
for i, (fov, th, ph, fb, fe) in enumerate(RCASES):
    nf = fe - fb
    view_in = randn(500 + i, RS["B"], RS["C"], nf, RS["h"], RS["w"])
    px = RingPanoramaLatentProxy(T(rp5))
    out[f"rsph/get_nearest/{i}"] = N(px.get_view_tensor_no_interpolate(fov, th, ph, RS["w"], RS["h"], frame_begin=fb, frame_end=fe)[0])
    out[f"rsph/get_bilinear/{i}"] = N(px.get_view_tensor_interpolate(fov, th, ph, RS["w"], RS["h"], frame_begin=fb, frame_end=fe))
    px = RingPanoramaLatentProxy(T(rp5)); px.set_view_tensor_no_interpolation(T(view_in), fov, th, ph, frame_begin=fb, frame_end=fe)
    out[f"rsph/set_nearest/{i}"] = N(px.get_equirect_tensor())
    px = RingPanoramaLatentProxy(T(rp5)); px.set_view_tensor_bilinear(T(view_in), fov, th, ph, frame_begin=fb, frame_end=fe)
    out[f"rsph/set_bilinear/{i}"] = N(px.get_equirect_tensor())
    if (RS["h"] * RS["w"]) % nf == 0:
        px = RingPanoramaLatentProxy(T(rp5)); px.set_view_tensor(T(view_in), fov, th, ph, frame_begin=fb, frame_end=fe)
        out[f"rsph/set_round/{i}"] = N(px.get_equirect_tensor())

# ---- 7. mask mix -----------------------------------------------------------
l1, add = randn(41, 1, 4, 5, 6, 10), randn(42, 1, 4, 5, 6, 10)
m3 = (randn(43, 1, 6, 10) > 0).astype(np.float32)
m5 = (randn(44, 1, 4, 5, 6, 10) > 0).astype(np.float32)
for i, (m, r) in enumerate([(m3, 0.5), (m3, 0.3), (m5, 0.8), (m5, 1.0), (m3, 0.0)]):
    out[f"mix/{i}"] = N(mix_latents_with_mask(T(l1), T(add), T(m), r))
out["mix/ratios"] = np.array([0.5, 0.3, 0.8, 1.0, 0.0])

# ---- 8. step compositions, small -------------------------------------------
# 8a sequential sphere step exactly as t2v_sphere_panorama_pipeline.py:176-291 composes it
#    (get latent, get mask, mix with injected re-noise, injected denoiser, 3 sets).
sviews = [(90, 0), (-90, 0), (45, 10), (45, 130), (45, 250), (0, 10), (0, 130), (0, 250), (-45, 70)]
out["step/views"] = np.array(sviews, dtype=np.float64)
lat = T(randn(51, 1, 2, 3, 16, 32))
h_lat = PanoramaLatentProxy(lat)
h_den = PanoramaLatentProxy(torch.zeros_like(lat))
h_msk = PanoramaTensor(torch.zeros_like(lat[0, 0, [0]]))
for n, (ph, th) in enumerate(sviews):
    vw, _ = h_lat.get_view_tensor_no_interpolate(fov=120, theta=th, phi=ph, width=10, height=6)
    mk, _ = h_msk.get_view_tensor_no_interpolate(fov=120, theta=th, phi=ph, width=10, height=6)
    noised = vw * 0.9 + T(randn(600 + n, 1, 2, 3, 6, 10)) * 0.4359
    vw = mix_latents_with_mask(latent_1=vw, latent_to_add=noised, mask=mk, mix_ratio=0.6)
    x_prev = vw * 0.5 + T(randn(700 + n, 1, 2, 3, 6, 10))
    x0 = vw * 0.25 - T(randn(800 + n, 1, 2, 3, 6, 10))
    h_lat.set_view_tensor_no_interpolation(view_tensor=x_prev, fov=120, theta=th, phi=ph)
    h_den.set_view_tensor_no_interpolation(view_tensor=x0, fov=120, theta=th, phi=ph)
    h_msk.set_view_tensor_no_interpolation(view_tensor=torch.ones_like(x_prev[0, 0, [0]]), fov=120, theta=th, phi=ph)
out["step/seq/latent"] = N(h_lat.get_equirect_tensor())
out["step/seq/denoised"] = N(h_den.get_equirect_tensor())
out["step/seq/mask"] = N(h_msk.equirect_tensor)

# 8b batched order (gather-all, then merge-all) with the reference classes, nearest + bilinear
for mode in ("nearest", "bilinear"):
    h_lat = PanoramaLatentProxy(lat)
    h_den = PanoramaLatentProxy(torch.zeros_like(lat))
    h_msk = PanoramaTensor(torch.zeros_like(lat[0, 0, [0]]))
    wins = []
    for (ph, th) in sviews:
        if mode == "nearest":
            wins.append(h_lat.get_view_tensor_no_interpolate(120, th, ph, 10, 6)[0].clone())
        else:
            wins.append(h_lat.get_view_tensor_interpolate(120, th, ph, 10, 6))
    for n, (ph, th) in enumerate(sviews):
        x_prev = wins[n] * 0.5 + T(randn(700 + n, 1, 2, 3, 6, 10))
        x0 = wins[n] * 0.25 - T(randn(800 + n, 1, 2, 3, 6, 10))
        put = "set_view_tensor_no_interpolation" if mode == "nearest" else "set_view_tensor_bilinear"
        getattr(h_lat, put)(x_prev, 120, th, ph)
        getattr(h_den, put)(x0, 120, th, ph)
        getattr(h_msk, put)(torch.ones(1, 6, 10), 120, th, ph)
    out[f"step/batched_{mode}/windows"] = N(torch.stack(wins))
    out[f"step/batched_{mode}/latent"] = N(h_lat.get_equirect_tensor())
    out[f"step/batched_{mode}/denoised"] = N(h_den.get_equirect_tensor())
    out[f"step/batched_{mode}/mask"] = N(h_msk.equirect_tensor)

np.savez_compressed(os.path.join(HERE, "small.npz"), **out)
print("small.npz:", len(out), "arrays")

# ---- 9. full-size configs: digests -----------------------------------------
big = {}

# config 1: ring planar [1,4,16,64,128], window 40x64 stride 8, step i=3
i_step = 3
lat = T(randn(0, 1, 4, 16, 64, 128))
wins_desc = [((i_step % 16) * 4 + 8 * a, (i_step % 16) * 1 + 8 * b) for a in range(16) for b in range(8)]
big["cfg1/windows"] = np.array(wins_desc, dtype=np.int64)
r_lat, r_den, r_msk = RingLatent(lat), RingLatent(torch.zeros_like(lat)), RingLatent(torch.zeros_like(lat))
wins = [r_lat.get_window_latent(l, l + 64, t, t + 40, 0, 16) for (l, t) in wins_desc]
for n, (l, t) in enumerate(wins_desc):
    r_lat.set_window_latent(T(randn(1000 + n, 1, 4, 16, 40, 64)), l, l + 64, t, t + 40, 0, 16)
    r_den.set_window_latent(T(randn(2000 + n, 1, 4, 16, 40, 64)), l, l + 64, t, t + 40, 0, 16)
    r_msk.set_window_latent(torch.ones(1, 4, 16, 40, 64), l, l + 64, t, t + 40, 0, 16)
big["cfg1/windows_sha"] = digest(N(torch.stack(wins)))
big["cfg1/latent_sha"] = digest(N(r_lat.torch_latent))
big["cfg1/denoised_sha"] = digest(N(r_den.torch_latent))
big["cfg1/mask_sha"] = digest(N(r_msk.torch_latent))
big["cfg1/latent_sample"] = N(r_lat.torch_latent).reshape(-1)[::4099]

# config 2: sphere [1,4,16,128,256], 44 default views, step i=1 (theta offset 12), nearest + bilinear
i_step = 1
views = po.default_sphere_views(i_step)
big["cfg2/views"] = np.array(views, dtype=np.float64)
lat = T(randn(0, 1, 4, 16, 128, 256))
idx_r, idx_w = [], []
for (ph, th) in views:
    pt = PanoramaTensor(torch.zeros(1, 128, 256))
    u, v = pt._get_uv(fov=120, theta=th, phi=ph, width=64, height=40)
    u0 = torch.floor(u).long(); v0 = torch.floor(v).long()
    idx_r.append(N((torch.clamp(v0, 0, 127) * 256 + u0 % 256).reshape(-1)))
    idx_w.append(N((v0 * 256 + u0).reshape(-1)))
big["cfg2/idx_read"] = np.stack(idx_r).astype(np.int32)
big["cfg2/idx_write"] = np.stack(idx_w).astype(np.int32)
for mode in ("nearest", "bilinear"):
    h_lat = PanoramaLatentProxy(lat)
    h_den = PanoramaLatentProxy(torch.zeros_like(lat))
    h_msk = PanoramaTensor(torch.zeros(1, 128, 256))
    wins = []
    for (ph, th) in views:
        if mode == "nearest":
            wins.append(h_lat.get_view_tensor_no_interpolate(120, th, ph, 64, 40)[0].clone())
        else:
            wins.append(h_lat.get_view_tensor_interpolate(120, th, ph, 64, 40))
    put = "set_view_tensor_no_interpolation" if mode == "nearest" else "set_view_tensor_bilinear"
    for n, (ph, th) in enumerate(views):
        getattr(h_lat, put)(T(randn(1000 + n, 1, 4, 16, 40, 64)), 120, th, ph)
        getattr(h_den, put)(T(randn(2000 + n, 1, 4, 16, 40, 64)), 120, th, ph)
        getattr(h_msk, put)(torch.ones(1, 40, 64), 120, th, ph)
    W_ = N(torch.stack(wins))
    big[f"cfg2/{mode}/windows_sha"] = digest(W_)
    big[f"cfg2/{mode}/windows_sample"] = W_.reshape(-1)[::4099]
    for name, hnd in (("latent", h_lat.get_equirect_tensor()), ("denoised", h_den.get_equirect_tensor()),
                      ("mask", h_msk.equirect_tensor)):
        a = N(hnd)
        big[f"cfg2/{mode}/{name}_sha"] = digest(a)
        big[f"cfg2/{mode}/{name}_sample"] = a.reshape(-1)[::(4099 if a.size > 100000 else 7)]

# config 4 slice: frame-ring sphere [1,4,64,128,256], 2 wrapping frame windows x 5 views, nearest
lat = T(randn(4, 1, 4, 64, 128, 256))
jobs = [(fb, ph, th) for fb in (52, 60) for (ph, th) in views[:5]]
big["cfg4/jobs"] = np.array(jobs, dtype=np.float64)
h_lat = RingPanoramaLatentProxy(lat)
wins = [h_lat.get_view_tensor_no_interpolate(120, th, ph, 64, 40, frame_begin=fb, frame_end=fb + 16)[0].clone()
        for (fb, ph, th) in jobs]
for n, (fb, ph, th) in enumerate(jobs):
    h_lat.set_view_tensor_no_interpolation(T(randn(1000 + n, 1, 4, 16, 40, 64)), 120, th, ph,
                                           frame_begin=fb, frame_end=fb + 16)
big["cfg4/windows_sha"] = digest(N(torch.stack(wins)))
big["cfg4/latent_sha"] = digest(N(h_lat.get_equirect_tensor()))
big["cfg4/latent_sample"] = N(h_lat.get_equirect_tensor()).reshape(-1)[::8209]

np.savez_compressed(os.path.join(HERE, "configs.npz"), **big)
print("configs.npz:", len(big), "arrays")
for f in ("small.npz", "configs.npz"):
    print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
