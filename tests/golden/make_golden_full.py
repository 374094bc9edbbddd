"""Full-size digests of the remaining BASELINE configs, from the UNMODIFIED reference
(imported from /root/reference, CPU, fp32):

    cfg3  sphere [1,4,16,256,512], 44 default views, step i=1, nearest + bilinear
    cfg5  planar ring [1,4,16,512,1024], 16x16 windows of 40x64, step i=3
    cfg4  frame-ring sphere [1,4,64,128,256], 13 wrapping frame windows x 44 views (572 jobs),
          step i=1, nearest + bilinear, latent-shaped mask (I2V)

all in the batched order (every window read from the step-start panorama, then all windows
written back in the reference's loop order).  Only sha256 digests and strided samples are
stored (tests/golden/full.npz).  Run once in the build container (a few minutes of CPU):

    python tests/golden/make_golden_full.py

Inputs: ``numpy.random.default_rng(seed).standard_normal(dtype=float32)`` -- panorama seed per
config below, payload n of a step: seeds 1000 + n (x_prev) and 2000 + n (pred_x0).
"""
import hashlib
import os
import sys
import time
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

from utils.shift_window_utils import RingLatent  # noqa: E402
from utils.panorama_tensor_utils import PanoramaTensor, PanoramaLatentProxy  # noqa: E402
from utils.ring_panorama_tensor_utils import RingPanoramaLatentProxy  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import panorama_oracle as po  # noqa: E402  (window geometry helpers only)


def randn(seed, *shape):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def N(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def record(big, prefix, name, a, stride=8209):
    a = N(a) if isinstance(a, torch.Tensor) else a
    big[f"{prefix}/{name}_sha"] = digest(a)
    big[f"{prefix}/{name}_sample"] = a.reshape(-1)[::stride].copy()


big = {}
VSHAPE = (1, 4, 16, 40, 64)
t0 = time.time()

# ---- cfg3: north-star shape, sphere 256x512 ------------------------------------
views = po.default_sphere_views(1)
big["cfg3/views"] = np.array(views, dtype=np.float64)
lat = T(randn(3, 1, 4, 16, 256, 512))
for mode in ("nearest", "bilinear"):
    h_lat = PanoramaLatentProxy(lat)
    h_den = PanoramaLatentProxy(torch.zeros_like(lat))
    h_msk = PanoramaTensor(torch.zeros(1, 256, 512))
    wins = []
    for (ph, th) in views:
        if mode == "nearest":
            wins.append(h_lat.get_view_tensor_no_interpolate(120, th, ph, 64, 40)[0].clone())
        else:
            wins.append(h_lat.get_view_tensor_interpolate(120, th, ph, 64, 40))
    put = "set_view_tensor_no_interpolation" if mode == "nearest" else "set_view_tensor_bilinear"
    for n, (ph, th) in enumerate(views):
        getattr(h_lat, put)(T(randn(1000 + n, *VSHAPE)), 120, th, ph)
        getattr(h_den, put)(T(randn(2000 + n, *VSHAPE)), 120, th, ph)
        getattr(h_msk, put)(torch.ones(1, 40, 64), 120, th, ph)
    record(big, f"cfg3/{mode}", "windows", torch.stack(wins))
    record(big, f"cfg3/{mode}", "latent", h_lat.get_equirect_tensor())
    record(big, f"cfg3/{mode}", "denoised", h_den.get_equirect_tensor())
    record(big, f"cfg3/{mode}", "mask", h_msk.equirect_tensor, stride=61)
    print(f"cfg3 {mode} done at {time.time() - t0:.0f} s", flush=True)

# ---- cfg5: 4K-scale planar ring 512x1024, 16x16 windows -------------------------
i_step = 3
windows = po.ring_window_grid(i_step, 16, 16, 8192, 4096)
big["cfg5/windows"] = np.array(windows, dtype=np.int64)
lat = T(randn(5, 1, 4, 16, 512, 1024))
r_lat, r_den, r_msk = RingLatent(lat), RingLatent(torch.zeros_like(lat)), RingLatent(torch.zeros_like(lat))
wins = [r_lat.get_window_latent(l, l + 64, t, t + 40, 0, 16) for (l, t) in windows]
for n, (l, t) in enumerate(windows):
    r_lat.set_window_latent(T(randn(1000 + n, *VSHAPE)), l, l + 64, t, t + 40, 0, 16)
    r_den.set_window_latent(T(randn(2000 + n, *VSHAPE)), l, l + 64, t, t + 40, 0, 16)
    r_msk.set_window_latent(torch.ones(*VSHAPE), l, l + 64, t, t + 40, 0, 16)
record(big, "cfg5", "windows", torch.stack(wins))
record(big, "cfg5", "latent", r_lat.torch_latent)
record(big, "cfg5", "denoised", r_den.torch_latent)
record(big, "cfg5", "mask", r_msk.torch_latent)
print(f"cfg5 done at {time.time() - t0:.0f} s", flush=True)
del r_lat, r_den, r_msk, wins

# ---- cfg4: frame-ring sphere, all 13 frame windows x 44 views --------------------
views = po.default_sphere_views(1)
fbs = po.frame_window_begins(1, frames=16, total_f=64, overlap=0.75, loop_step_frame=8)
assert len(fbs) == 13
jobs = [(fb, ph, th) for fb in fbs for (ph, th) in views]            # frame-window-major (i2v_..:284-325)
big["cfg4/frame_begins"] = np.array(fbs, dtype=np.int64)
big["cfg4/views"] = np.array(views, dtype=np.float64)
lat = T(randn(4, 1, 4, 64, 128, 256))
for mode in ("nearest", "bilinear"):
    h_lat = RingPanoramaLatentProxy(lat)
    h_den = RingPanoramaLatentProxy(torch.zeros_like(lat))
    h_msk = RingPanoramaLatentProxy(torch.zeros_like(lat))
    sha = hashlib.sha256()
    sample = []
    for (fb, ph, th) in jobs:                                        # windows digest streamed job by job
        if mode == "nearest":
            w_ = h_lat.get_view_tensor_no_interpolate(120, th, ph, 64, 40, frame_begin=fb, frame_end=fb + 16)[0]
        else:
            w_ = h_lat.get_view_tensor_interpolate(120, th, ph, 64, 40, frame_begin=fb, frame_end=fb + 16)
        a = N(w_)
        assert a.shape == VSHAPE
        sha.update(a.tobytes())
        sample.append(a.reshape(-1)[::8209].copy())
    big[f"cfg4/{mode}/windows_sha"] = np.frombuffer(sha.digest(), dtype=np.uint8)
    big[f"cfg4/{mode}/windows_sample"] = np.concatenate(sample)
    put = "set_view_tensor_no_interpolation" if mode == "nearest" else "set_view_tensor_bilinear"
    ones = torch.ones(*VSHAPE)
    for n, (fb, ph, th) in enumerate(jobs):
        kw = dict(frame_begin=fb, frame_end=fb + 16)
        getattr(h_lat, put)(T(randn(1000 + n, *VSHAPE)), 120, th, ph, **kw)
        getattr(h_den, put)(T(randn(2000 + n, *VSHAPE)), 120, th, ph, **kw)
        getattr(h_msk, put)(ones, 120, th, ph, **kw)
    record(big, f"cfg4/{mode}", "latent", h_lat.get_equirect_tensor())
    record(big, f"cfg4/{mode}", "denoised", h_den.get_equirect_tensor())
    record(big, f"cfg4/{mode}", "mask", h_msk.get_equirect_tensor())
    print(f"cfg4 {mode} done at {time.time() - t0:.0f} s", flush=True)

np.savez_compressed(os.path.join(HERE, "full.npz"), **big)
print("full.npz:", len(big), "arrays,", os.path.getsize(os.path.join(HERE, "full.npz")) // 1024, "KiB")
