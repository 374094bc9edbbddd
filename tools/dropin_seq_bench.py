"""Regime (i) of SURVEY §7: the reference's UNMODIFIED sequential call sites (get latent, get mask,
mix, set latent / denoised / mask per view; t2v_sphere_panorama_pipeline.py:176-291) through the
drop-in classes, one kernel launch per call.  cfg2 geometry, 44 views; prints ms per step."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicscaler_b200.geometry import sphere_views  # noqa: E402
from dynamicscaler_b200.panorama_tensor_utils import PanoramaLatentProxy, PanoramaTensor  # noqa: E402
from dynamicscaler_b200.tensor_utils import mix_latents_with_mask  # noqa: E402

dev = torch.device("cuda:0")
lat = torch.randn(1, 4, 16, 128, 256, device=dev)
h_lat, h_den = PanoramaLatentProxy(lat), PanoramaLatentProxy(torch.zeros_like(lat))
xp = torch.randn(1, 4, 16, 40, 64, device=dev)
x0 = torch.randn(1, 4, 16, 40, 64, device=dev)
ones = torch.ones(1, 40, 64, device=dev)


def step(i):
    h_msk = PanoramaTensor(torch.zeros(1, 128, 256, device=dev))      # re-created every step (:159)
    for ph, th in sphere_views(i):
        vw, _ = h_lat.get_view_tensor_no_interpolate(fov=120, theta=th, phi=ph, width=64, height=40)
        mk, _ = h_msk.get_view_tensor_no_interpolate(fov=120, theta=th, phi=ph, width=64, height=40)
        vw = mix_latents_with_mask(latent_1=vw, latent_to_add=vw, mask=mk, mix_ratio=0.5)
        h_lat.set_view_tensor_no_interpolation(view_tensor=xp, fov=120, theta=th, phi=ph)
        h_den.set_view_tensor_no_interpolation(view_tensor=x0, fov=120, theta=th, phi=ph)
        h_msk.set_view_tensor_no_interpolation(view_tensor=ones, fov=120, theta=th, phi=ph)


for i in range(10):
    step(i)                      # fills the map caches for the 10 longitude offsets
torch.cuda.synchronize()
t = time.perf_counter()
n = 30
for i in range(n):
    step(i)
torch.cuda.synchronize()
ms = (time.perf_counter() - t) / n * 1e3
print(json.dumps({"regime": "sequential drop-in call sites", "views": 44, "calls_per_step": 44 * 6,
                  "ms_per_step": round(ms, 3), "us_per_call": round(ms * 1e3 / (44 * 6), 1)}))
