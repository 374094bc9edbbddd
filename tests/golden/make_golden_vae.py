"""Golden vectors for the pixel-side neighbours of the path -> tests/golden/vae.npz.

Runs the UNMODIFIED ``VC2_Pipeline_I2V_SpherePano.tiled_vae_encode_tensor_simple``
(pipeline/i2v_sphere_panorama_pipeline.py:505-562) imported from /root/reference with a deterministic
stand-in encoder (8x8 average pooling + a fixed 3->4 channel mix: the real VAE weights are not
available and do not matter for the tiling arithmetic), and restates the inline wrap-padded chunk
logic of :974-991 with the same torch.chunk / torch.cat calls.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_vae.py
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import _ref_stubs  # noqa: E402,F401  (registers the pytorch_lightning / diffusers stubs, adds the reference to sys.path)
from conftest import randn  # noqa: E402
import builtins  # noqa: E402
_print, builtins.print = builtins.print, (lambda *a, **k: None)
from pipeline.i2v_sphere_panorama_pipeline import VC2_Pipeline_I2V_SpherePano  # noqa: E402
builtins.print = _print

MIX = torch.tensor([[0.5, -0.25, 0.125], [0.1, 0.2, 0.3], [-0.4, 0.3, 0.2], [0.05, 0.6, -0.35]])


def encode(x):
    """[B, 3, F, H, W] -> [B, 4, F, H/8, W/8]; also depends on the tile's absolute content only."""
    B, C, Fr, H, W = x.shape
    p = F.avg_pool2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), 8)
    y = torch.einsum("oc,nchw->nohw", MIX, p) + 0.01 * p.mean(dim=(1, 2, 3), keepdim=True)   # tile-dependent term
    return y.view(B, Fr, 4, H // 8, W // 8).permute(0, 2, 1, 3, 4)


out = {}
fake = types.SimpleNamespace(vae_scale_factor=8, pretrained_t2v=types.SimpleNamespace(encode_first_stage_2DAE=encode))
for n, (shape, kw) in enumerate([((1, 3, 1, 256, 512), dict(overlap_h=4, overlap_w=4)),
                                 ((1, 3, 2, 256, 512), dict(h_tile_num=2, w_tile_num=4, overlap_h=8, overlap_w=2)),
                                 ((2, 3, 1, 128, 256), dict(h_tile_num=4, w_tile_num=2, overlap_h=1, overlap_w=3))]):
    img = torch.from_numpy(randn(600 + n, *shape))
    lat = VC2_Pipeline_I2V_SpherePano.tiled_vae_encode_tensor_simple(fake, image_tensor=img, **kw)
    out[f"encode/{n}/shape"] = np.array(shape)
    out[f"encode/{n}/kw"] = np.array([kw.get("h_tile_num", 4), kw.get("w_tile_num", 4), kw["overlap_h"], kw["overlap_w"]])
    out[f"encode/{n}/latent"] = lat.numpy()

# wrap-padded chunked decode (i2v_sphere_panorama_pipeline.py:974-991), same torch calls, a x2 "decoder"
for n, (shape, chunk_num) in enumerate([((1, 4, 3, 8, 64), 16), ((1, 4, 2, 6, 40), 16), ((1, 4, 2, 4, 36), 8)]):
    denoised = torch.from_numpy(randn(700 + n, *shape))
    chunked = list(torch.chunk(denoised, chunk_num, dim=4))
    padded = torch.cat([chunked[-1]] + chunked + [chunked[0]], dim=4)
    frames = [F.interpolate(padded[:, :, f], scale_factor=2, mode="nearest").unsqueeze(2) * (f + 1) for f in range(shape[2])]
    videos = torch.cat(frames, dim=2)
    vc = torch.chunk(videos, chunk_num + 2, dim=4)
    final = torch.cat(vc[1:-1], dim=4)
    out[f"decode/{n}/shape"] = np.array(shape)
    out[f"decode/{n}/chunk_num"] = np.array(chunk_num)
    out[f"decode/{n}/padded"] = padded.numpy()
    out[f"decode/{n}/videos"] = final.numpy()
np.savez_compressed(os.path.join(HERE, "vae.npz"), **out)
print("vae.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "vae.npz")) // 1024, "KiB")
