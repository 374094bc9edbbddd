"""Pixel-side neighbours of the path (dynamicscaler_b200/vae_tiling.py, csrc/tiles.cu)
against tests/golden/vae.npz (the reference's own ``tiled_vae_encode_tensor_simple`` and the inline
wrap-padded decode of i2v_sphere_panorama_pipeline.py:974-991) and the oracle.  ``pytest -m gpu``."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, randn
from oracle import panorama_oracle as po

pytestmark = pytest.mark.gpu

MIX = [[0.5, -0.25, 0.125], [0.1, 0.2, 0.3], [-0.4, 0.3, 0.2], [0.05, 0.6, -0.35]]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from dynamicscaler_b200 import _lib, lib
    assert lib.ds_abi_version() == _lib.ABI_VERSION
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def vae():
    return np.load(os.path.join(GOLDEN, "vae.npz"))


def encode_cpu(x):
    """The stand-in encoder of tests/golden/make_golden_vae.py (deterministic on the CPU)."""
    B, C, Fr, H, W = x.shape
    p = F.avg_pool2d(x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, H, W), 8)
    y = torch.einsum("oc,nchw->nohw", torch.tensor(MIX), p) + 0.01 * p.mean(dim=(1, 2, 3), keepdim=True)
    return y.view(B, Fr, 4, H // 8, W // 8).permute(0, 2, 1, 3, 4)


def test_tiled_encode_matches_reference(vae, dev):
    """Crop + float32 accumulate + count + normalise on the GPU == the reference method, bit for bit
    (the encoder runs on the CPU exactly as in the golden script, its tiles are uploaded)."""
    from dynamicscaler_b200.vae_tiling import tiled_vae_encode_tensor_simple
    for n in range(3):
        shape = tuple(int(v) for v in vae[f"encode/{n}/shape"])
        ht, wt, oh, ow = (int(v) for v in vae[f"encode/{n}/kw"])
        img = torch.from_numpy(randn(600 + n, *shape))
        got = tiled_vae_encode_tensor_simple(img.to(dev), lambda t: encode_cpu(t.cpu()).to(dev), h_tile_num=ht,
                                             w_tile_num=wt, overlap_h=oh, overlap_w=ow)
        assert got.dtype == torch.float32
        assert np.array_equal(got.cpu().numpy(), vae[f"encode/{n}/latent"]), n
        ref = po.tiled_encode(img.numpy(), lambda a: encode_cpu(torch.from_numpy(np.ascontiguousarray(a))).numpy(),
                              h_tile_num=ht, w_tile_num=wt, overlap_h=oh, overlap_w=ow)
        assert np.array_equal(ref, vae[f"encode/{n}/latent"]), n       # the oracle restatement is pinned too


def test_tiled_encode_bf16_tiles_accumulate_in_fp32(dev):
    from dynamicscaler_b200.vae_tiling import tiled_vae_encode_tensor_simple
    img = torch.from_numpy(randn(610, 1, 3, 1, 128, 256)).to(dev)
    enc = lambda t: encode_cpu(t.float().cpu()).to(dev).to(torch.bfloat16)
    got = tiled_vae_encode_tensor_simple(img, enc, h_tile_num=2, w_tile_num=2, overlap_h=2, overlap_w=2)
    ref = po.tiled_encode(img.cpu().numpy(), lambda a: encode_cpu(torch.from_numpy(np.ascontiguousarray(a))).to(torch.bfloat16).float().numpy(),
                          h_tile_num=2, w_tile_num=2, overlap_h=2, overlap_w=2)
    assert got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), ref)


def test_static_latent_cache_encodes_once(dev):
    from dynamicscaler_b200.vae_tiling import StaticLatentCache
    cache, calls = StaticLatentCache(), {"n": 0}

    def enc(t):
        calls["n"] += 1
        return encode_cpu(t.cpu()).to(dev)

    img = torch.from_numpy(randn(611, 1, 3, 1, 128, 256)).to(dev)
    kw = dict(h_tile_num=2, w_tile_num=2, overlap_h=2, overlap_w=2)
    a = cache.get(("pano.png", (128, 256)), lambda: img, enc, **kw)
    for _ in range(5):                                               # the next denoising steps (i2v_...:245)
        b = cache.get(("pano.png", (128, 256)), lambda: img, enc, **kw)
    assert b is a and calls["n"] == 4 and (cache.hits, cache.misses) == (5, 1)


def test_wrap_padded_decode_matches_reference(vae, dev):
    from dynamicscaler_b200.vae_tiling import crop_wrap_padded, decode_wrap_padded, wrap_pad_latent
    for n in range(3):
        shape = tuple(int(v) for v in vae[f"decode/{n}/shape"])
        chunk_num = int(vae[f"decode/{n}/chunk_num"])
        den = torch.from_numpy(randn(700 + n, *shape)).to(dev)
        padded = wrap_pad_latent(den, chunk_num)
        assert np.array_equal(padded.cpu().numpy(), vae[f"decode/{n}/padded"]), n
        assert np.array_equal(po.wrap_pad(den.cpu().numpy(), chunk_num), vae[f"decode/{n}/padded"]), n
        state = {"f": 0}

        def dec(x):                                                  # the x2 "decoder" of the golden script, per frame
            state["f"] += 1
            return F.interpolate(x[:, :, 0], scale_factor=2, mode="nearest").unsqueeze(2) * state["f"]

        vid = decode_wrap_padded(den, dec, chunk_num)
        assert np.array_equal(vid.cpu().numpy(), vae[f"decode/{n}/videos"]), n
        full = torch.from_numpy(vae[f"decode/{n}/padded"]).to(dev)
        assert np.array_equal(crop_wrap_padded(full, chunk_num).cpu().numpy(),
                              po.crop_wrap_padded(vae[f"decode/{n}/padded"], chunk_num)), n


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_repeat_renoise_vs_oracle(dev, dtype):
    """paste_on_static: cat([frame_0] * total_f, dim=2) + re_noise (i2v_...:248-252) in one pass."""
    from dynamicscaler_b200.step import renoise_coefficients
    from dynamicscaler_b200.vae_tiling import repeat_renoise
    ac = np.linspace(0.999, 0.01, 1000).astype(np.float32)
    c, s = renoise_coefficients(ac, np.array([1, 334, 667]), 0, 2)
    f0 = torch.from_numpy(randn(620, 1, 4, 1, 16, 32)).to(dev).to(dtype)
    eps = torch.from_numpy(randn(621, 1, 4, 5, 16, 32)).to(dev).to(dtype)
    got = repeat_renoise(f0, 5, c, s, eps)
    rep = torch.cat([f0] * 5, dim=2)
    assert torch.equal(repeat_renoise(f0, 5, 1.0, 0.0, None), rep)
    # the reference's ops: c, s are float32 0-dim tensors (pipeline/scheduler.py:104-105), x_a / eps have the latent dtype
    want = torch.tensor(c, device=dev) * rep + torch.tensor(s, device=dev) * eps
    assert torch.equal(got, want)
    if dtype == torch.float32:
        oc, os_ = po.renoise_scalars(ac, np.array([1, 334, 667]), 0, 2)
        assert np.array_equal(got.cpu().numpy(), po.re_noise(rep.cpu().numpy(), eps.cpu().numpy(), oc, os_))
