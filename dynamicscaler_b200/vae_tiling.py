"""Pixel-side neighbours of the tiling path: the data movement around the VAE calls of
``pipeline/i2v_sphere_panorama_pipeline.py``.  The VAE stays the reference's torch module (the caller
passes ``encode_fn`` / ``decode_fn`` = ``pretrained_t2v.encode_first_stage_2DAE`` /
``decode_first_stage_2DAE``); what runs here is the crop / accumulate / normalise, the wrap padding
and the repeat + re-noise, each one kernel launch (csrc/tiles.cu, csrc/ring.cu).

    tiled_vae_encode_tensor_simple   :505-562   tile crop -> float32 accumulate + count -> normalise
    StaticLatentCache                :245-246   the per-step re-encode of the SAME image, cached
    repeat_renoise                   :248-252   cat([frame_0] * total_f, dim=2) + re_noise, one pass
    wrap_pad_latent / crop_wrap_padded / decode_wrap_padded   :974-991   chunked decode with wrap padding
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

from . import _lib
from ._lib import lib


def tiled_vae_encode_tensor_simple(image_tensor: torch.Tensor, encode_fn: Callable, vae_scale_factor: int = 8,
                                   h_tile_num: int = 4, w_tile_num: int = 4, overlap_h: int = 32, overlap_w: int = 32
                                   ) -> torch.Tensor:
    """Reference ``tiled_vae_encode_tensor_simple`` (i2v_sphere_panorama_pipeline.py:505-562) with the
    encoder passed in.  image_tensor [B, C_img, F, H_dec, W_dec] -> float32 latent [B, 4, F, H/8, W/8]."""
    _lib.require_cuda(image_tensor, "image_tensor")
    B, C_img, F, H_dec, W_dec = image_tensor.shape
    sf = vae_scale_factor
    H_lat, W_lat = H_dec // sf, W_dec // sf
    th_lat, tw_lat = H_lat // h_tile_num, W_lat // w_tile_num
    th_img, tw_img = th_lat * sf, tw_lat * sf
    ov_h_img, ov_w_img = overlap_h * sf, overlap_w * sf
    dev = image_tensor.device
    planes = B * 4 * F
    acc = torch.zeros((B, 4, F, H_lat, W_lat), device=dev, dtype=torch.float32)          # :521
    count = torch.zeros((B, 1, 1, H_lat, W_lat), device=dev, dtype=torch.float32)        # :523
    sp = _lib.stream_ptr(dev)
    with _lib.on(dev):
        for i in range(h_tile_num):
            for j in range(w_tile_num):
                h0, h1, w0, w1 = i * th_img, (i + 1) * th_img, j * tw_img, (j + 1) * tw_img
                h0o, h1o = max(h0 - ov_h_img, 0), min(h1 + ov_h_img, H_dec)
                w0o, w1o = max(w0 - ov_w_img, 0), min(w1 + ov_w_img, W_dec)
                tile = encode_fn(image_tensor[:, :, :, h0o:h1o, w0o:w1o]).contiguous()   # [B, 4, F, h_sub, w_sub]
                assert tile.shape[:3] == (B, 4, F), f"encoder returned {tuple(tile.shape)}"
                top_cut, left_cut = (h0 - h0o) // sf, (w0 - w0o) // sf
                bottom_cut = tile.shape[3] - (h1o - h1) // sf
                right_cut = tile.shape[4] - (w1o - w1) // sf
                ch, cw = bottom_cut - top_cut, right_cut - left_cut
                # the reference's slice-assign raises when the cropped tile and the core region differ
                assert (ch, cw) == (th_lat, tw_lat), f"cropped tile {(ch, cw)} does not match core {(th_lat, tw_lat)}"
                lib.check(lib.ds_tile_accumulate(acc.data_ptr(), count.data_ptr(), tile.data_ptr(), planes, 4 * F,
                                                 H_lat, W_lat, tile.shape[3], tile.shape[4], top_cut, left_cut, ch, cw,
                                                 i * th_lat, j * tw_lat, _lib.dtype_code(tile.dtype), sp),
                          "ds_tile_accumulate")
        out = torch.empty_like(acc)
        lib.check(lib.ds_tile_normalise(acc.data_ptr(), count.data_ptr(), out.data_ptr(), planes, 4 * F,
                                        H_lat * W_lat, sp), "ds_tile_normalise")
    return out


class StaticLatentCache:
    """``paste_on_static`` re-encodes the SAME panorama image with 16 VAE tile calls at every denoising
    step (i2v_sphere_panorama_pipeline.py:245-246: ``self.tiled_vae_encode_image(image_path=pano_image_path,
    ...)`` inside the loop).  The result does not depend on the step, so it is encoded once per
    (key, size, dtype) and reused."""

    def __init__(self):
        self._store = {}
        self.hits = 0
        self.misses = 0

    def get(self, key, image_tensor_fn: Callable[[], torch.Tensor], encode_fn: Callable, **tile_kwargs) -> torch.Tensor:
        """``key``: e.g. (image_path, (H, W)); ``image_tensor_fn``: builds the [B, 3, 1, H, W] tensor on a miss."""
        if key in self._store:
            self.hits += 1
            return self._store[key]
        self.misses += 1
        lat = tiled_vae_encode_tensor_simple(image_tensor_fn(), encode_fn, **tile_kwargs)
        self._store[key] = lat
        return lat


def repeat_renoise(frame0_latent: torch.Tensor, total_f: int, c: float, s: float,
                   eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``re_noise(torch.cat([frame_0_latent] * total_f, dim=2), 0, step_b)`` (i2v_...:248-252 +
    pipeline/scheduler.py:98-110) in ONE pass: frame0_latent [B, C, 1, H, W] -> [B, C, total_f, H, W] =
    c * frame0 + s * eps, with (c, s) from ``step.renoise_coefficients`` and the epsilon injected
    (``eps`` None: the plain repeat)."""
    _lib.require_cuda(frame0_latent, "frame0_latent")
    B, C, one, H, W = frame0_latent.shape
    assert one == 1, "frame_0_latent has a single frame"
    x0 = frame0_latent.contiguous()
    out = torch.empty((B, C, total_f, H, W), dtype=x0.dtype, device=x0.device)
    if eps is not None:
        assert eps.shape == out.shape and eps.dtype == x0.dtype and eps.is_contiguous()
    with _lib.on(x0.device):
        lib.check(lib.ds_repeat_renoise(x0.data_ptr(), _lib.ptr(eps), out.data_ptr(), float(c), float(s), B * C, total_f,
                                        H * W, _lib.dtype_code(x0.dtype), _lib.stream_ptr(x0.device)), "ds_repeat_renoise")
    return out


def _chunk_sizes(n: int, chunks: int) -> Tuple[int, int, int]:
    """torch.chunk(t, chunks) along an axis of length n: (chunk size, number of chunks, size of the last)."""
    size = -(-n // chunks)
    count = -(-n // size)
    return size, count, n - size * (count - 1)


def _ring_window(x: torch.Tensor, left: int, width: int) -> torch.Tensor:
    """[..., H, W] -> [..., H, width], columns (left + j) mod W: one ds_ring_gather launch."""
    _lib.require_cuda(x, "tensor")
    x = x.contiguous()
    H, W = x.shape[-2], x.shape[-1]
    planes = x.numel() // (H * W)
    out = torch.empty(x.shape[:-1] + (width,), dtype=x.dtype, device=x.device)
    desc = torch.tensor([[left % W, 0, 0]], dtype=torch.int32, device=x.device)
    with _lib.on(x.device):
        lib.check(lib.ds_ring_gather(x.data_ptr(), out.data_ptr(), desc.data_ptr(), 1, 0, 0, 0, planes, 1, H, W, 1, H, width,
                                     _lib.dtype_code(x.dtype), _lib.stream_ptr(x.device)), "ds_ring_gather(wrap pad)")
    return out


def wrap_pad_latent(denoised: torch.Tensor, chunk_num: int = 16) -> torch.Tensor:
    """i2v_sphere_panorama_pipeline.py:976-979: ``chunks = torch.chunk(denoised, chunk_num, dim=4);
    cat([chunks[-1]] + chunks + [chunks[0]], dim=4)`` -- the last chunk in front, the first behind --
    as one ring window over the longitude axis."""
    W = denoised.shape[-1]
    size, _, last = _chunk_sizes(W, chunk_num)
    return _ring_window(denoised, W - last, last + W + size)


def crop_wrap_padded(videos: torch.Tensor, chunk_num: int = 16) -> torch.Tensor:
    """:990-991: ``torch.cat(torch.chunk(videos, chunk_num + 2, dim=4)[1:-1], dim=4)``: drop the decoded
    padding chunks again."""
    Wv = videos.shape[-1]
    size, count, last = _chunk_sizes(Wv, chunk_num + 2)
    assert count >= 3, "nothing left between the padding chunks"
    return _ring_window(videos, size, Wv - size - last)


def decode_wrap_padded(denoised: torch.Tensor, decode_fn: Callable, chunk_num: int = 16) -> torch.Tensor:
    """The wrap-padded, frame-by-frame decode of :974-991: pad the latent with its own far ends so the
    VAE sees a seamless longitude wrap, decode every frame on its own (``decode_fn`` on [B, C, 1, H, W']),
    concatenate over frames and cut the padding off."""
    padded = wrap_pad_latent(denoised, chunk_num)
    frames = [decode_fn(padded[:, :, [f]]) for f in range(padded.shape[2])]
    return crop_wrap_padded(torch.cat(frames, dim=2), chunk_num)
