"""Drop-in for the reference's ``utils/diffusion_utils.py``: ``resize_video_latent`` on the GPU
(ds_resize_planes, csrc/resize.cu) plus the two frame-padding helpers."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import lib

_MODES = {"nearest": 0, "bilinear": 1, "bicubic": 2}


def padding_latents_at_front(source_latents, front_padding_num):
    """Reference diffusion_utils.py:4-10: repeat the first frame ``front_padding_num`` times."""
    first = source_latents[:, :, [0]]
    return torch.cat([first] * front_padding_num + [source_latents], dim=2)


def padding_latents_at_end(source_latents, end_padding_num):
    """Reference diffusion_utils.py:13-18: repeat the last frame ``end_padding_num`` times."""
    last = source_latents[:, :, [-1]]
    return torch.cat([source_latents] + [last] * end_padding_num, dim=2)


def resize_video_latent(input_latent, target_height, target_width, mode='bilinear', align_corners=False):
    """Reference diffusion_utils.py:21-33: per-frame ``F.interpolate`` of a [B, C, F, h, w]
    latent to [B, C, F, target_height, target_width].  One kernel; every (b, c, f) plane is
    resized independently, so the reference's permute/view round trip is not needed."""
    if mode not in _MODES:
        raise NotImplementedError(f"mode={mode!r}: kernels exist for 'nearest', 'bilinear' and 'bicubic'")
    if mode == 'nearest':
        align_corners = None
    _lib.require_cuda(input_latent, "input_latent")
    batch, channel, frame, h, w = input_latent.shape
    if batch > 1 and frame > 1 and channel > 1 and input_latent.is_contiguous():
        # the reference's .view(batch * frame, ...) of the permuted tensor (:28) cannot merge B and F
        raise RuntimeError("view size is not compatible with input tensor's size and stride (at least one "
                           "dimension spans across two contiguous subspaces). Use .reshape(...) instead.")
    src = input_latent.contiguous()
    out = torch.empty((batch, channel, frame, int(target_height), int(target_width)), dtype=src.dtype,
                      device=src.device)
    with _lib.on(src.device):
        lib.check(lib.ds_resize_planes(src.data_ptr(), out.data_ptr(), batch * channel * frame, h, w,
                                       int(target_height), int(target_width), _MODES[mode],
                                       1 if align_corners else 0, _lib.dtype_code(src.dtype),
                                       _lib.stream_ptr(src.device)), "ds_resize_planes")
    return out
