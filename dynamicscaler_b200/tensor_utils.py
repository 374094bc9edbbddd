"""Drop-in for the reference's ``utils/tensor_utils.py``: ``mix_latents_with_mask`` on the
GPU (ds_mix_mask, csrc/elementwise.cu) plus the host image loader."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import lib
from .shift_window_utils import _load_image_tensor_from_path


def load_image_tensor_from_path(image_path: str, height: int, width: int, norm_to_1=True):
    """Reference tensor_utils.py:7-16 (host file I/O, unchanged semantics)."""
    return _load_image_tensor_from_path(image_path, height, width, norm_to_1)


def mix_latents_with_mask(latent_1, latent_to_add, mask, mix_ratio):
    """Reference tensor_utils.py:19-39:
        mixed = latent_1*(1-r) + latent_to_add*r;   out = latent_1*(1-mask) + mixed*mask
    One fused kernel instead of six temporaries; every product/sum is rounded separately
    like the reference's elementwise chain.  mask: [1, h, w] (broadcast over B, C, F) or 5-D."""
    if len(mask.shape) == 3:
        mask_hw = mask.shape[-2] * mask.shape[-1]
        if mask.shape[0] != 1 or tuple(mask.shape[-2:]) != tuple(latent_1.shape[-2:]):
            raise RuntimeError(f"mask shape {tuple(mask.shape)} does not broadcast over latent {tuple(latent_1.shape)}")
    elif len(mask.shape) == 5:
        mask_hw = 0
    else:
        print("noise shape should be [1, H, W] or [B, N, C, H, W]")
        raise NotImplementedError
    _lib.require_cuda(latent_1, "latent_1")
    l1 = latent_1.contiguous()
    add = latent_to_add.to(device=l1.device, dtype=l1.dtype).expand_as(l1).contiguous()
    m = mask.to(device=l1.device, dtype=l1.dtype)
    if mask_hw == 0 and m.shape != l1.shape:
        m = m.expand_as(l1)
    m = m.contiguous()
    out = torch.empty_like(l1)
    r = float(mix_ratio)
    with _lib.on(l1.device):
        lib.check(lib.ds_mix_mask(l1.data_ptr(), add.data_ptr(), m.data_ptr(), out.data_ptr(),
                                  r, 1 - r, l1.numel(), mask_hw, _lib.dtype_code(l1.dtype),
                                  _lib.stream_ptr(l1.device)), "ds_mix_mask")
    return out
