"""Drop-in for the reference's ``utils/shift_window_utils.py``: wrap-around ("ring")
windows on the panorama latent, same classes / signatures / assertion messages, computed
by ds_ring_gather / ds_ring_scatter (csrc/ring.cu) instead of slice + cat + clone.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import lib


def get_dimension_slices_and_sizes(begin, end, size):
    """Split ``[begin, end)`` into contiguous slices modulo ``size`` (reference
    shift_window_utils.py:14-38).  Host-only helper kept for API compatibility; the
    kernels apply the same wrap rule ``(begin + j) mod size`` arithmetically."""
    slices, sizes = [], []
    pos = begin
    while pos < end:
        start = pos % size
        stop_pos = min(end, (pos // size + 1) * size)
        length = stop_pos - pos
        slices.append(slice(start, start + length))
        sizes.append(length)
        pos = stop_pos
    return slices, sizes


def _resolve(shape, pos_left, pos_right, pos_top, pos_down, frame_begin, frame_end):
    depth, height, width = shape[2], shape[-2], shape[-1]
    pos_left = 0 if pos_left is None else pos_left
    pos_right = width if pos_right is None else pos_right
    pos_top = 0 if pos_top is None else pos_top
    pos_down = height if pos_down is None else pos_down
    frame_begin = 0 if frame_begin is None else frame_begin
    frame_end = depth if frame_end is None else frame_end
    # same checks and messages as shift_window_utils.py:73-75 / 141-143
    assert 0 <= pos_left < pos_right <= width * 2, f"Invalid pos_left {pos_left} and pos_right {pos_right}"
    assert 0 <= pos_top < pos_down <= height * 2, f"Invalid pos_top {pos_top} and pos_down {pos_down}"
    assert 0 <= frame_begin < frame_end <= depth * 2, f"Invalid frame_begin {frame_begin} and frame_end {frame_end}"
    return int(pos_left), int(pos_right), int(pos_top), int(pos_down), int(frame_begin), int(frame_end)


class RingLatent:
    """Reference ``RingLatent`` (shift_window_utils.py:40-206).  ``torch_latent`` is the
    [B, C, F, H, W] panorama; callers read and assign it, so every method re-validates it."""

    def __init__(self, init_latent):
        self.torch_latent = init_latent.clone(memory_format=torch.contiguous_format)
        assert len(self.torch_latent.shape) == 5, f"[RingLatent.__init__] init_latent shape {init_latent.shape} not legal"

    def get_shape(self):
        return self.torch_latent.shape

    def _storage(self):
        lat = self.torch_latent
        _lib.require_cuda(lat, "RingLatent.torch_latent")
        if not lat.is_contiguous():
            lat = self.torch_latent = lat.contiguous()
        return lat

    def get_window_latent(self, pos_left: int = None, pos_right: int = None, pos_top: int = None,
                          pos_down: int = None, frame_begin: int = None, frame_end: int = None):
        l, r, t, d, fb, fe = _resolve(self.get_shape(), pos_left, pos_right, pos_top, pos_down, frame_begin, frame_end)
        lat = self._storage()
        B, C, F, H, W = lat.shape
        out = torch.empty((B, C, fe - fb, d - t, r - l), dtype=lat.dtype, device=lat.device)
        with _lib.on(lat.device):
            lib.check(lib.ds_ring_gather(lat.data_ptr(), out.data_ptr(), None, 1, l, t, fb,
                                         B * C, F, H, W, fe - fb, d - t, r - l,
                                         _lib.dtype_code(lat.dtype), _lib.stream_ptr(lat.device)),
                      "ds_ring_gather")
        return out

    def set_window_latent(self, input_latent: torch.Tensor, pos_left: int = None, pos_right: int = None,
                          pos_top: int = None, pos_down: int = None, frame_begin: int = None,
                          frame_end: int = None):
        l, r, t, d, fb, fe = _resolve(self.get_shape(), pos_left, pos_right, pos_top, pos_down, frame_begin, frame_end)
        lat = self._storage()
        B, C, F, H, W = lat.shape
        assert r - l <= W, f"warp should not occur"                     # :145-147
        assert d - t <= H, f"warp should not occur"
        assert fe - fb <= F, f"warp should not occur"
        target = (fe - fb, d - t, r - l)
        assert tuple(input_latent.shape[2:]) == target, \
            f"Input latent shape {input_latent.shape[2:]} does not match target window shape {target}"
        src = input_latent.to(device=lat.device, dtype=lat.dtype)
        if tuple(src.shape[:2]) != (B, C):
            src = src.expand(B, C, *target)                              # slice-assign broadcasts
        src = src.contiguous()
        with _lib.on(lat.device):
            lib.check(lib.ds_ring_scatter(lat.data_ptr(), src.data_ptr(), l, t, fb, B * C, F, H, W,
                                          fe - fb, d - t, r - l, _lib.dtype_code(lat.dtype),
                                          _lib.stream_ptr(lat.device)), "ds_ring_scatter")


def _load_image_tensor_from_path(image_path, height, width, norm_to_1=True):
    # host file I/O, as reference utils/tensor_utils.py:7-16
    import cv2
    from PIL import Image
    rgb = np.array(Image.open(image_path).convert("RGB"), np.float32)
    rgb = cv2.resize(rgb, (width, height), interpolation=cv2.INTER_LINEAR)
    img = torch.from_numpy(rgb).permute(2, 0, 1)
    return (img / 255. - 0.5) * 2 if norm_to_1 else img


class RingImageTensor:
    """Reference ``RingImageTensor`` (shift_window_utils.py:209-276): ring windows of the
    [3, H, W] condition image.  The image is uploaded once; windows are cut on the GPU."""

    def __init__(self, image_path: str, image_tensor: torch.Tensor = None, height: int = 320, width: int = 512):
        if image_tensor is None:
            self.image_tensor = _load_image_tensor_from_path(image_path, height, width)
        else:
            self.image_tensor = image_tensor
        assert list(self.image_tensor.shape) == [3, height, width], \
            f"[RingImageTensor] image shape {self.image_tensor.shape} does not match {[3, height, width]}"
        self._dev = None

    def get_shape(self):
        return self.image_tensor.shape

    def _storage(self):
        img = self.image_tensor
        if self._dev is None or self._dev[0] is not img:
            dev = img if img.is_cuda else img.cuda()
            self._dev = (img, dev.contiguous())
        return self._dev[1]

    def get_window_tensor(self, pos_left: int, pos_right: int, pos_top: int = None, pos_down: int = None):
        height, width = self.get_shape()[-2], self.get_shape()[-1]
        pos_top = 0 if pos_top is None else pos_top
        pos_down = height if pos_down is None else pos_down
        assert 0 <= pos_left < pos_right <= width * 2, \
            f"[RingImageTensor.get_window_tensor] pos_left {pos_left}, pos_right {pos_right} not legal"
        assert 0 <= pos_top < pos_down <= height * 2, \
            f"[RingImageTensor.get_window_tensor] pos_top {pos_top}, pos_down {pos_down} not legal"
        img = self._storage()
        h, w = int(pos_down - pos_top), int(pos_right - pos_left)
        out = torch.empty((3, h, w), dtype=img.dtype, device=img.device)
        with _lib.on(img.device):
            lib.check(lib.ds_ring_gather(img.data_ptr(), out.data_ptr(), None, 1, int(pos_left), int(pos_top), 0,
                                         3, 1, height, width, 1, h, w, _lib.dtype_code(img.dtype),
                                         _lib.stream_ptr(img.device)), "ds_ring_gather")
        return out

    def get_encoded_image_cond(self, pretrained_t2v, pos_left: int, pos_right: int,
                               pos_top: int = None, pos_down: int = None):
        cond = self.get_window_tensor(pos_left=pos_left, pos_right=pos_right, pos_top=pos_top, pos_down=pos_down)
        cond = cond.to(pretrained_t2v.device).unsqueeze(dim=0)       # get_image_embeds expects [b, c, h, w]
        return pretrained_t2v.get_image_embeds(cond)
