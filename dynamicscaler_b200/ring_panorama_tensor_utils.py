"""Drop-in for the reference's ``utils/ring_panorama_tensor_utils.py``: the sphere ops over
a frame-axis ring window ``[frame_begin, frame_end)`` (which may wrap: temporal loop).

The reference materialises the whole H x W frame window on every call (8.4 MB copied to
move 0.66 MB).  Here the frame ring is an index computation inside the
same gather / merge kernels: ``(frame_begin + fo) mod F``.
"""
from __future__ import annotations

from .panorama_tensor_utils import PanoramaTensor, _Block
from .shift_window_utils import RingLatent


class RingLatentProxy:
    """Reference ``RingLatentProxy`` (ring_panorama_tensor_utils.py:316-337): holds the
    [B, N, C, H, W]-ordered panorama inside a ``RingLatent`` in [B, C, N, H, W] order."""

    def __init__(self, init_latent):
        assert init_latent.dim() >= 4
        self.original_shape = init_latent.shape
        B, C, N, H, W = self.original_shape
        self.managed_ring_latent = RingLatent(init_latent.permute(0, 2, 1, 3, 4))

    def get_torch_latent(self):
        return self.managed_ring_latent.torch_latent.permute(0, 2, 1, 3, 4)

    def get_operating_shape(self, frame_begin, frame_end):
        return self.get_window_latent(frame_begin, frame_end).shape

    def get_window_latent(self, frame_begin, frame_end):
        return self.managed_ring_latent.get_window_latent(
            frame_begin=frame_begin, frame_end=frame_end).permute(0, 2, 1, 3, 4)

    def set_window_latent(self, input_latent, frame_begin, frame_end):
        self.managed_ring_latent.set_window_latent(input_latent=input_latent.permute(0, 2, 1, 3, 4),
                                                   frame_begin=frame_begin, frame_end=frame_end)


class RingPanoramaTensor(PanoramaTensor):
    """Reference ``RingPanoramaTensor`` (ring_panorama_tensor_utils.py:8-258).  Shares every
    kernel path with :class:`PanoramaTensor`; only the storage block differs (a frame window
    of the ring latent instead of the whole tensor)."""

    def __init__(self, equirect_tensor):
        assert equirect_tensor.dim() >= 2
        H, W = equirect_tensor.shape[-2], equirect_tensor.shape[-1]
        assert W == 2 * H
        if equirect_tensor.dim() == 2:
            equirect_tensor = equirect_tensor.unsqueeze(0)
        self.C = equirect_tensor.shape[0] if equirect_tensor.dim() == 3 else equirect_tensor.shape[-3]
        self.equirect_tensor_handler = RingLatentProxy(init_latent=equirect_tensor)
        self.H, self.W = H, W
        self.device = equirect_tensor.device
        self.dtype = equirect_tensor.dtype

    @property
    def equirect_tensor(self):
        return self.equirect_tensor_handler.get_torch_latent()

    def _block(self, frame_begin=None, frame_end=None) -> _Block:
        ring = self.equirect_tensor_handler.managed_ring_latent
        store = ring._storage()                                      # [B, C, F, H, W], contiguous
        B, C, F = store.shape[:3]
        fb = 0 if frame_begin is None else int(frame_begin)
        fe = F if frame_end is None else int(frame_end)
        # the window asserts of RingLatent.get_window_latent (shift_window_utils.py:75)
        assert 0 <= fb < fe <= F * 2, f"Invalid frame_begin {fb} and frame_end {fe}"
        return _Block(store, B * C, F, fb, fe - fb, True, (B, fe - fb))

    def _check_write(self, frame_begin, frame_end):
        F = self.equirect_tensor_handler.managed_ring_latent.get_shape()[2]
        fb = 0 if frame_begin is None else frame_begin
        fe = F if frame_end is None else frame_end
        assert fe - fb <= F, f"warp should not occur"               # shift_window_utils.py:147

    def get_view_tensor_interpolate(self, fov, theta, phi, width, height, frame_begin=None, frame_end=None,
                                    interpolate_mode='bilinear', interpolate_align_corners=True):
        """Reference :31-57."""
        return super().get_view_tensor_interpolate(fov, theta, phi, width, height, interpolate_mode,
                                                   interpolate_align_corners, _frames=(frame_begin, frame_end))

    def get_view_tensor_no_interpolate(self, fov, theta, phi, width, height, frame_begin=None, frame_end=None):
        """Reference :59-78."""
        return super().get_view_tensor_no_interpolate(fov, theta, phi, width, height,
                                                      _frames=(frame_begin, frame_end))

    def set_view_tensor(self, view_tensor, fov, theta, phi, frame_begin=None, frame_end=None):
        """Reference :80-105 (quirk Q3 with L = B * window frames)."""
        self._check_write(frame_begin, frame_end)
        super().set_view_tensor(view_tensor, fov, theta, phi, _frames=(frame_begin, frame_end))

    def set_view_tensor_bilinear(self, view_tensor, fov, theta, phi, frame_begin=None, frame_end=None):
        """Reference :107-166."""
        self._check_write(frame_begin, frame_end)
        super().set_view_tensor_bilinear(view_tensor, fov, theta, phi, _frames=(frame_begin, frame_end))

    def set_view_tensor_no_interpolation(self, view_tensor, fov, theta, phi, frame_begin=None, frame_end=None):
        """Reference :170-198."""
        self._check_write(frame_begin, frame_end)
        super().set_view_tensor_no_interpolation(view_tensor, fov, theta, phi, _frames=(frame_begin, frame_end))


class RingPanoramaLatentProxy:
    """Reference ``RingPanoramaLatentProxy`` (ring_panorama_tensor_utils.py:262-314)."""

    def __init__(self, equirect_tensor):
        assert equirect_tensor.dim() >= 4, "输入张量必须至少具有四个维度 [B, C, N, H, W]"
        self.original_shape = equirect_tensor.shape
        B, C, N, H, W = self.original_shape
        self.panorama_tensor = RingPanoramaTensor(equirect_tensor.permute(0, 2, 1, 3, 4))

    def get_view_tensor_interpolate(self, fov, theta, phi, width, height, interpolate_mode='bilinear',
                                    interpolate_align_corners=True, frame_begin=None, frame_end=None):
        view = self.panorama_tensor.get_view_tensor_interpolate(
            fov, theta, phi, width, height, frame_begin=frame_begin, frame_end=frame_end,
            interpolate_mode=interpolate_mode, interpolate_align_corners=interpolate_align_corners)
        return view.permute(0, 2, 1, 3, 4).clone()

    def get_view_tensor_no_interpolate(self, fov, theta, phi, width, height, frame_begin=None, frame_end=None):
        view, mask = self.panorama_tensor.get_view_tensor_no_interpolate(
            fov, theta, phi, width, height, frame_begin=frame_begin, frame_end=frame_end)
        return view.permute(0, 2, 1, 3, 4), mask

    def set_view_tensor(self, view_tensor, fov, theta, phi, frame_begin=None, frame_end=None):
        self.panorama_tensor.set_view_tensor(view_tensor.permute(0, 2, 1, 3, 4), fov, theta, phi,
                                             frame_begin=frame_begin, frame_end=frame_end)

    def set_view_tensor_bilinear(self, view_tensor, fov, theta, phi, frame_begin=None, frame_end=None):
        self.panorama_tensor.set_view_tensor_bilinear(view_tensor.permute(0, 2, 1, 3, 4), fov, theta, phi,
                                                      frame_begin=frame_begin, frame_end=frame_end)

    def get_equirect_tensor(self):
        return self.panorama_tensor.equirect_tensor_handler.get_torch_latent().permute(0, 2, 1, 3, 4)

    def set_view_tensor_no_interpolation(self, view_tensor, fov, theta, phi, frame_begin=None, frame_end=None):
        self.panorama_tensor.set_view_tensor_no_interpolation(view_tensor.permute(0, 2, 1, 3, 4), fov, theta, phi,
                                                              frame_begin=frame_begin, frame_end=frame_end)
