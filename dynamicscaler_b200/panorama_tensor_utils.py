"""Drop-in for the reference's ``utils/panorama_tensor_utils.py``: equirectangular <->
perspective view extraction and merge, same classes and signatures.

Every get is one gather kernel and every set one gather-form merge kernel
(csrc/sphere.cu) driven by cached host-built index maps (maps.py) -- instead of the
reference's ~30 tiny ops of ``_get_uv`` per call plus index / index_put_ / 512 index_add_.

Storage model.  The reference keeps whatever strides ``permute + clone`` left behind
(SURVEY §7 "Layout decision").  Here ``PanoramaTensor`` always operates on a dense
[P, F, H*W] block: either its own contiguous ``equirect_tensor`` ([..., C, H, W] => P planes,
F = 1) or, when created by ``PanoramaLatentProxy``, the proxy's [B, C, F, H, W] storage of
which ``equirect_tensor`` is the permuted [B, F, C, H, W] alias.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, maps
from ._lib import lib


# ------------------------------------------------------------------ kernel glue ----
def _launch_gather_nearest(store, out, d_idx, d_jobs, f0, n_jobs, P, F, HW, f_len, hw):
    with _lib.on(store.device):
        lib.check(lib.ds_gather_nearest(store.data_ptr(), out.data_ptr(), d_idx.data_ptr(), _lib.ptr(d_jobs),
                                        f0, n_jobs, P, F, HW, f_len, hw, _lib.dtype_code(store.dtype),
                                        _lib.stream_ptr(store.device)), "ds_gather_nearest")


def _launch_gather_bilinear(store, out, d_idx4, d_wt4, d_jobs, f0, n_jobs, P, F, HW, f_len, hw):
    with _lib.on(store.device):
        lib.check(lib.ds_gather_bilinear(store.data_ptr(), out.data_ptr(), d_idx4.data_ptr(), d_wt4.data_ptr(),
                                         _lib.ptr(d_jobs), f0, n_jobs, P, F, HW, f_len, hw,
                                         _lib.dtype_code(store.dtype), _lib.stream_ptr(store.device)),
                  "ds_gather_bilinear")


def _launch_merge_nearest(panos, payloads, mask, mask_mode, mask_planes, d_dst, d_src, T, d_ftab,
                          vpf, P, F, HW, f_len, hw, vec=1, job_begin=0, job_count=1 << 30):
    ref = panos[0] if panos else mask
    with _lib.on(ref.device):
        lib.check(lib.ds_merge_nearest(_lib.ptr_array(panos), _lib.ptr_array(payloads), len(panos),
                                       _lib.ptr(mask), mask_mode, mask_planes, _lib.ptr(d_dst), d_src.data_ptr(),
                                       T, _lib.ptr(d_ftab), vpf, job_begin, job_count, P, F, HW, f_len, hw, vec,
                                       _lib.dtype_code(ref.dtype), _lib.stream_ptr(ref.device)), "ds_merge_nearest")


def _launch_merge_bilinear(panos, payloads, mask, mask_mode, mask_planes, plan, d_ftab, vpf, P, F, HW, f_len, hw,
                           job_begin=0, job_count=1 << 30):
    row_dst, row_ptr, ent_src, ent_w = plan
    ref = panos[0] if panos else mask
    with _lib.on(ref.device):
        lib.check(lib.ds_merge_bilinear(_lib.ptr_array(panos), _lib.ptr_array(payloads), len(panos),
                                        _lib.ptr(mask), mask_mode, mask_planes, row_dst.data_ptr(),
                                        row_ptr.data_ptr(), row_dst.numel(), ent_src.data_ptr(), ent_w.data_ptr(),
                                        _lib.ptr(d_ftab), vpf, job_begin, job_count, P, F, HW, f_len, hw,
                                        _lib.dtype_code(ref.dtype), _lib.stream_ptr(ref.device)), "ds_merge_bilinear")


def _frame_table(F, f0, f_len, device):
    """Device [F] table of the frame ring for a single window (None when it is the identity)."""
    if f0 == 0 and f_len == F:
        return None
    return maps.on_device(("ftab1", F, f0, f_len), device, lambda: maps.frame_table(F, [f0], f_len))


_VIEW_CALL_CACHE = {}      # (kind, fov, theta, phi, w, h, H, W, device) -> (ViewMap, device arrays)
_ONE = {}


def _ones_view(height, width, dtype, device):
    one = _ONE.get((dtype, device))
    if one is None:
        one = _ONE[(dtype, device)] = torch.ones((1, 1), dtype=dtype, device=device)
    return one.expand(height, width)


class _Block:
    """A dense [P, F, H*W] view of a panorama plus the shape bookkeeping of one call."""
    __slots__ = ("store", "P", "F", "f0", "f_len", "proxied", "lead")

    def __init__(self, store, P, F, f0, f_len, proxied, lead):
        self.store, self.P, self.F, self.f0, self.f_len, self.proxied, self.lead = store, P, F, f0, f_len, proxied, lead


class PanoramaTensor:
    """Reference ``PanoramaTensor`` (panorama_tensor_utils.py:5-245)."""

    def __init__(self, equirect_tensor):
        assert equirect_tensor.dim() >= 2
        H, W = equirect_tensor.shape[-2], equirect_tensor.shape[-1]
        assert W == 2 * H
        if equirect_tensor.dim() == 2:
            equirect_tensor = equirect_tensor.unsqueeze(0)
        self.C = equirect_tensor.shape[0] if equirect_tensor.dim() == 3 else equirect_tensor.shape[-3]
        # clone like the reference (:20); dense memory order is kept, so a proxy's permuted
        # [B,F,C,H,W] input stays an alias-free copy in [B,C,F,H,W] order.
        self.equirect_tensor = equirect_tensor.clone()
        self.H, self.W = H, W
        self.device = equirect_tensor.device
        self.dtype = equirect_tensor.dtype

    # per-call host overhead matters for the unmodified call sites (one launch per call): the
    # storage block and the device maps of a view are looked up once and remembered
    _blk_cache = None
    _proxied = False            # True: equirect_tensor is the [B, F, C, H, W] alias of a proxy's [B, C, F, H, W] storage

    def _view_maps(self, what, fov, theta, phi, width, height, device, make):
        """Device arrays of one view (``what`` names the kind): one dict lookup per call (handlers
        are re-created every step, e.g. the mask panorama, so the cache is shared)."""
        key = (what, fov, theta, phi, width, height, self.H, self.W, device)
        hit = _VIEW_CALL_CACHE.get(key)
        if hit is None:
            vm = maps.view_map(fov, theta, phi, width, height, self.H, self.W)
            hit = _VIEW_CALL_CACHE[key] = (vm, make(vm, device))
            if len(_VIEW_CALL_CACHE) > 8192:
                _VIEW_CALL_CACHE.pop(next(iter(_VIEW_CALL_CACHE)))
        return hit

    # -- storage resolution -------------------------------------------------------
    def _block(self, frame_begin=None, frame_end=None) -> _Block:   # frames: ring twin only
        t = self.equirect_tensor
        hit = self._blk_cache
        if hit is not None and hit[0] is t:
            return hit[1]
        if not t.is_cuda:
            # the I2V pipeline wraps the CPU condition image (i2v_sphere_panorama_pipeline.py:223-224,
            # torch.from_numpy) and samples views from it every window (:367): upload it ONCE, like
            # RingImageTensor._storage; a CUDA device must exist -- there is no CPU path
            if not torch.cuda.is_available():
                _lib.require_cuda(t, "PanoramaTensor.equirect_tensor")
            t = self.equirect_tensor = t.cuda()
            self.device = t.device
        if self._proxied and t.dim() == 5:
            # created by PanoramaLatentProxy: storage is dense [B, C, F, H, W], equirect_tensor its
            # permuted [B, F, C, H, W] alias.  The flag -- not the strides -- says so: with F == 1 or
            # C == 1 the alias is itself contiguous
            store = t.permute(0, 2, 1, 3, 4)
            if not store.is_contiguous():                                 # a caller assigned other strides
                store = store.contiguous()
                t = self.equirect_tensor = store.permute(0, 2, 1, 3, 4)
            B, F, C = t.shape[0], t.shape[1], t.shape[2]
            blk = _Block(store, B * C, F, 0, F, True, (B, F))
        else:
            if not t.is_contiguous():
                t = self.equirect_tensor = t.contiguous()
            lead = tuple(t.shape[:-3]) if t.dim() > 3 else ()
            blk = _Block(t, t.numel() // (self.H * self.W), 1, 0, 1, False, lead)
        self._blk_cache = (t, blk)
        return blk

    def _view_shape(self, blk: _Block, height, width):
        if blk.proxied:
            return (blk.lead[0], self.C, blk.f_len, height, width)       # kernel order [B,C,f,h,w]
        return (*blk.lead, self.C, height, width)

    def _finish_view(self, blk: _Block, out, kernel_order=False):
        if blk.proxied and not kernel_order:
            return out.permute(0, 2, 1, 3, 4)                             # caller order [B,f,C,h,w]
        return out

    def _payload(self, blk: _Block, view_tensor, height, width, kernel_order=False):
        """View tensor -> dense [P, f_len, h*w] block in the panorama's dtype.  ``kernel_order``:
        the proxies hand over their callers' [B, C, f, h, w] tensors as they are."""
        v = view_tensor
        if v.device != blk.store.device or v.dtype != blk.store.dtype:
            v = v.to(device=blk.store.device, dtype=blk.store.dtype)
        if blk.proxied and not kernel_order:
            if v.dim() != 5:
                v = v.reshape(blk.lead[0], blk.f_len, self.C, height, width)
            v = v.permute(0, 2, 1, 3, 4)
        v = v.contiguous()
        assert v.numel() == blk.P * blk.f_len * height * width, \
            f"view tensor {tuple(view_tensor.shape)} does not match panorama planes {blk.P}x{blk.f_len}"
        return v

    # -- reads ----------------------------------------------------------------------
    def get_view_tensor_interpolate(self, fov, theta, phi, width, height,
                                    interpolate_mode='bilinear', interpolate_align_corners=True,
                                    _frames=(None, None)):
        """Reference :28-51 -- F.grid_sample(mode, padding_mode='border')."""
        if interpolate_mode not in ("bilinear", "nearest"):
            raise NotImplementedError(f"interpolate_mode={interpolate_mode!r}: only 'bilinear' and 'nearest' have kernels")
        blk = self._block(*_frames)
        vm = maps.view_map(fov, theta, phi, width, height, self.H, self.W)
        key = ("taps", vm.key, bool(interpolate_align_corners), interpolate_mode)
        d_idx4, d_wt4 = maps.on_device(key, blk.store.device,
                                       lambda: vm.read_taps(bool(interpolate_align_corners), interpolate_mode))
        out = torch.empty(self._view_shape(blk, height, width), dtype=blk.store.dtype, device=blk.store.device)
        _launch_gather_bilinear(blk.store, out, d_idx4, d_wt4, None, blk.f0, 1, blk.P, blk.F,
                                self.H * self.W, blk.f_len, vm.hw)
        return self._finish_view(blk, out)

    def get_view_tensor_no_interpolate(self, fov, theta, phi, width, height, _frames=(None, None),
                                       _kernel_order=False):
        """Reference :53-70 + :185-202 -- nearest read, returns (view, ones[h, w])."""
        blk = self._block(*_frames)
        dev = blk.store.device
        vm, d_idx = self._view_maps("idx_read", fov, theta, phi, width, height, dev,
                                    lambda vm, dev: maps.on_device(("idx_read", vm.key), dev, lambda: vm.idx_read))
        out = torch.empty(self._view_shape(blk, height, width), dtype=blk.store.dtype, device=dev)
        _launch_gather_nearest(blk.store, out, d_idx, None, blk.f0, 1, blk.P, blk.F, self.H * self.W, blk.f_len, vm.hw)
        if vm.all_valid:
            # second return value of the reference (:196-202): ones[h, w].  Every caller discards it
            # (`view, _ = ...`); a fresh torch.ones would cost an allocation + a fill kernel per call, so
            # hand out a broadcast view of one cached element (same shape, dtype, values; in-place
            # writes to it raise instead of corrupting the cache)
            mask = _ones_view(height, width, self.dtype, dev)
        else:                                                             # :198-200 (never taken: u, v in range)
            mask = torch.ones((height, width), dtype=self.dtype, device=dev)
            bad = torch.from_numpy(~vm.read_valid).to(dev)
            mask[bad] = 0
            out[..., bad] = 0
        return self._finish_view(blk, out, _kernel_order), mask

    # -- writes ---------------------------------------------------------------------
    def set_view_tensor_no_interpolation(self, view_tensor, fov, theta, phi, _frames=(None, None),
                                         _kernel_order=False):
        """Reference :154-183 -- forward nearest scatter, last view pixel wins (P4)."""
        if view_tensor.dim() == 3:
            view_tensor = view_tensor.unsqueeze(0)
        height, width = view_tensor.shape[-2], view_tensor.shape[-1]
        blk = self._block(*_frames)
        dev = blk.store.device
        vm, (d_dst, d_src) = self._view_maps(
            "write_plan", fov, theta, phi, width, height, dev,
            lambda vm, dev: maps.on_device(("write_plan", vm.key), dev, lambda: vm.write_plan))
        pay = self._payload(blk, view_tensor, height, width, _kernel_order)
        _launch_merge_nearest([blk.store], [pay], None, _lib.DS_MASK_NONE, 0, d_dst, d_src, d_dst.numel(),
                              _frame_table(blk.F, blk.f0, blk.f_len, dev), 1,
                              blk.P, blk.F, self.H * self.W, blk.f_len, vm.hw)

    def set_view_tensor_bilinear(self, view_tensor, fov, theta, phi, _frames=(None, None)):
        """Reference :98-152 -- weighted scatter-accumulate + per-pixel weight normalisation."""
        if view_tensor.dim() == 3:
            view_tensor = view_tensor.unsqueeze(0)
        height, width = view_tensor.shape[-2], view_tensor.shape[-1]
        blk = self._block(*_frames)
        vm = maps.view_map(fov, theta, phi, width, height, self.H, self.W)
        plan = maps.on_device(("bilinear_plan", vm.key), blk.store.device, lambda: maps.bilinear_merge_plan([vm]))
        pay = self._payload(blk, view_tensor, height, width)
        _launch_merge_bilinear([blk.store], [pay], None, _lib.DS_MASK_NONE, 0, plan,
                               _frame_table(blk.F, blk.f0, blk.f_len, blk.store.device), 1,
                               blk.P, blk.F, self.H * self.W, blk.f_len, vm.hw)

    def set_view_tensor(self, view_tensor, fov, theta, phi, _frames=(None, None)):
        """Reference :72-96 -- round-to-nearest ``scatter_``, INCLUDING quirk Q3: with L =
        prod(leading dims) > 1 images the [h*w] index is split into L chunks, image b
        scatters its first h*w/L view pixels to the targets of chunk b."""
        if view_tensor.dim() == 3:
            view_tensor = view_tensor.unsqueeze(0)
        height, width = view_tensor.shape[-2], view_tensor.shape[-1]
        blk = self._block(*_frames)
        vm = maps.view_map(fov, theta, phi, width, height, self.H, self.W)
        L = blk.P * blk.f_len // self.C
        if vm.hw % L != 0:
            raise RuntimeError(f"shape '[{L}, -1]' is invalid for input of size {vm.hw}")   # the reference's .view error
        pay = self._payload(blk, view_tensor, height, width)
        HW = self.H * self.W
        for b in range(L):
            d_dst, d_src = maps.on_device(("round_plan", vm.key, L, b), blk.store.device,
                                          lambda: maps.round_merge_plan(vm, L, b))
            if blk.proxied:      # image (batch, window frame): planes c of one batch entry, one frame
                bb, fo = divmod(b, blk.f_len)
                pano = blk.store.view(-1, self.C, blk.F, HW)[bb]
                src = pay.view(-1, self.C, blk.f_len, vm.hw)[bb]
                ftab = maps.on_device(("ftab_one", blk.F, (blk.f0 + fo) % blk.F, fo, blk.f_len), blk.store.device,
                                      lambda: _single_frame_table(blk.F, (blk.f0 + fo) % blk.F, fo))
                _launch_merge_nearest([pano], [src], None, _lib.DS_MASK_NONE, 0, d_dst, d_src, d_dst.numel(),
                                      ftab, 1, self.C, blk.F, HW, blk.f_len, vm.hw)
            else:
                pano = blk.store.view(L, self.C, HW)[b]
                src = pay.view(L, self.C, vm.hw)[b]
                _launch_merge_nearest([pano], [src], None, _lib.DS_MASK_NONE, 0, d_dst, d_src, d_dst.numel(),
                                      None, 1, self.C, 1, HW, 1, vm.hw)
        if L == 1 and not blk.proxied:
            self.equirect_tensor = blk.store.view(self.C, self.H, self.W)        # ":96 pano.squeeze(0)"

    # -- kept for API compatibility ---------------------------------------------------
    def _sample_equirect_tensor_nearest(self, pano, u, v):
        """Reference :185-202 on caller-supplied (u, v) tensors: uploads the index map and
        runs the nearest gather kernel on ``pano`` [B, C, H, W]."""
        _lib.require_cuda(pano, "pano")
        u0 = torch.floor(u).long() % self.W
        v0 = torch.clamp(torch.floor(v).long(), 0, self.H - 1)
        d_idx = (v0 * self.W + u0).reshape(1, -1).to(device=pano.device, dtype=torch.int32)
        pano = pano.contiguous()
        B, C = pano.shape[0], pano.shape[1]
        out = torch.empty((B, C, *u.shape), dtype=pano.dtype, device=pano.device)
        _launch_gather_nearest(pano, out, d_idx, None, 0, 1, B * C, 1, self.H * self.W, 1, u.numel())
        mask = torch.ones_like(u0, dtype=self.dtype, device=pano.device)
        valid = ((u >= 0) & (u < self.W) & (v >= 0) & (v < self.H)).to(pano.device)
        mask[~valid] = 0
        out[:, :, ~valid] = 0
        return out, mask

    def _get_uv(self, fov, theta, phi, width, height, focal_length=None):
        """Reference :204-245.  Returns the cached host-computed float32 maps as tensors on
        the panorama's device (``focal_length`` is accepted but, as in every call site, unused)."""
        if focal_length is not None:
            raise NotImplementedError("focal_length override is not supported by the cached map generator")
        vm = maps.view_map(fov, theta, phi, width, height, self.H, self.W)
        dev = self.equirect_tensor.device
        return torch.from_numpy(vm.u).to(dev), torch.from_numpy(vm.v).to(dev)


def _single_frame_table(F, f, fo):
    ftab = np.full(F, -1, np.int32)
    ftab[f] = fo
    return ftab


class PanoramaLatentProxy:
    """Reference ``PanoramaLatentProxy`` (panorama_tensor_utils.py:249-290): adapts the
    pipelines' [B, C, N, H, W] latents.  Storage stays in that order; the inner
    ``PanoramaTensor`` sees it as the permuted [B, N, C, H, W] alias, exactly the shapes the
    reference exposes, with no data movement."""

    def __init__(self, equirect_tensor):
        assert equirect_tensor.dim() >= 4, "输入张量必须至少具有四个维度 [B, C, N, H, W]"
        self.original_shape = equirect_tensor.shape
        B, C, N, H, W = self.original_shape
        self.panorama_tensor = PanoramaTensor(equirect_tensor.permute(0, 2, 1, 3, 4))
        # make the storage dense in [B, C, N, H, W] order whatever strides the input had
        store = self.panorama_tensor.equirect_tensor.permute(0, 2, 1, 3, 4).contiguous()
        self.panorama_tensor.equirect_tensor = store.permute(0, 2, 1, 3, 4)
        self.panorama_tensor._proxied = True

    def get_view_tensor_interpolate(self, fov, theta, phi, width, height,
                                    interpolate_mode='bilinear', interpolate_align_corners=True):
        view = self.panorama_tensor.get_view_tensor_interpolate(
            fov, theta, phi, width, height, interpolate_mode, interpolate_align_corners)
        return view.permute(0, 2, 1, 3, 4).clone()

    def get_view_tensor_no_interpolate(self, fov, theta, phi, width, height):
        # the kernel already produces the caller's [B, C, N, h, w] order
        return self.panorama_tensor.get_view_tensor_no_interpolate(fov, theta, phi, width, height, _kernel_order=True)

    def set_view_tensor(self, view_tensor, fov, theta, phi):
        self.panorama_tensor.set_view_tensor(view_tensor.permute(0, 2, 1, 3, 4), fov, theta, phi)

    def set_view_tensor_bilinear(self, view_tensor, fov, theta, phi):
        self.panorama_tensor.set_view_tensor_bilinear(view_tensor.permute(0, 2, 1, 3, 4), fov, theta, phi)

    def get_equirect_tensor(self):
        return self.panorama_tensor.equirect_tensor.permute(0, 2, 1, 3, 4)

    def set_view_tensor_no_interpolation(self, view_tensor, fov, theta, phi):
        if view_tensor.dim() == 5:
            self.panorama_tensor.set_view_tensor_no_interpolation(view_tensor, fov, theta, phi, _kernel_order=True)
        else:
            self.panorama_tensor.set_view_tensor_no_interpolation(view_tensor.permute(0, 2, 1, 3, 4), fov, theta, phi)
