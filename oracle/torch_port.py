"""torch-CPU port of the reference's tiling path, used ONLY as the timed CPU baseline
(``bench.py`` ``cpu_baseline`` / ``--impl reference``) and cross-checked in tests.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Why a second restatement: the reference *is* torch eager code, and its cost on a CPU is
dominated by per-call overheads (``_get_uv`` recomputed on every get/set, advanced-index
kernels, 512 ``index_add_`` calls per bilinear write, full frame-window copies in the ring
twin).  ``panorama_oracle`` (numpy) states the semantics but runs 1.3-6x slower than the
real reference, which would flatter the GPU/CPU ratio.  This port issues the same torch
ops per call as the reference does (file:line cited per function), so its timing stands
in for the reference's own; in the build container it is within a few percent of the
imported reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def uv(fov, theta, phi, width, height, H, W, dtype=torch.float32):
    """panorama_tensor_utils.py:204-245, recomputed on every call like the reference."""
    mk = lambda a: torch.tensor(a, dtype=dtype)
    fov_r, th, ph = torch.deg2rad(mk(fov)), torch.deg2rad(mk(theta)), torch.deg2rad(mk(phi))
    f = 0.5 * width / torch.tan(fov_r / 2)
    x = torch.linspace(-width / 2, width / 2 - 1, steps=width, dtype=dtype)
    y = torch.linspace(-height / 2, height / 2 - 1, steps=height, dtype=dtype)
    yv, xv = torch.meshgrid(y, x, indexing="ij")
    xyz = torch.stack([xv, yv, torch.full_like(xv, f)], dim=-1)
    xyz = xyz / torch.norm(xyz, dim=-1, keepdim=True)
    r_phi = torch.tensor([[1, 0, 0], [0, torch.cos(ph), -torch.sin(ph)], [0, torch.sin(ph), torch.cos(ph)]], dtype=dtype)
    r_th = torch.tensor([[torch.cos(th), 0, torch.sin(th)], [0, 1, 0], [-torch.sin(th), 0, torch.cos(th)]], dtype=dtype)
    rot = torch.matmul(r_th, r_phi)
    d = torch.matmul(xyz.view(-1, 3), rot.t()).view(height, width, 3)
    lon = (torch.atan2(d[..., 0], d[..., 2]) + 2 * torch.pi) % (2 * torch.pi)
    lat = torch.asin(d[..., 1])
    return lon / (2 * torch.pi) * (W - 1), (lat + torch.pi / 2) / torch.pi * (H - 1)


class SpherePano:
    """Holds a [B, C, F, H, W] latent the way PanoramaLatentProxy + PanoramaTensor do
    (permuted [B, F, C, H, W] clone; panorama_tensor_utils.py:20,249-258)."""

    def __init__(self, latent):
        self.t = latent.permute(0, 2, 1, 3, 4).clone()
        self.C, self.H, self.W = self.t.shape[-3], self.t.shape[-2], self.t.shape[-1]

    def latent(self):
        return self.t.permute(0, 2, 1, 3, 4)

    def get_nearest(self, fov, theta, phi, width, height):           # :53-70, 185-202, 268-274
        pano = self.t.view(-1, self.C, self.H, self.W)
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        u0 = torch.floor(u).long() % self.W
        v0 = torch.clamp(torch.floor(v).long(), 0, self.H - 1)
        out = pano[:, :, v0, u0].clone()
        mask = torch.ones_like(u0, dtype=self.t.dtype)
        valid = (u >= 0) & (u < self.W) & (v >= 0) & (v < self.H)
        mask[~valid] = 0
        out[:, :, ~valid] = 0
        return out.view(*self.t.shape[:-3], self.C, height, width).permute(0, 2, 1, 3, 4), mask

    def get_bilinear(self, fov, theta, phi, width, height):          # :28-51, 260-266
        pano = self.t.view(-1, self.C, self.H, self.W)
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        grid = torch.stack(((u / (self.W - 1)) * 2 - 1, (v / (self.H - 1)) * 2 - 1), dim=-1)
        grid = grid.unsqueeze(0).repeat(pano.shape[0], 1, 1, 1)
        out = F.grid_sample(pano, grid, mode="bilinear", padding_mode="border", align_corners=True)
        return out.view(*self.t.shape[:-3], self.C, height, width).permute(0, 2, 1, 3, 4).clone()

    def set_nearest(self, view, fov, theta, phi):                     # :154-183, 288-290
        view = view.permute(0, 2, 1, 3, 4)
        B = view.shape[0] * view.shape[1]
        height, width = view.shape[-2:]
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        ui, vi = torch.floor(u).long(), torch.floor(v).long()
        valid = ((ui >= 0) & (ui < self.W) & (vi >= 0) & (vi < self.H)).view(-1)
        lin = (vi * self.W + ui).view(-1)[valid]
        src = view.reshape(B * self.C, -1)[:, valid]
        flat = self.t.view(-1, self.C, self.H * self.W).reshape(B * self.C, -1)
        flat[:, lin] = src
        self.t = flat.reshape(self.t.shape)

    def set_bilinear(self, view, fov, theta, phi):                    # :98-152, 280-282
        view = view.permute(0, 2, 1, 3, 4)
        B = view.shape[0] * view.shape[1]
        height, width = view.shape[-2:]
        vw = view.reshape(B, self.C, -1).clone()
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        u0, v0 = torch.floor(u).long(), torch.floor(v).long()
        u1, v1 = (u0 + 1) % self.W, torch.clamp(v0 + 1, 0, self.H - 1)
        du, dv = (u - u0.float()).view(-1), (v - v0.float()).view(-1)
        ws = [(1 - du) * (1 - dv), (1 - du) * dv, du * (1 - dv), du * dv]
        ids = [(v0 * self.W + u0).view(-1), (v1 * self.W + u0).view(-1),
               (v0 * self.W + u1).view(-1), (v1 * self.W + u1).view(-1)]
        flat = self.t.view(B, self.C, -1)
        for b in range(B):
            acc = torch.zeros_like(flat[b])
            wsum = torch.zeros_like(flat[b])
            for c in range(self.C):
                for i, w in zip(ids, ws):
                    acc[c].index_add_(0, i, vw[b, c] * w)
                for i, w in zip(ids, ws):
                    wsum[c].index_add_(0, i, w)
            m = wsum > 0
            flat[b][m] = acc[m] / wsum[m]


def sphere_step(lat, den, views, x_prev, x0, fov=120, width=64, height=40, mode="nearest"):
    """gather-all -> merge-all with the reference's per-call op sequence: the composition
    SURVEY §8d times.  lat/den: [1, C, F, H, W] tensors; x_prev/x0: lists of window tensors.
    Returns (windows, latent, denoised, mask)."""
    h_lat, h_den = SpherePano(lat), SpherePano(den)
    h_msk = SpherePano(torch.zeros(1, 1, 1, lat.shape[-2], lat.shape[-1], dtype=lat.dtype))
    if mode == "nearest":
        wins = [h_lat.get_nearest(fov, th, ph, width, height)[0].clone() for (ph, th) in views]
    else:
        wins = [h_lat.get_bilinear(fov, th, ph, width, height) for (ph, th) in views]
    put = "set_nearest" if mode == "nearest" else "set_bilinear"
    ones = torch.ones(1, 1, 1, height, width, dtype=lat.dtype)
    for n, (ph, th) in enumerate(views):
        getattr(h_lat, put)(x_prev[n], fov, th, ph)
        getattr(h_den, put)(x0[n], fov, th, ph)
        getattr(h_msk, put)(ones, fov, th, ph)
    return wins, h_lat.latent(), h_den.latent(), h_msk.latent()[0, 0]


def ring_get(lat, left, top, fb, w, h, f):
    """RingLatent.get_window_latent (shift_window_utils.py:48-114): slice + 3 nested cat + clone."""
    _, _, Ft, H, W = lat.shape

    def runs(b, e, size):
        out, pos = [], b
        while pos < e:
            stop = min(e, (pos // size + 1) * size)
            out.append(slice(pos % size, pos % size + stop - pos))
            pos = stop
        return out

    fr, hr, wr = runs(fb, fb + f, Ft), runs(top, top + h, H), runs(left, left + w, W)
    return torch.cat([torch.cat([torch.cat([lat[:, :, a, b, c] for c in wr], dim=4) for b in hr], dim=3)
                      for a in fr], dim=2).clone()


def ring_set(lat, x, left, top, fb):
    """RingLatent.set_window_latent (shift_window_utils.py:116-206): strided slice assigns."""
    _, _, Ft, H, W = lat.shape
    f, h, w = x.shape[2:]

    def runs(b, e, size):
        out, pos = [], b
        while pos < e:
            stop = min(e, (pos // size + 1) * size)
            out.append((pos % size, stop - pos, pos - b))
            pos = stop
        return out

    for fs, fl, fo in runs(fb, fb + f, Ft):
        for hs, hl, ho in runs(top, top + h, H):
            for ws_, wl, wo in runs(left, left + w, W):
                lat[:, :, fs:fs + fl, hs:hs + hl, ws_:ws_ + wl] = x[:, :, fo:fo + fl, ho:ho + hl, wo:wo + wl]


def ring_step(lat, den, windows, x_prev, x0, w=64, h=40):
    """Planar ring gather-all -> merge-all (t2v_sphere_panorama_pipeline.py:537-632 order)."""
    lat, den = lat.clone(), den.clone()
    msk = torch.zeros_like(lat)
    Ft = lat.shape[2]
    wins = [ring_get(lat, l, t, 0, w, h, Ft) for (l, t) in windows]
    ones = torch.ones(1, lat.shape[1], Ft, h, w, dtype=lat.dtype)
    for n, (l, t) in enumerate(windows):
        ring_set(lat, x_prev[n], l, t, 0)
        ring_set(den, x0[n], l, t, 0)
        ring_set(msk, ones, l, t, 0)
    return wins, lat, den, msk


class FrameRingSpherePano:
    """``RingPanoramaLatentProxy`` + ``RingPanoramaTensor`` + ``RingLatentProxy``
    (ring_panorama_tensor_utils.py:8-337): the panorama lives in a RingLatent in
    [B, C, F, H, W] order and EVERY get / set first materialises the whole H x W frame window
    ``[frame_begin, frame_end)`` (slice + cat + clone, :324-337), works on it like SpherePano
    and -- for sets -- writes the whole window back (B == 1 only, like the reference)."""

    def __init__(self, latent):
        self.lat = latent.clone()                                      # RingLatent.__init__ (:42)
        self.C, self.H, self.W = latent.shape[1], latent.shape[-2], latent.shape[-1]

    def latent(self):
        return self.lat

    def _window(self, fb, fe):                                         # RingLatentProxy.get_window_latent
        return ring_get(self.lat, 0, 0, fb, self.W, self.H, fe - fb).permute(0, 2, 1, 3, 4)

    def _put_window(self, win, fb):                                    # RingLatentProxy.set_window_latent
        ring_set(self.lat, win.permute(0, 2, 1, 3, 4), 0, 0, fb)

    def get_nearest(self, fov, theta, phi, width, height, fb, fe):     # :59-78, 201-214
        op = self._window(fb, fe)
        pano = op.view(-1, self.C, self.H, self.W)
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        u0 = torch.floor(u).long() % self.W
        v0 = torch.clamp(torch.floor(v).long(), 0, self.H - 1)
        out = pano[:, :, v0, u0].clone()
        mask = torch.ones_like(u0, dtype=op.dtype)
        valid = (u >= 0) & (u < self.W) & (v >= 0) & (v < self.H)
        mask[~valid] = 0
        out[:, :, ~valid] = 0
        return out.view(*op.shape[:-3], self.C, height, width).permute(0, 2, 1, 3, 4), mask

    def get_bilinear(self, fov, theta, phi, width, height, fb, fe):    # :31-57
        op = self._window(fb, fe)
        pano = op.view(-1, self.C, self.H, self.W)
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        grid = torch.stack(((u / (self.W - 1)) * 2 - 1, (v / (self.H - 1)) * 2 - 1), dim=-1)
        grid = grid.unsqueeze(0).repeat(pano.shape[0], 1, 1, 1)
        out = F.grid_sample(pano, grid, mode="bilinear", padding_mode="border", align_corners=True)
        return out.view(*op.shape[:-3], self.C, height, width).permute(0, 2, 1, 3, 4).clone()

    def set_nearest(self, view, fov, theta, phi, fb, fe):              # :170-198
        view = view.permute(0, 2, 1, 3, 4)
        B = view.shape[0] * view.shape[1]
        height, width = view.shape[-2:]
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        ui, vi = torch.floor(u).long(), torch.floor(v).long()
        valid = ((ui >= 0) & (ui < self.W) & (vi >= 0) & (vi < self.H)).view(-1)
        lin = (vi * self.W + ui).view(-1)[valid]
        src = view.reshape(B * self.C, -1)[:, valid]
        cur = self._window(fb, fe)
        flat = cur.view(-1, self.C, self.H * self.W).reshape(B * self.C, -1)
        flat[:, lin] = src
        self._put_window(flat.reshape(cur.shape), fb)

    def set_bilinear(self, view, fov, theta, phi, fb, fe):             # :107-166
        view = view.permute(0, 2, 1, 3, 4)
        B = view.shape[0] * view.shape[1]
        height, width = view.shape[-2:]
        vw = view.reshape(B, self.C, -1).clone()
        u, v = uv(fov, theta, phi, width, height, self.H, self.W)
        u0, v0 = torch.floor(u).long(), torch.floor(v).long()
        u1, v1 = (u0 + 1) % self.W, torch.clamp(v0 + 1, 0, self.H - 1)
        du, dv = (u - u0.float()).view(-1), (v - v0.float()).view(-1)
        ws = [(1 - du) * (1 - dv), (1 - du) * dv, du * (1 - dv), du * dv]
        ids = [(v0 * self.W + u0).view(-1), (v1 * self.W + u0).view(-1),
               (v0 * self.W + u1).view(-1), (v1 * self.W + u1).view(-1)]
        for b in range(B):
            cur = self._window(fb, fe)                                 # the whole window, once per frame (:143)
            acc = torch.zeros_like(cur.view(B, self.C, -1)[b])
            wsum = torch.zeros_like(acc)
            for c in range(self.C):
                for i, w in zip(ids, ws):
                    acc[c].index_add_(0, i, vw[b, c] * w)
                for i, w in zip(ids, ws):
                    wsum[c].index_add_(0, i, w)
            m = wsum > 0
            cur.view(B, self.C, -1)[b][m] = acc[m] / wsum[m]
            self._put_window(cur, fb)


def frame_ring_sphere_step(lat, den, frame_begins, frames, views, x_prev, x0, fov=120, width=64, height=40,
                           mode="nearest"):
    """gather-all -> merge-all over (frame window, view) jobs in the reference's loop order
    (i2v_sphere_panorama_pipeline.py:284-476: frame windows outer, views inner), latent-shaped
    mask (:220).  x_prev / x0: lists of window tensors in job order.
    Returns (windows, latent, denoised, mask)."""
    h_lat, h_den, h_msk = FrameRingSpherePano(lat), FrameRingSpherePano(den), FrameRingSpherePano(torch.zeros_like(lat))
    jobs = [(fb, ph, th) for fb in frame_begins for (ph, th) in views]
    if mode == "nearest":
        wins = [h_lat.get_nearest(fov, th, ph, width, height, fb, fb + frames)[0].clone() for (fb, ph, th) in jobs]
    else:
        wins = [h_lat.get_bilinear(fov, th, ph, width, height, fb, fb + frames) for (fb, ph, th) in jobs]
    put = "set_nearest" if mode == "nearest" else "set_bilinear"
    ones = torch.ones(1, lat.shape[1], frames, height, width, dtype=lat.dtype)
    for n, (fb, ph, th) in enumerate(jobs):
        getattr(h_lat, put)(x_prev[n], fov, th, ph, fb, fb + frames)
        getattr(h_den, put)(x0[n], fov, th, ph, fb, fb + frames)
        getattr(h_msk, put)(ones, fov, th, ph, fb, fb + frames)
    return wins, h_lat.latent(), h_den.latent(), h_msk.latent()
