"""Host-side geometry of the tiling path: uv maps, index/wrap maps and merge plans.

The kernels never evaluate the equirect<->perspective trigonometry.  ``floor(u)`` /
``floor(v)`` flip for thousands of view pixels between libm implementations,
so "index maps bit-exact with the reference" is only attainable by evaluating the
reference's own torch-CPU float32 op sequence (utils/panorama_tensor_utils.py:204-245) on
the host.  The maps are static geometry: each (fov, theta, phi, w, h, H, W) is computed
once, cached, and uploaded once per device; a 50-step run touches 440 of them.

Everything here is integer/fp32 bookkeeping on [h*w]-sized arrays; the latents themselves
are only ever touched by the CUDA kernels.
"""
from __future__ import annotations

import threading
from collections import OrderedDict

import numpy as np
import torch

F32 = np.float32
_LOCK = threading.Lock()


class _LRU(OrderedDict):
    def __init__(self, cap):
        super().__init__()
        self.cap = cap

    def get_or(self, key, make):
        with _LOCK:
            if key in self:
                self.move_to_end(key)
                return self[key]
        val = make()
        with _LOCK:
            self[key] = val
            while len(self) > self.cap:
                self.popitem(last=False)
        return val


# --------------------------------------------------------------------------- uv ----
def compute_uv(fov, theta, phi, width, height, H, W):
    """(u, v) float32 [height, width]: the reference's ``_get_uv`` op for op, on the CPU in
    float32 (panorama_tensor_utils.py:204-245; ring_panorama_tensor_utils.py:216-258).
    Angles are NOT reduced modulo 360."""
    dt = torch.float32
    as_t = lambda a: torch.tensor(a, dtype=dt)
    fov_r, th, ph = torch.deg2rad(as_t(fov)), torch.deg2rad(as_t(theta)), torch.deg2rad(as_t(phi))
    focal = 0.5 * width / torch.tan(fov_r / 2)
    x = torch.linspace(-width / 2, width / 2 - 1, steps=width, dtype=dt)
    y = torch.linspace(-height / 2, height / 2 - 1, steps=height, dtype=dt)
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    ray = torch.stack([xx, yy, torch.full_like(xx, focal)], dim=-1)
    ray = ray / torch.norm(ray, dim=-1, keepdim=True)
    cph, sph, cth, sth = torch.cos(ph), torch.sin(ph), torch.cos(th), torch.sin(th)
    r_phi = torch.tensor([[1, 0, 0], [0, cph, -sph], [0, sph, cph]], dtype=dt)
    r_theta = torch.tensor([[cth, 0, sth], [0, 1, 0], [-sth, 0, cth]], dtype=dt)
    rot = torch.matmul(r_theta, r_phi)
    ray = torch.matmul(ray.view(-1, 3), rot.t()).view(height, width, 3)
    lon = torch.atan2(ray[..., 0], ray[..., 2])
    lat = torch.asin(ray[..., 1])
    lon = (lon + 2 * torch.pi) % (2 * torch.pi)
    u = lon / (2 * torch.pi) * (W - 1)
    v = (lat + torch.pi / 2) / torch.pi * (H - 1)
    return u.numpy(), v.numpy()


class ViewMap:
    """All index/weight maps of one view, derived lazily from (u, v)."""

    def __init__(self, key):
        self.key = key
        fov, theta, phi, w, h, H, W = key
        self.w, self.h, self.H, self.W = w, h, H, W
        self.hw = w * h
        self.u, self.v = compute_uv(fov, theta, phi, w, h, H, W)
        self._cache = {}

    def _memo(self, name, make):
        if name not in self._cache:
            self._cache[name] = make()
        return self._cache[name]

    # A.2 nearest read: column wraps, row clamps (panorama_tensor_utils.py:185-193)
    @property
    def idx_read(self):
        def make():
            c = np.floor(self.u).astype(np.int64) % self.W
            r = np.clip(np.floor(self.v).astype(np.int64), 0, self.H - 1)
            return (r * self.W + c).reshape(-1).astype(np.int32)
        return self._memo("idx_read", make)

    @property
    def read_valid(self):
        """[h, w] bool of panorama_tensor_utils.py:198 (vacuously all True)."""
        return self._memo("read_valid", lambda: (self.u >= 0) & (self.u < self.W) & (self.v >= 0) & (self.v < self.H))

    @property
    def all_valid(self):
        """True when the validity mask of :198 is all ones (always, in practice)."""
        return self._memo("all_valid", lambda: bool(self.read_valid.all()))

    # A.3 nearest write: no wrap, validity mask, last writer wins (:166-181)
    @property
    def write_plan(self):
        """(dst[T], src[T]) int32 sorted by dst: panorama pixel <- winning view pixel."""
        def make():
            c = np.floor(self.u).astype(np.int64).reshape(-1)
            r = np.floor(self.v).astype(np.int64).reshape(-1)
            ok = (c >= 0) & (c < self.W) & (r >= 0) & (r < self.H)
            return last_writer_plan((r * self.W + c)[ok], np.nonzero(ok)[0])
        return self._memo("write_plan", make)

    # A.4 round-to-nearest write targets (:87-88)
    @property
    def idx_round(self):
        def make():
            c = np.clip(np.rint(self.u).astype(np.int64), 0, self.W - 1)
            r = np.clip(np.rint(self.v).astype(np.int64), 0, self.H - 1)
            return (r * self.W + c).reshape(-1)
        return self._memo("idx_round", make)

    # A.5 bilinear read: the reference grid through ATen's grid_sample coordinate chain
    def read_taps(self, align_corners=True, mode="bilinear"):
        """(idx4[4, hw] int32, wt4[4, hw] float32), taps nw, ne, sw, se.  Out-of-range taps
        (x1 == W or y1 == H, whose weight is 0) are clamped in bounds with weight 0.
        mode='nearest' yields grid_sample's nearest tap with weight 1 in slot 0."""
        def make():
            W, H = self.W, self.H
            gx = (self.u / F32(W - 1)) * F32(2) - F32(1)          # panorama_tensor_utils.py:38-39
            gy = (self.v / F32(H - 1)) * F32(2) - F32(1)
            if align_corners:                                       # ATen grid_sampler_unnormalize
                ix = (gx + F32(1)) * F32((W - 1) / 2)
                iy = (gy + F32(1)) * F32((H - 1) / 2)
            else:
                ix = (gx + F32(1)) * F32(W / 2) - F32(0.5)
                iy = (gy + F32(1)) * F32(H / 2) - F32(0.5)
            ix = np.minimum(F32(W - 1), np.maximum(ix, F32(0)))     # padding_mode='border'
            iy = np.minimum(F32(H - 1), np.maximum(iy, F32(0)))
            if mode == "nearest":
                xn = np.clip(np.rint(ix).astype(np.int64), 0, W - 1)
                yn = np.clip(np.rint(iy).astype(np.int64), 0, H - 1)
                idx = np.tile((yn * W + xn).reshape(1, -1), (4, 1))
                wt = np.zeros((4, self.hw), F32)
                wt[0] = 1
                return idx.astype(np.int32), wt
            x0, y0 = np.floor(ix), np.floor(iy)
            west = ix - x0
            east = F32(1) - west
            north = iy - y0
            south = F32(1) - north
            wt = np.stack([south * east, south * west, north * east, north * west]).reshape(4, -1)
            x0, y0 = x0.astype(np.int64).reshape(-1), y0.astype(np.int64).reshape(-1)
            xs = np.stack([x0, x0 + 1, x0, x0 + 1])
            ys = np.stack([y0, y0, y0 + 1, y0 + 1])
            inside = (xs >= 0) & (xs < W) & (ys >= 0) & (ys < H)
            wt = np.where(inside, wt, F32(0)).astype(F32)
            idx = np.clip(ys, 0, H - 1) * W + np.clip(xs, 0, W - 1)
            return idx.astype(np.int32), np.ascontiguousarray(wt)
        return self._memo(("read_taps", align_corners, mode), make)

    # A.6 bilinear write taps: longitude wraps, latitude clamps (:110-133)
    @property
    def write_taps(self):
        """(idx[4, hw] int64, wt[4, hw] float32) in the reference's tap order 00,01,10,11."""
        def make():
            W, H = self.W, self.H
            u0 = np.floor(self.u).astype(np.int64)
            v0 = np.floor(self.v).astype(np.int64)
            u1 = (u0 + 1) % W
            v1 = np.clip(v0 + 1, 0, H - 1)
            du = self.u - u0.astype(F32)
            dv = self.v - v0.astype(F32)
            one = F32(1)
            wt = np.stack([(one - du) * (one - dv), (one - du) * dv, du * (one - dv), du * dv])
            idx = np.stack([v0 * W + u0, v1 * W + u0, v0 * W + u1, v1 * W + u1])
            return idx.reshape(4, -1), wt.reshape(4, -1).astype(F32)
        return self._memo("write_taps", make)


_VIEWS = _LRU(4096)


def view_map(fov, theta, phi, width, height, H, W) -> ViewMap:
    key = (float(fov), float(theta), float(phi), int(width), int(height), int(H), int(W))
    return _VIEWS.get_or(key, lambda: ViewMap(key))


# ------------------------------------------------------------------ merge plans ----
def last_writer_plan(targets, sources):
    """Collapse an ordered forward scatter ``pano[targets[i]] = payload[sources[i]]``
    (applied for i ascending) into its net effect: for every distinct target the LAST i
    wins.  Returns (dst, src) int32 sorted by dst."""
    if targets.size == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    rev_t = targets[::-1]
    dst, first = np.unique(rev_t, return_index=True)
    src = sources[::-1][first]
    return dst.astype(np.int32), src.astype(np.int32)


def nearest_merge_plan(vmaps):
    """Plan of "set_view_tensor_no_interpolation for every view in order" (t2v_sphere_
    panorama_pipeline.py:276-291).  src = view_index*hw + k."""
    t, s = [], []
    for n, vm in enumerate(vmaps):
        dst, src = vm.write_plan
        t.append(dst.astype(np.int64))
        s.append(src.astype(np.int64) + n * vm.hw)
    return last_writer_plan(np.concatenate(t), np.concatenate(s))


def _view_lut(n_views, local_views):
    lut = np.full(n_views, -1, np.int64)
    lut[np.asarray(local_views, np.int64)] = np.arange(len(local_views))
    return lut


def filter_nearest_plan(plan, hw, n_views, local_views):
    """Restrict a (dst, src) last-writer plan to the views one rank owns and renumber them
    0..len(local_views)-1 (src = local_view*hw + k).  Winners are decided GLOBALLY first, so the
    ranks' filtered plans have disjoint support and their union is the global plan."""
    dst, src = plan
    lut = _view_lut(n_views, local_views)
    lv = lut[src // hw]
    keep = lv >= 0
    return dst[keep].astype(np.int32), (lv[keep] * hw + src[keep] % hw).astype(np.int32)


def filter_bilinear_plan(plan, hw, n_views, local_views):
    """Same for a CSR weighted plan: keep the rows owned by this rank's views."""
    row_dst, row_ptr, ent_src, ent_w = plan
    if row_dst.size == 0:
        return plan
    row_ptr = row_ptr.astype(np.int64)
    lut = _view_lut(n_views, local_views)
    owner = lut[ent_src[row_ptr[:-1]] // hw]
    keep = owner >= 0
    counts = np.diff(row_ptr)[keep]
    new_ptr = np.concatenate([[0], np.cumsum(counts)])
    ent_index = np.repeat(row_ptr[:-1][keep] - new_ptr[:-1], counts) + np.arange(new_ptr[-1])
    src = ent_src[ent_index].astype(np.int64)
    return (row_dst[keep].astype(np.int32), new_ptr.astype(np.int32),
            (lut[src // hw] * hw + src % hw).astype(np.int32), np.ascontiguousarray(ent_w[ent_index], dtype=F32))


def round_merge_plan(vm, n_images, b):
    """Quirk Q3 of ``set_view_tensor`` (panorama_tensor_utils.py:89-93): image ``b`` of
    ``n_images`` scatters its first hw/n_images view pixels to the targets of view pixels
    [b*per, (b+1)*per)."""
    per = vm.hw // n_images
    return last_writer_plan(vm.idx_round[b * per:(b + 1) * per], np.arange(per, dtype=np.int64))


def bilinear_merge_plan(vmaps):
    """CSR plan of "set_view_tensor_bilinear for every view in order".

    Per view the reference accumulates acc/wsum over the 4 taps and then overwrites the
    panorama where wsum > 0 (panorama_tensor_utils.py:136-152), so across an ordered list
    the LAST view with wsum > 0 at a pixel owns it, and the value is that view's own
    acc/wsum.  Entries of a row keep index_add_'s order: tap-major, then k ascending.
    Returns (row_dst[T], row_ptr[T+1], ent_src[E], ent_w[E])."""
    p_all, key_all, src_all, w_all, n_all = [], [], [], [], []
    for n, vm in enumerate(vmaps):
        idx, wt = vm.write_taps                       # [4, hw]
        hw = vm.hw
        k = np.arange(hw, dtype=np.int64)
        for t in range(4):
            p_all.append(idx[t])
            key_all.append(t * hw + k)                # order inside one view
            src_all.append(n * hw + k)
            w_all.append(wt[t])
            n_all.append(np.full(hw, n, np.int64))
    p, key, src, w, n = map(np.concatenate, (p_all, key_all, src_all, w_all, n_all))
    HW = vmaps[0].H * vmaps[0].W
    owner = np.full(HW, -1, np.int64)
    pos = w > 0
    np.maximum.at(owner, p[pos], n[pos])              # last view with positive weight
    keep = owner[p] == n
    p, key, src, w = p[keep], key[keep], src[keep], w[keep]
    order = np.lexsort((key, p))
    p, src, w = p[order], src[order], w[order]
    row_dst, start = np.unique(p, return_index=True)
    row_ptr = np.concatenate([start, [p.size]])
    return (row_dst.astype(np.int32), row_ptr.astype(np.int32), src.astype(np.int32),
            np.ascontiguousarray(w, dtype=F32))


# ------------------------------------------- pixel-major (scatter-form) merge plans ----
def _view_ptr(view_of_item, n_views):
    """[n_views + 1] offsets of each view's slice in a list sorted by view."""
    return np.searchsorted(view_of_item, np.arange(n_views + 1)).astype(np.int32)


def pm_nearest_plan(vmaps, base=None, n_views=None):
    """``nearest_merge_plan`` regrouped for the scatter-form kernel: the surviving view
    pixels sorted by (view, k).  Returns (view_ptr[n_views + 1], ent_src[T] = view*hw + k,
    ent_dst[T]).  ``base`` / ``n_views``: an already filtered (dst, src) plan and its view count."""
    dst, src = nearest_merge_plan(vmaps) if base is None else base
    n_views = len(vmaps) if n_views is None else n_views
    order = np.argsort(src, kind="stable")
    src, dst = src[order], dst[order]
    return _view_ptr(src // vmaps[0].hw, n_views), src.astype(np.int32), dst.astype(np.int32)


def pm_bilinear_plan(vmaps, base=None, n_views=None):
    """``bilinear_merge_plan`` regrouped by owning view: rows sorted by (view, dst), CSR
    entries kept in index_add_ order inside each row.
    Returns (view_ptr[n_views + 1] over rows, row_dst[T], row_ptr[T + 1], ent_src[E], ent_w[E]).
    ``base`` / ``n_views``: an already filtered CSR plan and its view count."""
    row_dst, row_ptr, ent_src, ent_w = bilinear_merge_plan(vmaps) if base is None else base
    n_views = len(vmaps) if n_views is None else n_views
    hw = vmaps[0].hw
    if row_dst.size == 0:
        return (np.zeros(n_views + 1, np.int32), row_dst, np.zeros(1, np.int32), np.zeros(0, np.int32),
                np.zeros(0, F32))
    row_ptr = row_ptr.astype(np.int64)
    owner = ent_src[row_ptr[:-1]] // hw
    order = np.lexsort((row_dst, owner))
    counts = np.diff(row_ptr)[order]
    new_ptr = np.concatenate([[0], np.cumsum(counts)])
    # entry gather: for new row r, entries row_ptr[order[r]] + [0, counts[r])
    ent_index = np.repeat(row_ptr[:-1][order] - new_ptr[:-1], counts) + np.arange(new_ptr[-1])
    return (_view_ptr(owner[order], n_views), row_dst[order].astype(np.int32), new_ptr.astype(np.int32),
            ent_src[ent_index].astype(np.int32), np.ascontiguousarray(ent_w[ent_index], dtype=F32))


PM_TILE_ROWS, PM_TILE_ENTRIES = 32, 64        # csrc/pm.cu kPmTR, kPmEC


def pm_bilinear_tiled_plan(vmaps, base=None, n_views=None):
    """``pm_bilinear_plan`` cut into tiles for ds_pm_merge_bilinear_tiled: tile t owns the
    padded entries [64t, 64t+64) (ent_k = -1, w = 0 padding; k relative to the tile's view) and
    at most 32 whole rows of ONE view.  Returns None when a row has more than 64 entries
    (tiny panoramas; the untiled kernel handles those), else
    (view_ptr[n_views+1] over tiles, tile[n_tiles, 4] = (first row, view, rows, entries),
     row_dst[T], row_start[T] into the padded arrays, ent_k[64*n_tiles], ent_w[64*n_tiles])."""
    view_ptr_rows, row_dst, row_ptr, ent_src, ent_w = pm_bilinear_plan(vmaps, base, n_views)
    n_views = len(vmaps) if n_views is None else n_views
    hw = vmaps[0].hw
    n_rows = row_dst.size
    counts = np.diff(row_ptr.astype(np.int64))
    if n_rows and counts.max() > PM_TILE_ENTRIES:
        return None
    ce = np.concatenate([[0], np.cumsum(counts)])                  # entries before row r
    tiles, view_ptr = [], [0]
    for v in range(n_views):
        i, end = int(view_ptr_rows[v]), int(view_ptr_rows[v + 1])
        while i < end:
            # furthest row boundary within 64 entries, 32 rows and the view
            j = int(np.searchsorted(ce, ce[i] + PM_TILE_ENTRIES, side="right")) - 1
            j = min(j, i + PM_TILE_ROWS, end)
            tiles.append((i, v, j - i, int(ce[j] - ce[i])))
            i = j
        view_ptr.append(len(tiles))
    n_tiles = len(tiles)
    ent_k = np.full(n_tiles * PM_TILE_ENTRIES, -1, np.int32)
    ent_wp = np.zeros(n_tiles * PM_TILE_ENTRIES, F32)
    row_start = np.zeros(n_rows, np.int32)
    for t, (r0, v, nr, ne) in enumerate(tiles):
        e0, base = int(ce[r0]), t * PM_TILE_ENTRIES
        ent_k[base:base + ne] = ent_src[e0:e0 + ne] - v * hw
        ent_wp[base:base + ne] = ent_w[e0:e0 + ne]
        row_start[r0:r0 + nr] = base + (ce[r0:r0 + nr] - e0)
    tile = np.asarray(tiles, np.int32).reshape(n_tiles, 4)
    return np.asarray(view_ptr, np.int32), tile, row_dst, row_start, ent_k, ent_wp


def frame_table(F, frame_begins, f_len):
    """ftab[f] = s*f_len + fo of the LAST frame window s (in order) containing panorama
    frame f = (frame_begins[s] + fo) mod F, else -1 (shift_window_utils.py:181 composed
    over the ordered window list)."""
    ftab = np.full(F, -1, np.int32)
    for s, fb in enumerate(frame_begins):
        fo = np.arange(f_len)
        ftab[(fb + fo) % F] = s * f_len + fo
    return ftab


# -------------------------------------------------------------- device residency ----
_DEVICE = _LRU(8192)
_HOST = _LRU(1024)


def host_cached(key, make):
    """Build-once cache for host-side plans whose builder may return None (= not applicable)."""
    return _HOST.get_or(key, make)


def on_device(key, device, make_numpy):
    """Upload-once cache of small int32/float32 map arrays, keyed by (key, device)."""
    device = torch.device(device)
    index = device.index
    if index is None and device.type == "cuda":
        index = torch.cuda.current_device()
    dkey = (key, device.type, index)

    def make():
        arrs = make_numpy()
        single = not isinstance(arrs, (tuple, list))
        arrs = [arrs] if single else arrs
        out = [torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=False) for a in arrs]
        return out[0] if single else tuple(out)

    return _DEVICE.get_or(dkey, make)


def clear_caches():
    with _LOCK:
        _VIEWS.clear()
        _DEVICE.clear()
        _HOST.clear()
    from . import panorama_tensor_utils
    panorama_tensor_utils._VIEW_CALL_CACHE.clear()
