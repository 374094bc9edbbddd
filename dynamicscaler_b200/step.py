"""Batched per-denoising-step driver (additive API).

The reference walks the windows of a step one by one (get -> UNet -> set; window k+1
reads what window k wrote).  The batched order used here -- and by the window-sharded
multi-GPU path BASELINE.json names -- is

    gather_all  : every window of the step is cut from the step-start panorama (1 launch)
    denoise     : the caller's UNet + ddim_step over the stacked windows
    merge_all   : all windows written back "in window order, last writer wins" or
                  "weighted splat + normalise" (1 launch)

which equals the reference's own get_*/set_* functions composed in that order
(oracle.panorama_oracle.sphere_step_batched / ring_step_batched).  ``merge_denoised`` fuses the
denoise epilogue (classifier-free guidance + ddim_step) into the merge.

Panoramas live either in the reference's [B, C, F, H, W] order (any tensor; csrc/sphere.cu,
csrc/ring.cu) or, for the sphere path, in a resident pixel-major ``PanoramaState`` (csrc/pm.cu).

With world_size > 1 the views of every frame window are dealt round-robin over the ranks
(planar ring windows: contiguous blocks) and each rank merges only the pixels its own windows win.
With ``make_state(..., peer=True)`` the replicas sit in symmetric memory and the merge kernels
store the surviving rows straight into every rank's replica over NVLink; otherwise the
disjoint-support partial merges go into zeroed sum / weight buffers, one NCCL all-reduce adds
them and a final select produces the next step-start replica on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, maps
from ._lib import lib
from .panorama_tensor_utils import (_launch_gather_bilinear, _launch_gather_nearest, _launch_merge_bilinear,
                                    _launch_merge_nearest)


def ddim_coefficients(alpha_t, alpha_prev, sigma_t, sqrt_one_minus_alpha_t, guidance_scale=1.0):
    """float32 scalars (sa, s1m, sp, sd, sig, g) of one ``ddim_step`` for the fused denoise epilogue
    (include/dynscaler.h ds_pm_merge_*_ddim), derived with the same float32 torch ops the reference
    applies to its ``torch.full`` tensors (pipeline/scheduler.py:78-86).  Arguments: the schedule
    entries ``ddim_alphas[index]``, ``ddim_alphas_prev[index]``, ``ddim_sigmas[index]``,
    ``ddim_sqrt_one_minus_alphas[index]`` and the classifier-free guidance scale."""
    f = lambda v: torch.full((1,), float(v), dtype=torch.float32)
    a_t, a_prev, sig, s1m = f(alpha_t), f(alpha_prev), f(sigma_t), f(sqrt_one_minus_alpha_t)
    sd = (1. - a_prev - sig ** 2).sqrt()
    return (float(a_t.sqrt()), float(s1m), float(a_prev.sqrt()), float(sd), float(sig), float(f(guidance_scale)))


def renoise_coefficients(alphas_cumprod, ddim_timesteps, step_a: int, step_b: int):
    """float32 scalars (c, s) of ``re_noise(x_a, step_a, step_b)`` -- ``x_b = c*x_a + s*eps`` -- derived
    with the float32 torch ops of the reference (pipeline/scheduler.py:98-110): ``alphas_cumprod`` is the
    scheduler's float32 table, ``ddim_timesteps`` its integer DDIM time steps."""
    ac = torch.as_tensor(alphas_cumprod, dtype=torch.float32)
    a, b = ac[int(ddim_timesteps[step_a])], ac[int(ddim_timesteps[step_b])]
    return float(torch.sqrt(b / a)), float(torch.sqrt(1 - b / a))


def shard_range(n_jobs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of the ordered job list: (begin, count) of ``rank``."""
    base, rem = divmod(n_jobs, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


@dataclass
class StepPlan:
    kind: str                       # "sphere_nearest" | "sphere_bilinear" | "ring"
    n_jobs: int
    vpf: int                        # spatial windows (views) per frame window
    f_len: int
    F: int
    H: int
    W: int
    h: int
    w: int
    d_jobs: Optional[torch.Tensor] = None      # [n_jobs, 2] (map row, frame_begin) / ring: [n_jobs, 3]
    d_idx: Optional[torch.Tensor] = None       # nearest gather maps [vpf, hw]
    d_taps: Optional[tuple] = None             # bilinear gather (idx4, wt4)
    d_dst: Optional[torch.Tensor] = None       # nearest merge plan (None: dense)
    d_src: Optional[torch.Tensor] = None
    T: int = 0
    vec: int = 1
    d_csr: Optional[tuple] = None              # bilinear merge plan
    d_ftab: Optional[torch.Tensor] = None
    d_fbegin: Optional[torch.Tensor] = None    # [n_fwin] frame_begin per frame window (pixel-major merges)
    pm_merge: Optional[tuple] = None           # scatter-form plan of the pixel-major merge (maps.pm_*_plan)
    coverage: float = 0.0                      # touched fraction of H*W (bytes accounting)
    n_local: int = 0                           # jobs of THIS rank (== n_jobs on one GPU)
    job_ids: Optional[List[int]] = None        # their indices in the global ordered job list, local order
    extra: dict = field(default_factory=dict)

    @property
    def hw(self):
        return self.h * self.w

    @property
    def n_fwin(self):
        return self.extra.get("n_fwin", 1)


class PanoramaState:
    """Panoramas of one denoising stage, resident on the device in the pixel-major
    ``pm[H*W][F][P]`` order of csrc/pm.cu (P = B*C = 4).  Created by
    ``PanoramaStepper.make_state`` from reference-order [B, C, F, H, W] tensors and turned
    back by ``export()``; between the two, ``gather_all`` / ``merge_all`` work on it in place.

    ``mask``: optional denoised-mask panorama -- a [1, H, W] plane (T2V,
    t2v_sphere_panorama_pipeline.py:159; kept as is) or a latent-shaped tensor (I2V,
    i2v_sphere_panorama_pipeline.py:220; kept pixel-major)."""

    def __init__(self, stepper, tensors, mask=None, peer=False):
        t0 = tensors[0]
        _lib.require_cuda(t0, "panorama")
        B, C, F, H, W = t0.shape
        assert B * C == 4, "the pixel-major path needs B*C == 4 planes (use reference-order tensors otherwise)"
        assert (F, H, W) == (stepper.F, stepper.H, stepper.W)
        assert len(tensors) <= _lib.DS_MAX_PAYLOADS
        self.stepper, self.shape, self.dtype, self.device = stepper, tuple(t0.shape), t0.dtype, t0.device
        self.mask_plane, self.mask_pm = None, None
        self._mask_shape = None if mask is None else tuple(mask.shape)
        self.peer_table, self._peer_buf, self._peer_hdl = None, None, None
        mask_is_plane = mask is not None and mask.numel() == H * W
        if mask is not None:
            assert mask.dtype == t0.dtype and (mask_is_plane or tuple(mask.shape) == self.shape)
        if peer and stepper.world > 1:
            self._init_peer(tensors, mask, mask_is_plane)
            return
        self.pms = [stepper._pack(t) for t in tensors]
        if mask is not None:
            if mask_is_plane:
                self.mask_plane = mask.contiguous().clone()
            else:
                self.mask_pm = stepper._pack(mask)

    # ------------------------------------------------------------ peer replicas ----
    def _init_peer(self, tensors, mask, mask_is_plane):
        """Allocate [table | panoramas | mask] in ONE symmetric (peer-mapped) buffer, exchange the
        addresses, fill this rank's ds_peer_table_t.  Collective over the stepper's group."""
        import ctypes as C
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        st = self.stepper
        assert st.world <= _lib.DS_MAX_PEERS
        esz = tensors[0].element_size()
        n = tensors[0].numel()
        align = lambda x: (x + 255) // 256 * 256
        tab_bytes = align(lib.ds_peer_table_size())
        assert lib.ds_peer_table_size() == C.sizeof(_lib.PeerTable)
        pm_bytes = align(n * esz)
        mask_bytes = 0 if mask is None else align(mask.numel() * esz)
        total = tab_bytes + len(tensors) * pm_bytes + mask_bytes
        group = st.group if st.group is not None else dist.group.WORLD
        buf = symm.empty(total, dtype=torch.uint8, device=self.device)
        hdl = symm.rendezvous(buf, group.group_name)
        buf.zero_()
        base = [int(p) for p in hdl.buffer_ptrs]
        F, HW = self.shape[2], self.shape[3] * self.shape[4]
        self.pms = []
        for q, t in enumerate(tensors):
            off = tab_bytes + q * pm_bytes
            pm = buf[off:off + n * esz].view(self.dtype).view(HW, F, 4)
            st._pack(t, out=pm)
            self.pms.append(pm)
        m_off = tab_bytes + len(tensors) * pm_bytes
        if mask is not None:
            mview = buf[m_off:m_off + mask.numel() * esz].view(self.dtype)
            if mask_is_plane:
                self.mask_plane = mview.view(mask.shape)
                self.mask_plane.copy_(mask)
            else:
                self.mask_pm = mview.view(HW, F, 4)
                st._pack(mask, out=self.mask_pm)
        tab = _lib.PeerTable()
        tab.rank, tab.world = st.rank, st.world
        for r in range(st.world):
            for q in range(len(tensors)):
                tab.pm[r][q] = base[r] + tab_bytes + q * pm_bytes
            tab.mask[r] = base[r] + m_off if mask is not None else 0
            tab.ctl[r] = base[r]
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if st.use_multicast else 0
        if mc:                                     # NVLS: one multimem.st reaches every replica
            for q in range(len(tensors)):
                tab.mc_pm[q] = mc + tab_bytes + q * pm_bytes
            tab.mc_mask = mc + m_off if mask is not None else 0
        self.multicast = bool(mc)
        raw = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8)
        buf[:raw.numel()].copy_(raw)
        self.peer_table, self._peer_buf, self._peer_hdl = buf.data_ptr(), buf, hdl
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)                  # every table is filled before any kernel signals into it

    def reset_mask(self):
        """Zero the denoised-mask panorama: the reference builds a fresh zero mask every step
        (t2v_sphere_panorama_pipeline.py:173, i2v_...:242).  With peer replicas every rank clears its
        own replica; call it before the step's first gather on every rank."""
        m = self.mask_plane if self.mask_plane is not None else self.mask_pm
        if m is not None:
            if self.peer_table is not None:
                # the peers' previous merges push mask rows into THIS replica: they must have landed before it
                # is cleared (the next gather would wait for them anyway, but the memset would not)
                self.stepper._peer_sync(self, _lib.DS_PEER_WAIT_MERGES, wait=True)
            m.zero_()

    def peer_error(self) -> bool:
        """True if a cross-GPU wait of this state timed out (results are then undefined)."""
        if self.peer_table is None:
            return False
        off = _lib.PeerTable.error.offset
        return bool(self._peer_buf[off:off + 4].view(torch.int32).item())

    def export(self, index=None):
        """Reference-order [B, C, F, H, W] copies of the resident panoramas (all, or one).  With
        peer replicas this is a collective: it first waits until every rank's pushes have landed."""
        if self.peer_table is not None:
            import torch.distributed as dist
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.stepper.group)
            assert not self.peer_error(), "a cross-GPU wait timed out"
        if index is not None:
            return self.stepper._unpack(self.pms[index], self.shape)
        return [self.stepper._unpack(pm, self.shape) for pm in self.pms]

    def export_mask(self):
        if self.mask_plane is not None:
            return self.mask_plane.view(self._mask_shape).clone()
        if self.mask_pm is not None:
            return self.stepper._unpack(self.mask_pm, self.shape)
        return None


class RingPanoramaState:
    """Reference-order [B, C, F, H, W] panoramas of the PLANAR ring path
    (t2v_sphere_panorama_pipeline.py:537-632) replicated on every rank in symmetric (peer-mapped)
    memory: ``merge_all`` then stores every element this rank's windows win straight into every
    replica over NVLink (ds_merge_nearest_peer) instead of zero buffer + NCCL all-reduce + select.
    Collective constructor (one rendezvous per state); ``tensors`` / ``mask`` alias the replica."""

    def __init__(self, stepper, tensors, mask=None):
        import ctypes as C
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        t0 = tensors[0]
        _lib.require_cuda(t0, "panorama")
        assert stepper.world > 1 and stepper.world <= _lib.DS_MAX_PEERS and len(tensors) <= _lib.DS_MAX_PAYLOADS
        self.stepper, self.shape, self.dtype, self.device = stepper, tuple(t0.shape), t0.dtype, t0.device
        esz, n = t0.element_size(), t0.numel()
        align = lambda x: (x + 255) // 256 * 256
        tab_bytes = align(lib.ds_peer_table_size())
        t_bytes = align(n * esz)
        m_bytes = 0 if mask is None else align(mask.numel() * esz)
        group = stepper.group if stepper.group is not None else dist.group.WORLD
        buf = symm.empty(tab_bytes + len(tensors) * t_bytes + m_bytes, dtype=torch.uint8, device=self.device)
        hdl = symm.rendezvous(buf, group.group_name)
        buf.zero_()
        base = [int(p_) for p_ in hdl.buffer_ptrs]
        self.tensors = []
        for q, t in enumerate(tensors):
            assert tuple(t.shape) == self.shape and t.dtype == self.dtype
            off = tab_bytes + q * t_bytes
            v = buf[off:off + n * esz].view(self.dtype).view(self.shape)
            v.copy_(t)
            self.tensors.append(v)
        m_off = tab_bytes + len(tensors) * t_bytes
        self.mask = None
        if mask is not None:
            assert mask.dtype == self.dtype
            self.mask = buf[m_off:m_off + mask.numel() * esz].view(self.dtype).view(mask.shape)
            self.mask.copy_(mask)
        tab = _lib.PeerTable()
        tab.rank, tab.world = stepper.rank, stepper.world
        for r in range(stepper.world):
            for q in range(len(tensors)):
                tab.pm[r][q] = base[r] + tab_bytes + q * t_bytes
            tab.mask[r] = base[r] + m_off if mask is not None else 0
            tab.ctl[r] = base[r]
        raw = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8)
        buf[:raw.numel()].copy_(raw)
        self.peer_table, self._peer_buf, self._peer_hdl = buf.data_ptr(), buf, hdl
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def peer_error(self) -> bool:
        off = _lib.PeerTable.error.offset
        return bool(self._peer_buf[off:off + 4].view(torch.int32).item())

    def export(self):
        """Clones of the replicated panoramas (collective: waits for every rank's stores)."""
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.stepper.group)
        assert not self.peer_error(), "a cross-GPU wait timed out"
        return [t.clone() for t in self.tensors]

    def export_mask(self):
        return None if self.mask is None else self.mask.clone()


class _PlanCache(dict):
    """Bounded plan cache (oldest entry out first).  A whole-step loop needs one plan per distinct window list
    (10-16 per stage); the reference's one-window-at-a-time order builds one per window and step offset, and a
    planar plan holds a dense H*W map (2 MB at 4K scale), so the cache must not grow without bound."""

    def __init__(self, capacity: int = 1024):
        super().__init__()
        self.capacity = int(capacity)

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        while len(self) > self.capacity:
            del self[next(iter(self))]


class PanoramaStepper:
    """Plans and runs gather_all / merge_all for one panorama geometry on one device."""

    def __init__(self, H: int, W: int, F: int, device=None, rank: int = 0, world: int = 1, group=None):
        self.H, self.W, self.F = int(H), int(W), int(F)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rank, self.world, self.group = rank, world, group
        # peer states: NVLS multicast stores win from 4 GPUs up (one store instead of world-1 copies);
        # at 2 GPUs the per-replica TMA row copies are faster (tools/peer_probe.py)
        self.use_multicast = world > 2
        self._plans = _PlanCache()
        self._xbuf = {}

    @property
    def merge_publish_in_kernel(self) -> bool:
        """Exchange ordering protocol: False = host-side (one ds_pm_peer_sync launch after every gather / merge),
        True = deferred (kernels count themselves as issued, the next kernel publishes).  PROCESS-WIDE: the switch
        lives in the library (ds_tune DS_TUNE_PEER_MERGE_PUBLISH) and every stepper of the process reads the one
        copy kept next to the binding, so host and kernels can never disagree about who publishes."""
        return _lib.PEER_DEFERRED_PUBLISH[0]

    @merge_publish_in_kernel.setter
    def merge_publish_in_kernel(self, on: bool):
        lib.check(lib.ds_tune(6, 1 if on else 0), "ds_tune")
        _lib.PEER_DEFERRED_PUBLISH[0] = bool(on)

    # ------------------------------------------------------------------ planning ----
    def _frames(self, frame_begins, frames):
        if frame_begins is None:
            return [0], self.F
        assert frames is not None and 0 < frames <= self.F, "warp should not occur"
        for fb in frame_begins:
            assert 0 <= fb < fb + frames <= 2 * self.F, f"Invalid frame_begin {fb} and frame_end {fb + frames}"
        return [int(fb) for fb in frame_begins], int(frames)

    def _ftab(self, fbs, f_len):
        if len(fbs) == 1 and fbs[0] == 0 and f_len == self.F:
            return None
        return maps.on_device(("ftab", self.F, tuple(fbs), f_len), self.device,
                              lambda: maps.frame_table(self.F, fbs, f_len))

    def plan_sphere(self, views: Sequence[Tuple[float, float]], width: int, height: int, fov=120,
                    mode: str = "nearest", frame_begins=None, frames=None) -> StepPlan:
        """``views``: ordered (phi, theta) list of the step (pipeline loop order); with
        ``frame_begins`` the job order is frame-window-major, like the reference's loops
        (i2v_sphere_panorama_pipeline.py:284-325)."""
        assert self.W == 2 * self.H
        assert mode in ("nearest", "bilinear")
        fbs, f_len = self._frames(frame_begins, frames)
        key = ("sphere", mode, tuple(views), width, height, fov, tuple(fbs), f_len)
        if key in self._plans:
            return self._plans[key]
        vms = [maps.view_map(fov, th, ph, width, height, self.H, self.W) for (ph, th) in views]
        nv = len(vms)
        # Sharding: views are dealt round-robin over the ranks -- rank r owns views
        # r, r + world, ... of every frame window -- which balances both the job count and the number
        # of surviving pixels (with last-writer-wins the late views own most of the panorama).
        # Every plan array below is restricted to the local views and renumbers them 0..nlv-1; the
        # winners themselves are decided globally, so the ranks' plans have disjoint support.
        local_views = list(range(self.rank, nv, self.world))
        nlv = len(local_views)
        plan = StepPlan(kind="sphere_" + mode, n_jobs=nv * len(fbs), vpf=max(nlv, 1), f_len=f_len, F=self.F,
                        H=self.H, W=self.W, h=height, w=width)
        plan.n_local = nlv * len(fbs)
        plan.extra["n_fwin"] = len(fbs)
        plan.job_ids = [s_ * nv + v for s_ in range(len(fbs)) for v in local_views]
        vkey = tuple(vm.key for vm in vms)
        lkey = (vkey, self.rank, self.world)
        hw = width * height
        jobs = np.array([[v, fb] for fb in fbs for v in local_views], dtype=np.int32).reshape(-1, 2)
        plan.d_jobs = maps.on_device(("jobs", nv, tuple(fbs), self.rank, self.world), self.device, lambda: jobs)
        plan.d_ftab = self._ftab(fbs, f_len)
        if plan.d_ftab is not None:
            plan.d_fbegin = maps.on_device(("fbegin", tuple(fbs)), self.device, lambda: np.array(fbs, dtype=np.int32))
        plan.extra["vms"] = vms
        sharded = self.world > 1
        if mode == "nearest":
            plan.d_idx = maps.on_device(("idx_all", vkey), self.device, lambda: np.stack([vm.idx_read for vm in vms]))
            whole = maps.host_cached(("nplan", vkey), lambda: maps.nearest_merge_plan(vms))
            base = maps.filter_nearest_plan(whole, hw, nv, local_views) if sharded else whole
            plan.d_dst, plan.d_src = maps.on_device(("nplan", lkey), self.device, lambda: base)
            plan.T = base[0].size
            plan.extra["host_plan"] = base
            plan.pm_merge = maps.on_device(("pm_nplan", lkey), self.device,
                                           lambda: maps.pm_nearest_plan(vms, base, nlv))
        else:
            plan.d_taps = maps.on_device(
                ("taps_all", vkey), self.device,
                lambda: (np.stack([vm.read_taps()[0] for vm in vms]), np.stack([vm.read_taps()[1] for vm in vms])))
            whole = maps.host_cached(("bplan", vkey), lambda: maps.bilinear_merge_plan(vms))
            base = maps.filter_bilinear_plan(whole, hw, nv, local_views) if sharded else whole
            plan.d_csr = maps.on_device(("bplan", lkey), self.device, lambda: base)
            plan.T = base[0].size
            plan.extra["host_plan"] = base
            plan.pm_merge = maps.on_device(("pm_bplan", lkey), self.device,
                                           lambda: maps.pm_bilinear_plan(vms, base, nlv))
            if f_len == 16:
                host = maps.host_cached(("pm_btiled", lkey), lambda: maps.pm_bilinear_tiled_plan(vms, base, nlv))
                plan.extra["pm_tiled"] = None if host is None else maps.on_device(("pm_btiled", lkey), self.device,
                                                                                    lambda: host)
        plan.coverage = whole[0].size / float(self.H * self.W)
        self._plans[key] = plan
        return plan

    def plan_ring(self, windows: Sequence[Tuple[int, int]], w: int, h: int,
                  frame_begins=None, frames=None) -> StepPlan:
        """``windows``: ordered (left, top) list of the step's planar shift windows
        (t2v_sphere_panorama_pipeline.py:537-560 loop order)."""
        fbs, f_len = self._frames(frame_begins, frames)
        assert w <= self.W and h <= self.H, "warp should not occur"
        for (l, t) in windows:
            assert 0 <= l < l + w <= 2 * self.W, f"Invalid pos_left {l} and pos_right {l + w}"
            assert 0 <= t < t + h <= 2 * self.H, f"Invalid pos_top {t} and pos_down {t + h}"
        key = ("ring", tuple(windows), w, h, tuple(fbs), f_len)
        if key in self._plans:
            return self._plans[key]
        nv = len(windows)
        plan = StepPlan(kind="ring", n_jobs=nv * len(fbs), vpf=nv, f_len=f_len, F=self.F, H=self.H, W=self.W, h=h, w=w)
        jobs3 = np.array([[l, t, fb] for fb in fbs for (l, t) in windows], dtype=np.int32)
        plan.d_jobs = torch.from_numpy(jobs3).to(self.device)
        d_wins2 = torch.from_numpy(np.array(windows, dtype=np.int32).reshape(-1, 2)).to(self.device)
        plan.d_src = torch.empty(self.H * self.W, dtype=torch.int32, device=self.device)
        with _lib.on(self.device):
            lib.check(lib.ds_ring_merge_plan(plan.d_src.data_ptr(), d_wins2.data_ptr(), nv, self.H, self.W, h, w,
                                             _lib.stream_ptr(self.device)), "ds_ring_merge_plan")
        aligned = self.W % 4 == 0 and w % 4 == 0 and all(l % 4 == 0 for l, _ in windows)
        plan.vec = 4 if aligned else 1
        plan.T = self.H * self.W // plan.vec
        plan.d_ftab = self._ftab(fbs, f_len)
        plan.coverage = float((plan.d_src >= 0).float().mean().item())
        plan.extra["n_fwin"] = len(fbs)
        j0, nj = shard_range(plan.n_jobs, self.rank, self.world)     # planar windows: contiguous blocks
        plan.n_local, plan.job_ids = nj, list(range(j0, j0 + nj))
        plan.extra["d_wins2"] = d_wins2
        self._plans[key] = plan
        return plan

    # --------------------------------------------------------------------- gather ----
    def local_jobs(self, plan: StepPlan) -> Tuple[int, int]:
        """(first job, count) of this rank INSIDE the plan's own job numbering.  Ring plans keep the
        global numbering (contiguous blocks); sphere plans are already restricted to the rank's
        views, so the range is (0, n_local).  ``plan.job_ids`` maps local jobs to global indices."""
        if plan.kind == "ring":
            return shard_range(plan.n_jobs, self.rank, self.world)
        return 0, plan.n_local

    # ------------------------------------------------------- pixel-major residency ----
    def make_ring_state(self, tensors: Sequence[torch.Tensor], mask: Optional[torch.Tensor] = None) -> RingPanoramaState:
        """Planar ring path, world_size > 1 (collective): replicate reference-order panoramas in symmetric
        memory so that ``merge_all`` exchanges rows by peer stores (see :class:`RingPanoramaState`)."""
        return RingPanoramaState(self, list(tensors), mask)

    def make_state(self, tensors: Sequence[torch.Tensor], mask: Optional[torch.Tensor] = None,
                   peer: bool = False) -> PanoramaState:
        """Pack reference-order panoramas (latent, denoised, ...) into a resident
        :class:`PanoramaState`; one pack kernel per tensor, once per stage.  ``peer=True`` (with
        world_size > 1; collective) puts the replicas in symmetric memory so that ``merge_all``
        pushes rows straight into every rank's replica instead of all-reducing."""
        return PanoramaState(self, list(tensors), mask, peer)

    def _pack(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, C, F, H, W] (B*C == 4) -> pm [H*W, F, 4]: ds_pm_merge_nearest with the identity plan."""
        B, C, F, H, W = t.shape
        t = t.contiguous()
        pm = torch.empty((H * W, F, 4), dtype=t.dtype, device=t.device) if out is None else out
        with _lib.on(t.device):
            lib.check(lib.ds_pm_merge_nearest(_lib.ptr_array([pm]), _lib.ptr_array([t]), 1, None, _lib.DS_MASK_NONE,
                                              None, None, None, H * W, None, None, 1, 1, 0, 1,
                                              4, F, H * W, F, H * W, None, _lib.dtype_code(t.dtype),
                                              _lib.stream_ptr(t.device)), "ds_pm_merge_nearest(pack)")
        return pm

    def _unpack(self, pm: torch.Tensor, shape) -> torch.Tensor:
        """pm [H*W, F, 4] -> [B, C, F, H, W]: ds_pm_gather with the identity map."""
        B, C, F, H, W = shape
        out = torch.empty(shape, dtype=pm.dtype, device=pm.device)
        with _lib.on(pm.device):
            lib.check(lib.ds_pm_gather(pm.data_ptr(), out.data_ptr(), None, None, 1, None, 0, 1, 4, F, H * W, F,
                                       H * W, None, _lib.dtype_code(pm.dtype), _lib.stream_ptr(pm.device)),
                      "ds_pm_gather(unpack)")
        return out

    def _pm_merge_plan(self, plan: StepPlan):
        return plan.pm_merge                      # built (and uploaded) by plan_sphere

    def _peer_sync(self, st, phase, wait=False):
        """Publish this rank's gather / merge count to the peers (the data kernels wait in-kernel)."""
        with _lib.on(st.device):
            lib.check(lib.ds_pm_peer_sync(st.peer_table, phase, 1 if wait else 0, _lib.stream_ptr(st.device)),
                      "ds_pm_peer_sync")

    def _pm_gather(self, pm, plan, out, j0, nj, peer=None):
        F, HW = plan.F, plan.H * plan.W
        d_jobs = plan.d_jobs[j0:j0 + max(nj, 1)]
        if plan.kind == "sphere_nearest":
            d_idx, d_wt, taps = plan.d_idx, None, 1
        else:
            (d_idx, d_wt), taps = plan.d_taps, 4
        with _lib.on(pm.device):
            lib.check(lib.ds_pm_gather(pm.data_ptr(), out.data_ptr(), d_idx.data_ptr(), _lib.ptr(d_wt), taps,
                                       d_jobs.data_ptr(), 0, nj, 4, F, HW, plan.f_len, plan.hw, peer,
                                       _lib.dtype_code(pm.dtype), _lib.stream_ptr(pm.device)), "ds_pm_gather")

    def _pm_gather_renoise(self, st, plan, out, index, j0, nj, rn):
        import ctypes as C
        assert plan.kind == "sphere_nearest", "the re-noise mix follows the nearest read (the reference's call sites)"
        eps = rn["eps"]
        assert eps.is_contiguous() and eps.dtype == st.dtype and tuple(eps.shape) == tuple(out.shape)
        mplan = rn.get("mask_plan") or plan
        assert mplan.kind == "sphere_nearest" and mplan.hw == plan.hw and mplan.n_local == plan.n_local
        if st.mask_plane is not None:
            assert plan.d_ftab is None, "a [1, H, W] mask panorama has no frame axis"
            mask, mode = st.mask_plane, _lib.DS_MASK_PLANE
            mshape = (nj, plan.h, plan.w)
        else:
            assert st.mask_pm is not None, "the state has no denoised-mask panorama"
            mask, mode = st.mask_pm, _lib.DS_MASK_FULL
            mshape = tuple(out.shape)
        raw = torch.empty_like(out) if rn.get("want_raw") else None
        mview = torch.empty(mshape, dtype=st.dtype, device=st.device) if rn.get("want_mask") else None
        r = float(rn["ratio"])
        h_coef = (C.c_float * 4)(float(rn["c"]), float(rn["s"]), r, 1 - r)
        F, HW = plan.F, plan.H * plan.W
        if nj > 0 or (st.peer_table is not None and self.merge_publish_in_kernel):
            d_jobs = plan.d_jobs[j0:j0 + max(nj, 1)]
            with _lib.on(st.device):
                lib.check(lib.ds_pm_gather_renoise(st.pms[index].data_ptr(), out.data_ptr(), _lib.ptr(raw), _lib.ptr(mview),
                                                   eps.data_ptr(), mask.data_ptr(), mode, plan.d_idx.data_ptr(),
                                                   mplan.d_idx.data_ptr(), h_coef, d_jobs.data_ptr(), 0, nj, 4, F, HW,
                                                   plan.f_len, plan.hw, st.peer_table, _lib.dtype_code(st.dtype),
                                                   _lib.stream_ptr(st.device)), "ds_pm_gather_renoise")
        if st.peer_table is not None and not self.merge_publish_in_kernel:
            self._peer_sync(st, _lib.DS_PEER_AFTER_GATHER)
        return out, raw, mview

    def _pm_merge_into(self, pms, payloads, mask, mask_mode, plan, j0, nj, peer=None):
        """Returns True when a merge kernel was launched (in exchange mode it then publishes the merge count)."""
        if plan.T == 0 or nj == 0:
            return False                           # none of this rank's views survives anywhere
        mp = self._pm_merge_plan(plan)
        F, HW = plan.F, plan.H * plan.W
        ref = pms[0] if pms else mask
        common = (_lib.ptr(plan.d_ftab), _lib.ptr(plan.d_fbegin), plan.n_fwin, plan.vpf, j0, nj,
                  4, F, HW, plan.f_len, plan.hw, peer, _lib.dtype_code(ref.dtype), _lib.stream_ptr(ref.device))
        with _lib.on(ref.device):
            if plan.kind == "sphere_nearest":
                view_ptr, ent_src, ent_dst = mp
                lib.check(lib.ds_pm_merge_nearest(_lib.ptr_array(pms), _lib.ptr_array(payloads), len(pms),
                                                  _lib.ptr(mask), mask_mode, view_ptr.data_ptr(),
                                                  ent_src.data_ptr(), ent_dst.data_ptr(), ent_src.numel(), *common),
                          "ds_pm_merge_nearest")
            elif plan.extra.get("pm_tiled") is not None and len(pms) in (1, 2, 4):
                view_ptr, tile, row_dst, row_start, ent_k, ent_w = plan.extra["pm_tiled"]
                lib.check(lib.ds_pm_merge_bilinear_tiled(_lib.ptr_array(pms), _lib.ptr_array(payloads), len(pms),
                                                         _lib.ptr(mask), mask_mode, view_ptr.data_ptr(),
                                                         tile.data_ptr(), tile.shape[0], row_dst.data_ptr(),
                                                         row_start.data_ptr(), row_dst.numel(), ent_k.data_ptr(),
                                                         ent_w.data_ptr(), *common),
                          "ds_pm_merge_bilinear_tiled")
            else:
                view_ptr, row_dst, row_ptr, ent_src, ent_w = mp
                lib.check(lib.ds_pm_merge_bilinear(_lib.ptr_array(pms), _lib.ptr_array(payloads), len(pms),
                                                   _lib.ptr(mask), mask_mode, view_ptr.data_ptr(),
                                                   row_dst.data_ptr(), row_ptr.data_ptr(), row_dst.numel(),
                                                   ent_src.data_ptr(), ent_w.data_ptr(), *common),
                          "ds_pm_merge_bilinear")
        return True

    # ---------------------------------------------- condition-image windows (§8f-3) ----
    def gather_image_windows(self, image: torch.Tensor, windows: Sequence[Tuple[int, int]], w: int, h: int
                             ) -> torch.Tensor:
        """All planar condition-image windows of a step in ONE launch: ``RingImageTensor.
        get_window_tensor`` (shift_window_utils.py:232-263) for every (left, top) in ``windows``
        (pixel units, may wrap: 0 <= begin < end <= 2*size).  image [3, Hi, Wi] -> [n, 3, h, w],
        ready to be fed to the CLIP image embedder as one batch
        (i2v_sphere_panorama_pipeline.py:889-893 calls it once per window)."""
        _lib.require_cuda(image, "image")
        assert image.dim() == 3 and image.is_contiguous()
        C, Hi, Wi = image.shape
        for (l, t) in windows:
            assert 0 <= l < l + w <= Wi * 2, f"[RingImageTensor.get_window_tensor] pos_left {l}, pos_right {l + w} not legal"
            assert 0 <= t < t + h <= Hi * 2, f"[RingImageTensor.get_window_tensor] pos_top {t}, pos_down {t + h} not legal"
        n = len(windows)
        d_wins = torch.tensor([[l, t, 0] for (l, t) in windows], dtype=torch.int32, device=image.device)
        out = torch.empty((n, C, h, w), dtype=image.dtype, device=image.device)
        with _lib.on(image.device):
            lib.check(lib.ds_ring_gather(image.data_ptr(), out.data_ptr(), d_wins.data_ptr(), n, 0, 0, 0,
                                         C, 1, Hi, Wi, 1, h, w, _lib.dtype_code(image.dtype),
                                         _lib.stream_ptr(image.device)), "ds_ring_gather(image windows)")
        return out

    def gather_image_views(self, image: torch.Tensor, views: Sequence[Tuple[float, float]], width: int, height: int,
                           fov=120) -> torch.Tensor:
        """All sphere condition-image views of a step in ONE launch: ``PanoramaTensor(image).
        get_view_tensor_no_interpolate`` (panorama_tensor_utils.py:53-70) for every (phi, theta)
        (i2v_sphere_panorama_pipeline.py:367-372 calls it once per view on the [3, 1024, 2048]
        image).  image [C, Hi, Wi] with Wi == 2*Hi -> [n, C, height, width]."""
        _lib.require_cuda(image, "image")
        assert image.dim() == 3 and image.is_contiguous()
        C, Hi, Wi = image.shape
        assert Wi == 2 * Hi
        vms = [maps.view_map(fov, th, ph, width, height, Hi, Wi) for (ph, th) in views]
        d_idx = maps.on_device(("idx_all", tuple(vm.key for vm in vms)), image.device,
                               lambda: np.stack([vm.idx_read for vm in vms]))
        out = torch.empty((len(vms), C, height, width), dtype=image.dtype, device=image.device)
        _launch_gather_nearest(image, out, d_idx, None, 0, len(vms), C, 1, Hi * Wi, 1, width * height)
        return out

    def gather_all(self, latent, plan: StepPlan, out: Optional[torch.Tensor] = None, index: int = 0,
                   renoise: Optional[dict] = None):
        """latent [B, C, F, H, W] -- or a :class:`PanoramaState` (panorama ``index``) -- ->
        windows [n_local_jobs, B, C, f_len, h, w] (this rank's slice of the ordered job list).

        ``renoise`` (states, nearest plans): fuse what follows the read in the
        reference's loops (t2v_sphere_panorama_pipeline.py:194-221) -- the nearest read of the state's
        denoised-mask panorama, ``re_noise`` with INJECTED epsilon and ``mix_latents_with_mask`` -- into
        the gather: ``dict(eps=<windows-shaped tensor>, c=, s=, ratio=, want_raw=False, want_mask=False,
        mask_plan=None)`` with (c, s) from :func:`renoise_coefficients`.  Returns ``(windows, raw,
        mask_view)``: the mixed windows, the unmixed ones (``want_raw``) and the mask view
        (``want_mask``; [n, h, w] for a [1, H, W] mask, windows-shaped for a latent-shaped mask) -- the
        last two feed ``merge_denoised(prev=, prev_mask=)``.  Planar ring plans on reference-order tensors take the
        same dict plus ``mask=<latent-shaped mask panorama>`` (a ``RingPanoramaState`` uses its own) and
        ``mask_frame0=True`` for the T2V planar loop's ``mask[0, 0, [0]]`` broadcast; ``eps`` None
        skips the re-noise mix and only reads the mask windows (-> ``merge_denoised_ref(prev=, prev_mask=)``)."""
        if isinstance(latent, PanoramaState):
            assert plan.kind != "ring", "planar ring windows run on reference-order tensors"
            st = latent
            B, C = st.shape[0], st.shape[1]
            j0, nj = self.local_jobs(plan)
            shape = (nj, B, C, plan.f_len, plan.h, plan.w)
            if out is None:
                out = torch.empty(shape, dtype=st.dtype, device=st.device)
            assert tuple(out.shape) == shape and out.is_contiguous()
            if renoise is not None:
                return self._pm_gather_renoise(st, plan, out, index, j0, nj, renoise)
            launched = nj > 0 or (st.peer_table is not None and self.merge_publish_in_kernel)
            if launched:                             # (in-kernel publication: an idle rank launches too)
                self._pm_gather(st.pms[index], plan, out, j0, nj, st.peer_table)
            if st.peer_table is not None and not self.merge_publish_in_kernel:
                self._peer_sync(st, _lib.DS_PEER_AFTER_GATHER)   # replicas may be overwritten once every rank has read
            return out
        ring_state = latent if isinstance(latent, RingPanoramaState) else None
        if ring_state is not None:
            assert plan.kind == "ring", "RingPanoramaState holds the planar ring path"
            latent = ring_state.tensors[index]
        _lib.require_cuda(latent, "latent")
        assert latent.is_contiguous() and latent.dim() == 5
        B, C, F, H, W = latent.shape
        assert (F, H, W) == (plan.F, plan.H, plan.W)
        j0, nj = self.local_jobs(plan)
        shape = (nj, B, C, plan.f_len, plan.h, plan.w)
        if out is None:
            out = torch.empty(shape, dtype=latent.dtype, device=latent.device)
        assert tuple(out.shape) == shape and out.is_contiguous()
        assert renoise is None or plan.kind == "ring", "fused re-noise on reference-order tensors: planar ring plans only"
        if nj == 0 and ring_state is None:
            if renoise is None:
                return out
            return (out, out if renoise.get("want_raw") else None, torch.empty_like(out) if renoise.get("want_mask") else None)
        if plan.kind == "ring":
            d_jobs = plan.d_jobs[j0:j0 + nj]

            def ring_gather(src, dst):
                if nj > 0:
                    with _lib.on(src.device):
                        lib.check(lib.ds_ring_gather(src.data_ptr(), dst.data_ptr(), d_jobs.data_ptr(), nj, 0, 0, 0,
                                                     B * C, F, H, W, plan.f_len, plan.h, plan.w,
                                                     _lib.dtype_code(src.dtype), _lib.stream_ptr(src.device)),
                                  "ds_ring_gather")

            ring_gather(latent, out)
            mview = None
            if renoise is not None:
                # planar I2V read side (i2v_sphere_panorama_pipeline.py:859-884): the latent-shaped denoised-mask
                # panorama is read through the same windows, then re_noise + mask mix run as one pass
                mask = ring_state.mask if ring_state is not None else renoise.get("mask")
                assert mask is not None and tuple(mask.shape) == tuple(latent.shape) and mask.dtype == latent.dtype \
                    and mask.is_contiguous(), "the planar loops keep a latent-shaped denoised-mask panorama"
                mview = torch.empty_like(out)
                ring_gather(mask, mview)
            if ring_state is not None:           # replicas may be overwritten once EVERY rank has read (full barrier)
                self._peer_sync(ring_state, _lib.DS_PEER_AFTER_GATHER, wait=True)
            if renoise is not None:
                mixed = out
                eps = renoise.get("eps")
                if eps is not None and nj > 0:
                    assert eps.shape == out.shape and eps.dtype == out.dtype and eps.is_contiguous()
                    r = float(renoise["ratio"])
                    mixed = torch.empty_like(out)
                    plane = plan.hw if renoise.get("mask_frame0") else 0     # t2v_sphere_...:553 mask[0, 0, [0]]
                    assert not plane or B == 1, "the plane-mask variant indexes batch 0"
                    with _lib.on(out.device):
                        lib.check(lib.ds_renoise_mix(out.data_ptr(), eps.data_ptr(), mview.data_ptr(), mixed.data_ptr(),
                                                     float(renoise["c"]), float(renoise["s"]), r, 1 - r, out.numel(),
                                                     out.numel() // nj, plane, _lib.dtype_code(out.dtype),
                                                     _lib.stream_ptr(out.device)), "ds_renoise_mix")
                return (mixed, out if renoise.get("want_raw") else None, mview if renoise.get("want_mask") else None)
        elif plan.kind == "sphere_nearest":
            _launch_gather_nearest(latent, out, plan.d_idx, plan.d_jobs[j0:j0 + nj], 0, nj, B * C, F, H * W,
                                   plan.f_len, plan.hw)
        else:
            _launch_gather_bilinear(latent, out, plan.d_taps[0], plan.d_taps[1], plan.d_jobs[j0:j0 + nj], 0, nj,
                                    B * C, F, H * W, plan.f_len, plan.hw)
        return out

    # ---------------------------------------------------------------------- merge ----
    def _merge_into(self, panos, payloads, mask, mask_mode, mask_planes, plan, P, j0, nj):
        if plan.T == 0:
            return                                 # none of this rank's views survives anywhere
        HW = plan.H * plan.W
        if plan.kind == "sphere_bilinear":
            _launch_merge_bilinear(panos, payloads, mask, mask_mode, mask_planes, plan.d_csr, plan.d_ftab, plan.vpf,
                                   P, plan.F, HW, plan.f_len, plan.hw, job_begin=j0, job_count=nj)
        else:
            _launch_merge_nearest(panos, payloads, mask, mask_mode, mask_planes, plan.d_dst, plan.d_src, plan.T,
                                  plan.d_ftab, plan.vpf, P, plan.F, HW, plan.f_len, plan.hw, vec=plan.vec,
                                  job_begin=j0, job_count=nj)

    def merge_all(self, panos: List[torch.Tensor], payloads: List[torch.Tensor], plan: StepPlan,
                  mask: Optional[torch.Tensor] = None):
        """In-place merge of this step's payloads into the panoramas.

        panos[q]: [B, C, F, H, W]; payloads[q]: [n_local_jobs, B, C, f_len, h, w] (x_prev ->
        latent panorama, pred_x0 -> denoised panorama, ...).  ``mask``: optional [1, H, W]
        (T2V, t2v_sphere_panorama_pipeline.py:159) or [B, C, F, H, W] (I2V, i2v_sphere_
        panorama_pipeline.py:220) denoised-mask panorama that receives 1 where written.
        With world_size > 1 every rank ends up with the identical merged replica."""
        if isinstance(panos, PanoramaState):
            return self._merge_state(panos, payloads, plan)
        if isinstance(panos, RingPanoramaState):
            return self._merge_ring_state(panos, payloads, plan)
        assert len(panos) == len(payloads) and 0 < len(panos) <= _lib.DS_MAX_PAYLOADS
        B, C, F, H, W = panos[0].shape
        P = B * C
        for p_, y in zip(panos, payloads):
            _lib.require_cuda(p_, "panorama")
            assert p_.is_contiguous() and y.is_contiguous() and p_.shape == panos[0].shape and y.dtype == p_.dtype
        j0, nj = self.local_jobs(plan)
        mask_mode, mask_planes = _lib.DS_MASK_NONE, 0
        if mask is not None:
            assert mask.is_contiguous() and mask.dtype == panos[0].dtype
            if mask.numel() == H * W:
                assert plan.d_ftab is None, "a [1, H, W] mask panorama has no frame axis"
                mask_mode = _lib.DS_MASK_PLANE
            else:
                assert mask.numel() % (F * H * W) == 0
                mask_mode, mask_planes = _lib.DS_MASK_FULL, mask.numel() // (F * H * W)
        if self.world == 1:
            self._merge_into(panos, payloads, mask, mask_mode, mask_planes, plan, P, 0, plan.n_local)
            return
        self._merge_sharded(panos, payloads, mask, plan, P, j0, nj)

    def merge_denoised(self, st: PanoramaState, x: torch.Tensor, eps_cond: torch.Tensor,
                       eps_uncond: Optional[torch.Tensor], coef, plan: StepPlan,
                       noise: Optional[torch.Tensor] = None, prev: Optional[torch.Tensor] = None,
                       prev_mask: Optional[torch.Tensor] = None, prev_ratio: Optional[float] = None):
        """merge_all fused with the denoise epilogue: the latent panorama (state
        panorama 0) receives ``x_prev`` and the denoised panorama (1) ``pred_x0`` of
        ``ddim_step(x, eps_uncond + g*(eps_cond - eps_uncond))`` -- computed inside the merge kernel
        for the surviving pixels only, so the ~75 % of window pixels that lose to a later window are
        never denoised, and x_prev / pred_x0 are never materialised.  ``x``: this rank's windows as
        returned by ``gather_all``; ``eps_*``: UNet outputs of the same shape; ``coef``:
        ``ddim_coefficients(...)``; ``noise``: the injected epsilon when sigma > 0.
        ``prev`` / ``prev_mask`` / ``prev_ratio``: the merge prologue of the I2V loops
        (i2v_sphere_panorama_pipeline.py:430-435): ``x_prev = mix_latents_with_mask(x_prev, prev,
        prev_mask, prev_ratio)`` before the latent set, with ``prev`` the windows as gathered and
        ``prev_mask`` the mask view read at gather time (``gather_all(..., renoise=dict(want_raw=True,
        want_mask=True))`` returns both)."""
        import ctypes as C
        assert plan.kind != "ring" and len(st.pms) >= 2, "needs latent + denoised panoramas"
        assert self.world == 1 or st.peer_table is not None, "sharded fused merge needs peer replicas (make_state(peer=True))"
        j0, nj = self.local_jobs(plan)
        for t in (x, eps_cond, eps_uncond, noise):
            assert t is None or (t.is_contiguous() and t.dtype == st.dtype and t.shape == x.shape)
        assert x.shape[0] == nj
        mask, mask_mode = None, _lib.DS_MASK_NONE
        if st.mask_plane is not None:
            assert plan.d_ftab is None, "a [1, H, W] mask panorama has no frame axis"
            mask, mask_mode = st.mask_plane, _lib.DS_MASK_PLANE
        elif st.mask_pm is not None:
            mask, mask_mode = st.mask_pm, _lib.DS_MASK_FULL
        self._pm_merge_plan(plan)
        mview_full = 0
        if prev is not None:
            assert prev_mask is not None and prev_ratio is not None
            assert prev.is_contiguous() and prev.dtype == st.dtype and prev.shape == x.shape
            assert prev_mask.is_contiguous() and prev_mask.dtype == st.dtype
            mview_full = 1 if prev_mask.numel() == x.numel() else 0
            assert mview_full or prev_mask.numel() == nj * plan.hw, "mask view: [n, h, w] or windows-shaped"
        pr = 0.0 if prev_ratio is None else float(prev_ratio)
        h_coef = (C.c_float * 8)(*coef, pr, 1 - pr)
        F, HW = plan.F, plan.H * plan.W
        pms = st.pms[:2]
        head = (_lib.ptr_array(pms), x.data_ptr(), eps_cond.data_ptr(), _lib.ptr(eps_uncond), _lib.ptr(noise),
                _lib.ptr(prev), _lib.ptr(prev_mask if prev is not None else None), mview_full, h_coef,
                _lib.ptr(mask), mask_mode)
        tail = (_lib.ptr(plan.d_ftab), _lib.ptr(plan.d_fbegin), plan.n_fwin, plan.vpf, j0, nj, 4, F, HW, plan.f_len,
                plan.hw, st.peer_table, _lib.dtype_code(st.dtype), _lib.stream_ptr(st.device))
        launched = nj > 0 and plan.T > 0
        with _lib.on(st.device):
            if nj > 0 and plan.T > 0 and plan.kind == "sphere_nearest":
                view_ptr, ent_src, ent_dst = plan.pm_merge
                lib.check(lib.ds_pm_merge_nearest_ddim(*head, view_ptr.data_ptr(), ent_src.data_ptr(), ent_dst.data_ptr(),
                                                       ent_src.numel(), *tail), "ds_pm_merge_nearest_ddim")
            elif nj > 0 and plan.T > 0:
                tiled = plan.extra.get("pm_tiled")
                assert tiled is not None, "the fused weighted merge needs 16-frame windows and rows of <= 64 entries"
                view_ptr, tile, row_dst, row_start, ent_k, ent_w = tiled
                lib.check(lib.ds_pm_merge_bilinear_tiled_ddim(*head, view_ptr.data_ptr(), tile.data_ptr(), tile.shape[0],
                                                              row_dst.data_ptr(), row_start.data_ptr(), row_dst.numel(),
                                                              ent_k.data_ptr(), ent_w.data_ptr(), *tail),
                          "ds_pm_merge_bilinear_tiled_ddim")
        if st.peer_table is not None and not (launched and self.merge_publish_in_kernel):
            self._peer_sync(st, _lib.DS_PEER_AFTER_MERGE)

    def merge_denoised_ref(self, target, x: torch.Tensor, eps_cond: torch.Tensor, eps_uncond: Optional[torch.Tensor],
                           coef, plan: StepPlan, noise: Optional[torch.Tensor] = None,
                           mask: Optional[torch.Tensor] = None, prev: Optional[torch.Tensor] = None,
                           prev_mask: Optional[torch.Tensor] = None, prev_ratio: Optional[float] = None):
        """:meth:`merge_denoised` for REFERENCE-ORDER panoramas and nearest plans (planar ring windows,
        ``plan_ring``; also ``plan_sphere(mode="nearest")``) -- the write side of the planar shift-window
        loop (pipeline/t2v_normal_pipeline.py:526-557): CFG + ``ddim_step`` run inside the merge kernel
        (ds_merge_nearest_ddim) for the winning elements only.  ``target``: ``[latent, denoised]`` tensors
        [B, C, F, H, W] (world_size > 1: every rank ends with the identical merged replica, all-reduce
        exchange) or a :class:`RingPanoramaState` (peer-store exchange).  ``mask`` as in ``merge_all``.
        ``prev`` / ``prev_mask`` / ``prev_ratio``: the ``merge_prev_denoised_ratio`` mix of the planar I2V loop
        (pipeline/i2v_sphere_panorama_pipeline.py:938-943) between ``ddim_step`` and the latent set, with the
        windows and mask windows as read at gather time (``gather_all(..., renoise=dict(want_raw=True,
        want_mask=True, ...))``), both shaped like ``x``."""
        import ctypes as C
        assert plan.kind in ("ring", "sphere_nearest"), "the reference-order fused merge takes nearest plans"
        j0, nj = self.local_jobs(plan)
        if prev is not None:
            assert prev_mask is not None and prev_ratio is not None, "prev needs prev_mask and prev_ratio"
        for t in (x, eps_cond, eps_uncond, noise, prev, prev_mask if prev is not None else None):
            assert t is None or (t.is_contiguous() and t.dtype == x.dtype and t.shape == x.shape)
        assert x.shape[0] == nj
        pr = 0.0 if prev_ratio is None else float(prev_ratio)
        h_coef = (C.c_float * 8)(*coef, pr, 1 - pr)
        ring_state = target if isinstance(target, RingPanoramaState) else None
        if ring_state is not None:
            assert mask is None, "a RingPanoramaState carries its own mask"
            panos, mask = ring_state.tensors[:2], ring_state.mask
        else:
            panos = list(target)
        assert len(panos) == 2, "needs the latent and the denoised panorama"
        B, Cc, F, H, W = panos[0].shape
        P = B * Cc
        for p_ in panos:
            _lib.require_cuda(p_, "panorama")
            assert p_.is_contiguous() and p_.shape == panos[0].shape and p_.dtype == x.dtype
        assert (F, H, W) == (plan.F, plan.H, plan.W)
        mask_mode, mask_planes = _lib.DS_MASK_NONE, 0
        if mask is not None:
            assert mask.is_contiguous() and mask.dtype == x.dtype
            if mask.numel() == H * W:
                assert plan.d_ftab is None, "a [1, H, W] mask panorama has no frame axis"
                mask_mode = _lib.DS_MASK_PLANE
            else:
                assert mask.numel() % (F * H * W) == 0
                mask_mode, mask_planes = _lib.DS_MASK_FULL, mask.numel() // (F * H * W)

        def launch(dst_panos, m, m_mode, m_planes, peer):
            if nj == 0 or plan.T == 0:
                return
            with _lib.on(x.device):
                lib.check(lib.ds_merge_nearest_ddim(_lib.ptr_array(dst_panos), x.data_ptr(), eps_cond.data_ptr(),
                                                    _lib.ptr(eps_uncond), _lib.ptr(noise), _lib.ptr(prev),
                                                    _lib.ptr(prev_mask if prev is not None else None), h_coef,
                                                    _lib.ptr(m), m_mode,
                                                    m_planes, _lib.ptr(plan.d_dst), plan.d_src.data_ptr(), plan.T,
                                                    _lib.ptr(plan.d_ftab), plan.vpf, j0, nj, P, F, H * W, plan.f_len,
                                                    plan.hw, plan.vec, peer, _lib.dtype_code(x.dtype),
                                                    _lib.stream_ptr(x.device)), "ds_merge_nearest_ddim")

        if ring_state is not None:
            launch(panos, mask, mask_mode, mask_planes, ring_state.peer_table)
            self._peer_sync(ring_state, _lib.DS_PEER_AFTER_MERGE, wait=True)
        elif self.world == 1:
            launch(panos, mask, mask_mode, mask_planes, None)
        else:
            self._merge_sharded(panos, None, mask, plan, P, j0, nj, launch=launch)

    def _merge_state(self, st: PanoramaState, payloads, plan: StepPlan):
        """merge_all on a resident state: payloads[q] -> st.pms[q]; the state's mask receives 1."""
        assert plan.kind != "ring", "planar ring windows run on reference-order tensors"
        assert 0 < len(payloads) <= len(st.pms)
        pms = st.pms[:len(payloads)]
        for y in payloads:
            assert y.is_contiguous() and y.dtype == st.dtype
        j0, nj = self.local_jobs(plan)
        mask, mask_mode = None, _lib.DS_MASK_NONE
        if st.mask_plane is not None:
            assert plan.d_ftab is None, "a [1, H, W] mask panorama has no frame axis"
            mask, mask_mode = st.mask_plane, _lib.DS_MASK_PLANE
        elif st.mask_pm is not None:
            mask, mask_mode = st.mask_pm, _lib.DS_MASK_FULL
        if self.world == 1:
            self._pm_merge_into(pms, payloads, mask, mask_mode, plan, 0, plan.n_local)
            return
        if st.peer_table is not None:
            # fused exchange: this rank's surviving rows go straight into every replica (csrc/pm.cu)
            launched = self._pm_merge_into(pms, payloads, mask, mask_mode, plan, j0, nj, st.peer_table)
            if not (launched and self.merge_publish_in_kernel):  # an idle rank (or the host-sync protocol) publishes here
                self._peer_sync(st, _lib.DS_PEER_AFTER_MERGE)   # replicas are complete once every rank has pushed
            return
        # sharded: disjoint-support partial merge into zeroed buffers -> all-reduce -> select
        import torch.distributed as dist
        Q, n = len(pms), pms[0].numel()
        per_frame = plan.d_ftab is not None
        HW = plan.H * plan.W
        wlen = n if per_frame else HW
        key = ("pm", Q, n, wlen, st.dtype)
        buf = self._xbuf.get(key)
        if buf is None:
            buf = self._xbuf[key] = torch.empty(Q * n + wlen, dtype=st.dtype, device=st.device)
        buf.zero_()
        sums = [buf[q * n:(q + 1) * n] for q in range(Q)]
        weight = buf[Q * n:]
        if nj > 0:
            self._pm_merge_into(sums, payloads, weight, _lib.DS_MASK_FULL if per_frame else _lib.DS_MASK_PLANE,
                                plan, j0, nj)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        group, period = (1, n) if per_frame else (n // HW, HW)
        code, sp = _lib.dtype_code(st.dtype), _lib.stream_ptr(st.device)
        with _lib.on(st.device):
            for q in range(Q):
                lib.check(lib.ds_select_merged(pms[q].data_ptr(), sums[q].data_ptr(), weight.data_ptr(), n, period,
                                               group, code, sp), "ds_select_merged")
            if mask is not None:
                okey = ("one", mask.numel(), mask.dtype)
                if okey not in self._xbuf:
                    self._xbuf[okey] = torch.ones(mask.numel(), dtype=mask.dtype, device=mask.device)
                if mask_mode == _lib.DS_MASK_PLANE:
                    lib.check(lib.ds_select_merged(mask.data_ptr(), self._xbuf[okey].data_ptr(), weight.data_ptr(),
                                                   HW, HW, 1, code, sp), "ds_select_merged")
                else:
                    lib.check(lib.ds_select_merged(mask.data_ptr(), self._xbuf[okey].data_ptr(), weight.data_ptr(),
                                                   n, period, group, code, sp), "ds_select_merged")

    def _merge_ring_state(self, rs: RingPanoramaState, payloads, plan: StepPlan):
        """Planar ring merge in exchange mode: every element this rank's windows win goes into every
        replica (peer stores), then a full cross-GPU barrier publishes the step."""
        assert plan.kind == "ring" and 0 < len(payloads) <= len(rs.tensors)
        panos = rs.tensors[:len(payloads)]
        B, C, F, H, W = rs.shape
        P = B * C
        for y in payloads:
            assert y.is_contiguous() and y.dtype == rs.dtype
        j0, nj = self.local_jobs(plan)
        mask, mask_mode, mask_planes = rs.mask, _lib.DS_MASK_NONE, 0
        if mask is not None:
            if mask.numel() == H * W:
                assert plan.d_ftab is None, "a [1, H, W] mask panorama has no frame axis"
                mask_mode = _lib.DS_MASK_PLANE
            else:
                mask_mode, mask_planes = _lib.DS_MASK_FULL, mask.numel() // (F * H * W)
        if nj > 0 and plan.T > 0:
            with _lib.on(rs.device):
                lib.check(lib.ds_merge_nearest_peer(_lib.ptr_array(panos), _lib.ptr_array(payloads), len(panos),
                                                    _lib.ptr(mask), mask_mode, mask_planes, None, plan.d_src.data_ptr(),
                                                    plan.T, _lib.ptr(plan.d_ftab), plan.vpf, j0, nj, P, F, H * W,
                                                    plan.f_len, plan.hw, plan.vec, rs.peer_table,
                                                    _lib.dtype_code(rs.dtype), _lib.stream_ptr(rs.device)),
                          "ds_merge_nearest_peer")
        self._peer_sync(rs, _lib.DS_PEER_AFTER_MERGE, wait=True)

    def _merge_sharded(self, panos, payloads, mask, plan, P, j0, nj, launch=None):
        """disjoint-support partial merge -> one all-reduce -> select.  ``launch(sums, weight,
        mask_mode, mask_planes, peer)``: the fused-epilogue merge in place of the payload merge."""
        import torch.distributed as dist
        Q = len(panos)
        n = panos[0].numel()
        per_frame = plan.d_ftab is not None
        wlen = plan.F * plan.H * plan.W if per_frame else plan.H * plan.W
        key = (Q, n, wlen, panos[0].dtype)
        buf = self._xbuf.get(key)
        if buf is None:
            buf = self._xbuf[key] = torch.empty(Q * n + wlen, dtype=panos[0].dtype, device=panos[0].device)
        buf.zero_()
        sums = [buf[q * n:(q + 1) * n] for q in range(Q)]
        weight = buf[Q * n:]
        if launch is not None:
            launch([t.view(panos[0].shape) for t in sums], weight, _lib.DS_MASK_FULL if per_frame else _lib.DS_MASK_PLANE,
                   1 if per_frame else 0, None)
        elif nj > 0:
            self._merge_into(sums, payloads, weight, _lib.DS_MASK_FULL if per_frame else _lib.DS_MASK_PLANE,
                             1 if per_frame else 0, plan, P, j0, nj)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        with _lib.on(buf.device):
            for q in range(Q):
                lib.check(lib.ds_select_merged(panos[q].data_ptr(), sums[q].data_ptr(), weight.data_ptr(), n, wlen, 1,
                                               _lib.dtype_code(buf.dtype), _lib.stream_ptr(buf.device)),
                          "ds_select_merged")
            if mask is not None:
                okey = ("one", mask.numel(), mask.dtype)
                if okey not in self._xbuf:
                    self._xbuf[okey] = torch.ones(mask.numel(), dtype=mask.dtype, device=mask.device)
                one = self._xbuf[okey]
                lib.check(lib.ds_select_merged(mask.data_ptr(), one.data_ptr(), weight.data_ptr(), mask.numel(),
                                               wlen, 1, _lib.dtype_code(buf.dtype), _lib.stream_ptr(buf.device)),
                          "ds_select_merged")

    # ------------------------------------------------------------- bytes accounting ----
    def necessary_bytes(self, plan: StepPlan, P: int, elem_size: int, k: int = 2):
        """Bytes one step of THIS rank cannot avoid moving, as opposed to the SURVEY §8d formula of
        ``algorithmic_bytes`` (which charges the whole payload of every window although a merge only
        reads what survives the ordered writes).  Returns (gather_bytes, merge_bytes, detail):

        gather : the window tensors written (all of them) + every DISTINCT panorama row a tap of
                 some view reads (once) + the index / weight map entries;
        merge  : the surviving payload elements, counted in whole 32-byte sectors of the
                 [job][plane][frame][h*w] payload tensors (the smallest unit HBM / L2 move) + the
                 panorama rows written + the mask + the plan arrays.
        Frame rings: only (pixel, frame) pairs that are live in the frame table are counted."""
        SECT = 32
        per_sect = max(SECT // elem_size, 1)
        hw, f_len = plan.hw, plan.f_len
        n_fwin = plan.n_fwin
        row_bytes_full = P * elem_size                       # one (pixel, frame) atom
        if plan.d_ftab is None:
            live_frames = [f_len] * n_fwin                   # frames of frame window s that are written
            frames_read = plan.F
        else:
            ftab = plan.d_ftab.cpu().numpy()
            live_frames = [int(((ftab >= 0) & (ftab // f_len == s_)).sum()) for s_ in range(n_fwin)]
            frames_read = f_len                              # a job reads f_len frames of its rows
        live_total = sum(live_frames)
        detail = {}
        if plan.kind == "ring":
            src = plan.d_src.cpu().numpy()
            ok = src >= 0
            j0, nj = self.local_jobs(plan)
            mine = ok & (src // hw >= j0) & (src // hw < j0 + nj)
            sectors = np.unique(src[mine] // per_sect).size
            written = int(mine.sum())
            pay = sectors * SECT * P * live_total * k
            out = written * elem_size * P * live_total * k + written * elem_size * P * live_total   # + latent-shaped mask
            planb = src.size * 4
            gather = (nj * P * f_len * hw * elem_size                          # windows out
                      + min(int(ok.sum()), nj * hw) * P * frames_read * elem_size  # panorama pixels read once
                      + nj * 12)                                                  # window descriptors
            detail.update(payload_sectors=int(sectors), rows_written=written)
            return float(gather), float(pay + out + planb), detail
        vms = plan.extra["vms"]
        local_views = list(range(self.rank, len(vms), self.world))
        # ---- gather: distinct panorama pixels touched by this rank's views
        touched = np.zeros(plan.H * plan.W, bool)
        taps = 1 if plan.kind == "sphere_nearest" else 4
        for v in local_views:
            if taps == 1:
                touched[vms[v].idx_read] = True
            else:
                idx4, wt4 = vms[v].read_taps()
                touched[idx4[wt4 > 0]] = True
        n_touched = int(touched.sum())
        gather = (plan.n_local * P * f_len * hw * elem_size                       # windows out
                  + n_touched * row_bytes_full * (plan.F if plan.d_ftab is None else min(plan.F, f_len * n_fwin))
                  + len(local_views) * hw * taps * (4 if taps == 1 else 8))           # maps
        # ---- merge: surviving payload sectors + rows written + plan
        host = plan.extra["host_plan"]
        if plan.kind == "sphere_nearest":
            dst, src = host
            sectors = np.unique(src // per_sect).size                             # (view, k // 8): hw % 8 == 0 at these sizes
            rows = dst.size
            planb = src.size * 8
        else:
            row_dst, row_ptr, ent_src, ent_w = host
            sectors = np.unique(ent_src // per_sect).size
            rows = row_dst.size
            tiled = plan.extra.get("pm_tiled")
            planb = (tiled[4].numel() * 8 + rows * 8 + tiled[1].numel() * 4) if tiled is not None else ent_src.size * 8 + rows * 8
        pay = sectors * SECT * P * live_total * k
        out = rows * row_bytes_full * live_total * k
        mask = rows * elem_size if plan.d_ftab is None else rows * row_bytes_full * live_total
        detail.update(payload_sectors=int(sectors), rows_written=int(rows), pano_pixels_read=n_touched,
                      payload_bytes=float(pay), written_bytes=float(out + mask), plan_bytes=float(planb))
        return float(gather), float(pay + out + mask + planb), detail

    def algorithmic_bytes(self, plan: StepPlan, P: int, elem_size: int, k: int = 2):
        """N*V*s [gather out] + cov*P*s [pano read once] + k*N*V*s [payloads in]
        + k*cov*P*s [pano written once] (+ mask).  Returns (gather_bytes, merge_bytes)."""
        V = P * plan.f_len * plan.hw * elem_size
        frames_touched = plan.F if plan.d_ftab is None else int((plan.d_ftab >= 0).sum().item())
        pano = plan.coverage * P * frames_touched * plan.H * plan.W * elem_size
        gather = plan.n_jobs * V + pano
        merge = k * plan.n_jobs * V + k * pano
        return gather, merge
