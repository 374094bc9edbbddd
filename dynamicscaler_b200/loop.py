"""Denoising-loop driver over the resident panorama state (VERDICT row N1):
the reference's DDIM loop (pipeline/t2v_sphere_panorama_pipeline.py:166-293, I2V twin
pipeline/i2v_sphere_panorama_pipeline.py:232-476) with the UNet -- the reference's own torch module,
untouched -- BATCHED over the windows of a group, and everything either side of it fused into two
launches per group:

    gather  (+ mask read + re_noise + mask mix)            ds_pm_gather / ds_pm_gather_renoise
    UNet    model(x[n, C, f, h, w], timesteps of length n*f, ...)   one call for the whole group (P9)
    merge   (CFG combine + ddim_step + merge_prev mix + latent / denoised / mask sets)
                                                           ds_pm_merge_*_ddim

``group_size``: how many consecutive windows are read before any of them is written back.
``None`` = the whole step (the batched order BASELINE.json config 3 shards over GPUs); ``1`` = the
reference's own sequential order, reproduced bit for bit (tests/golden/loop.npz comes from the real
``basic_sample_shift_shpere_panorama``).  With ``stepper.world > 1`` the views of every group are dealt
round-robin over the ranks, each rank runs the UNet on its own windows only (``balance_evals=True``
additionally moves single uncond evaluations between ranks so that 44 views x 2 CFG branches are 11
evaluations on each of 8 ranks instead of 12 / 10), and the merge kernels push the surviving rows into
every replica (step.py).

The planar shift-window loops (``basic_sample_shift_multi_windows`` of pipeline/t2v_normal_pipeline.py,
pipeline/t2v_sphere_panorama_pipeline.py and pipeline/i2v_sphere_panorama_pipeline.py) have the same driver on
reference-order tensors -- :func:`ring_denoise_loop` with the window schedules :func:`shift_window_schedule`,
:func:`t2v_planar_schedule`, :func:`i2v_planar_schedule`:

    gather  ds_ring_gather (+ mask windows + ds_renoise_mix)
    UNet    one call per group
    merge   ds_merge_nearest_ddim (CFG + ddim_step + merge_prev mix + latent / denoised / mask sets)

The model is called exactly like the reference calls ``self.pretrained_t2v.model`` (:236-254):
``model(x, ts, c_crossattn=[ctx], fps=fps, curr_time_steps=ts, temporal_length=frames, **kwargs)``.
``UNetModel.forward`` with a batch of b > 1 windows needs ``timesteps`` of length b*t, not b
(lvdm/modules/networks/openaimodel3d.py:667-682: the time embedding is not repeated over t), so the
driver passes ``ts`` of length n*frames for n > 1 and of length n for a single window, as the reference does.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch

from .step import PanoramaState, PanoramaStepper, ddim_coefficients, renoise_coefficients


def plan_eval_balance(n_local_per_rank: Sequence[int]):
    """Classifier-free guidance doubles every window into two UNet evaluations (cond, uncond).  Views are dealt
    round-robin, so ranks own n or n+1 windows -- 2n or 2n+2 evaluations (44 views on 8 ranks: 12,12,12,12,10,10,
    10,10).  Moving single UNCOND evaluations from the ranks with more to the ranks with fewer evens that out
    (88 = 8 x 11).  Returns the list of transfers ``(donor rank, receiver rank, local window index at the donor)``
    -- the donor's LAST windows first -- identical on every rank."""
    world = len(n_local_per_rank)
    evals = [2 * n for n in n_local_per_rank]
    base, extra = divmod(sum(evals), world)
    order = sorted(range(world), key=lambda r: (-evals[r], r))
    target = {r: base + (1 if i < extra else 0) for i, r in enumerate(order)}
    surplus = {r: evals[r] - target[r] for r in range(world)}
    donors = [r for r in order if surplus[r] > 0]
    receivers = [r for r in reversed(order) if surplus[r] < 0]
    transfers, given = [], {r: 0 for r in range(world)}
    for d in donors:
        while surplus[d] > 0 and receivers:
            q = receivers[0]
            if given[d] >= n_local_per_rank[d]:
                break                                          # nothing left to donate (every uncond already given away)
            given[d] += 1
            transfers.append((d, q, n_local_per_rank[d] - given[d]))
            surplus[d] -= 1
            surplus[q] += 1
            if surplus[q] == 0:
                receivers.pop(0)
    return transfers


def exchange_windows(tensors_out, tensors_in, group=None):
    """One batch of point-to-point transfers: ``tensors_out`` = [(tensor, peer rank)] to send, ``tensors_in`` =
    [(tensor, peer rank)] to receive into (torch.distributed P2P: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    ops = [dist.P2POp(dist.isend, t, peer, group) for t, peer in tensors_out]
    ops += [dist.P2POp(dist.irecv, t, peer, group) for t, peer in tensors_in]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def _get(schedule, name):
    return schedule[name] if isinstance(schedule, dict) else getattr(schedule, name)


def unet_eval(model: Callable, x: torch.Tensor, t: int, cond: dict, frames: int, model_kwargs: Optional[dict] = None,
              uncond: Optional[dict] = None, cfg_batch: bool = True, x_uncond: Optional[torch.Tensor] = None):
    """Run the UNet on ``x`` [n, C, f, h, w] for the conditional and (if ``uncond`` is given) the
    unconditional branch.  ``cond`` / ``uncond``: the reference's dicts ``{"c_crossattn": [emb], "fps": fps}``
    with ``emb`` [1 or n, L, D].  ``cfg_batch``: evaluate both branches as ONE batch of 2n windows.
    Returns (eps_cond, eps_uncond or None), each [n, C, f, h, w] contiguous."""
    n = x.shape[0]
    kwargs = dict(model_kwargs or {})
    xu = x if x_uncond is None else x_uncond        # eval balancing: the uncond branch may run on another window set

    def ctx_of(c, count=None):
        count = n if count is None else count
        emb = torch.cat(c["c_crossattn"], dim=1)
        return emb.expand(count, -1, -1) if emb.shape[0] == 1 and count != 1 else emb

    def call(xx, ctx, fps):
        b = xx.shape[0]
        ts = torch.full((b * frames if b > 1 else b,), int(t), device=xx.device, dtype=torch.long)
        return model(xx, ts, c_crossattn=[ctx], fps=fps, curr_time_steps=ts, temporal_length=frames, **kwargs)

    if uncond is None:
        return call(x, ctx_of(cond), cond.get("fps", 16)).contiguous(), None
    nu = xu.shape[0]
    if cfg_batch and n + nu > 0:
        y = call(torch.cat([x, xu], dim=0), torch.cat([ctx_of(cond), ctx_of(uncond, nu)], dim=0), cond.get("fps", 16))
        return y[:n].contiguous(), y[n:].contiguous()
    ec = call(x, ctx_of(cond), cond.get("fps", 16)).contiguous() if n else x
    eu = call(xu, ctx_of(uncond, nu), uncond.get("fps", 16)).contiguous() if nu else xu
    return ec, eu


def sphere_denoise_step(stepper: PanoramaStepper, state: PanoramaState, views: Sequence, width: int, height: int, *,
                        model: Callable, t: int, step_i: int, schedule, cond: dict, uncond: Optional[dict],
                        guidance_scale: float, frames: int, fov=120, group_size: Optional[int] = None,
                        renoise_ratio: Optional[float] = None, prev_ratio: Optional[float] = None,
                        eps_fn: Optional[Callable] = None, noise_fn: Optional[Callable] = None,
                        model_kwargs: Optional[dict] = None, cfg_batch: bool = True, frame_begins=None,
                        on_input: Optional[Callable] = None, balance_evals: bool = False):
    """One DDIM step i over ``views`` (ordered (phi, theta) list of the step).  ``schedule``: object or
    dict with the reference scheduler's ``ddim_timesteps, alphas_cumprod, ddim_alphas, ddim_alphas_prev,
    ddim_sigmas, ddim_sqrt_one_minus_alphas`` (pipeline/scheduler.py:17-57).  ``eps_fn(step_i, job_ids,
    shape, device)`` supplies the re_noise epsilon, ``noise_fn`` the ddim noise when sigma > 0 (the
    reference draws them from the global RNG).  ``on_input(step_i, job_ids, x)``: observer of
    the model inputs (tests).  ``frame_begins``: ordered frame-window begins (I2V temporal loop); the job list
    is frame windows (outer) x views (inner) and a ``group_size`` group never spans two frame windows.
    ``cond`` may be a callable ``cond(step_i, job_ids, boxes)`` -> dict with one context row per window
    (``boxes``: (phi, theta, frame_begin) per job; the I2V loop builds one image embedding per view, i2v :367-395)."""
    total = len(_get(schedule, "ddim_timesteps"))
    index = total - step_i - 1                                          # t2v_sphere_panorama_pipeline.py:263
    coef = ddim_coefficients(_get(schedule, "ddim_alphas")[index], _get(schedule, "ddim_alphas_prev")[index],
                             _get(schedule, "ddim_sigmas")[index], _get(schedule, "ddim_sqrt_one_minus_alphas")[index],
                             guidance_scale if uncond is not None else 1.0)
    sigma = float(_get(schedule, "ddim_sigmas")[index])
    mixing = step_i < total - 1                                         # :212, i2v :430
    use_renoise = renoise_ratio is not None and mixing
    use_prev = prev_ratio is not None and mixing
    if use_renoise or use_prev:
        assert eps_fn is not None or not use_renoise, "re_noise needs an epsilon source (eps_fn)"
        rc, rs = renoise_coefficients(_get(schedule, "alphas_cumprod"), _get(schedule, "ddim_timesteps"),
                                      total - step_i - 2, total - step_i - 1) if use_renoise else (1.0, 0.0)
    state.reset_mask()                                                  # :173 fresh mask panorama every step
    nv = len(views)
    if group_size is None or frame_begins is None:                      # whole step, or groups of views (no frame windows)
        gs = nv if group_size is None else int(group_size)
        groups = [(g0, 0, None if frame_begins is None else list(frame_begins)) for g0 in range(0, nv, gs)]
    else:                                                               # i2v :284: frame windows outer, views inner
        gs = int(group_size)
        groups = [(g0, s, [fb]) for s, fb in enumerate(frame_begins) for g0 in range(0, nv, gs)]
    for g0, s0, fbs in groups:
        fkw = {} if fbs is None else dict(frame_begins=fbs, frames=frames)
        plan = stepper.plan_sphere(list(views[g0:g0 + gs]), width, height, fov=fov, mode="nearest", **fkw)
        nvp = len(plan.extra["vms"])                                    # views of this group (all ranks)
        job_ids = [(s0 + j // nvp) * nv + g0 + j % nvp for j in plan.job_ids]   # indices in the step's ordered job list
        raw = mview = None
        if use_renoise or use_prev:
            shape = (plan.n_local, state.shape[0], state.shape[1], plan.f_len, plan.h, plan.w)
            eps = eps_fn(step_i, job_ids, shape, state.device) if use_renoise else \
                torch.zeros(shape, dtype=state.dtype, device=state.device)
            x, raw, mview = stepper.gather_all(state, plan, renoise=dict(
                eps=eps.contiguous(), c=rc, s=rs, ratio=renoise_ratio if use_renoise else 0.0,
                want_raw=use_prev, want_mask=use_prev))
        else:
            x = stepper.gather_all(state, plan)
        n = x.shape[0]
        xb = x.view(n * x.shape[1], *x.shape[2:])                       # [n*B, C, f, h, w]; B == 1 in every pipeline
        if on_input is not None and n > 0:
            on_input(step_i, job_ids, xb)
        if balance_evals and stepper.world > 1 and uncond is not None:
            assert not callable(cond), "eval balancing takes one shared cond dict"
            ec, eu = _balanced_eval(stepper, plan, model, xb, t, cond, uncond, frames, model_kwargs, cfg_batch)
            ec, eu = ec.view(x.shape), eu.view(x.shape)
        elif n > 0:
            c = cond
            if callable(cond):                  # one context row per window (i2v :367-395: one image embedding per view)
                boxes = [(views[g0 + j % nvp][0], views[g0 + j % nvp][1], None if fbs is None else fbs[j // nvp])
                         for j in plan.job_ids]
                c = cond(step_i, job_ids, boxes)
            ec, eu = unet_eval(model, xb, t, c, frames, model_kwargs, uncond, cfg_batch)
            ec = ec.view(x.shape)
            eu = None if eu is None else eu.view(x.shape)
        else:
            ec, eu = x, (x if uncond is not None else None)
        noise = noise_fn(step_i, job_ids, tuple(x.shape), state.device) if (noise_fn is not None and sigma > 0) else None
        stepper.merge_denoised(state, x, ec, eu, coef, plan, noise=noise, prev=raw if use_prev else None,
                               prev_mask=mview if use_prev else None, prev_ratio=prev_ratio if use_prev else None)


def _balanced_eval(stepper, plan, model, xb, t, cond, uncond, frames, model_kwargs, cfg_batch):
    """UNet evaluations of one group with single uncond evaluations moved between ranks
    (:func:`plan_eval_balance`): donors ship the window to the receiver, the receiver evaluates the uncond
    branch in its own batch and ships the result back; the merge stays at the window's owner."""
    n_views = len(plan.extra["vms"])
    n_fwin = plan.n_fwin
    per_rank = [len(range(r, n_views, stepper.world)) * n_fwin for r in range(stepper.world)]
    moves = plan_eval_balance(per_rank)
    mine_out = [j for (d, q, j) in moves if d == stepper.rank]
    peers_out = [q for (d, q, j) in moves if d == stepper.rank]
    mine_in = [(d, j) for (d, q, j) in moves if q == stepper.rank]
    n = xb.shape[0]
    x_in = xb.new_empty((len(mine_in),) + tuple(xb.shape[1:]))
    exchange_windows([(xb[j].contiguous(), q) for j, q in zip(mine_out, peers_out)],
                     [(x_in[i], d) for i, (d, j) in enumerate(mine_in)], stepper.group)
    keep = [j for j in range(n) if j not in set(mine_out)]
    xu = torch.cat([xb[keep], x_in], dim=0) if (len(keep) != n or len(mine_in)) else xb
    ec, eu_mix = unet_eval(model, xb, t, cond, frames, model_kwargs, uncond, cfg_batch, x_uncond=xu)
    eu = torch.empty_like(xb)
    if keep:
        eu[keep] = eu_mix[:len(keep)]
    back = eu_mix[len(keep):]
    recv = [eu.new_empty(xb.shape[1:]) for _ in mine_out]
    exchange_windows([(back[i].contiguous(), d) for i, (d, j) in enumerate(mine_in)],
                     [(recv[i], q) for i, q in enumerate(peers_out)], stepper.group)
    for j, r in zip(mine_out, recv):
        eu[j] = r
    return ec, eu


def sphere_denoise_loop(stepper: PanoramaStepper, state: PanoramaState, *, model: Callable, schedule,
                        views_fn: Callable, width: int, height: int, frames: int, cond: dict,
                        uncond: Optional[dict], guidance_scale: float = 7.5, fov=120,
                        timesteps: Optional[Sequence[int]] = None, group_size: Optional[int] = None,
                        renoise_ratio: Optional[float] = None, prev_ratio_list: Optional[Sequence[float]] = None,
                        eps_fn: Optional[Callable] = None, noise_fn: Optional[Callable] = None,
                        model_kwargs: Optional[dict] = None, cfg_batch: bool = True, frame_begins_fn=None,
                        on_input: Optional[Callable] = None, balance_evals: bool = False):
    """The whole DDIM loop (t2v_sphere_panorama_pipeline.py:164-293): ``timesteps`` defaults to the
    flipped ``ddim_timesteps`` (:103); ``views_fn(i)`` returns the ordered (phi, theta) list of step i
    (``geometry.sphere_views``).  The state's panorama 0 / 1 end up holding the final latents / the
    denoised prediction (:295-296)."""
    ddim_ts = [int(v) for v in _get(schedule, "ddim_timesteps")]
    timesteps = list(reversed(ddim_ts)) if timesteps is None else list(timesteps)
    kwargs = dict(model_kwargs or {})
    kwargs.setdefault("clean_cond", True)                               # :232
    for i, t in enumerate(timesteps):
        sphere_denoise_step(stepper, state, views_fn(i), width, height, model=model, t=t, step_i=i, schedule=schedule,
                            cond=cond, uncond=uncond, guidance_scale=guidance_scale, frames=frames, fov=fov,
                            group_size=group_size, renoise_ratio=renoise_ratio,
                            prev_ratio=None if prev_ratio_list is None else prev_ratio_list[i], eps_fn=eps_fn,
                            noise_fn=noise_fn, model_kwargs=kwargs, cfg_batch=cfg_batch,
                            frame_begins=None if frame_begins_fn is None else frame_begins_fn(i), on_input=on_input,
                            balance_evals=balance_evals)
    return state


# ------------------------------------------------------------------ planar shift-window loop ----
def shift_window_schedule(step_i: int, *, frames: int, latent_height: int, latent_width: int, num_windows_f: int,
                          num_windows_h: int, num_windows_w: int, loop_step: int, vae_scale_factor: int = 8,
                          shift_jump_odd_w: bool = False, shift_jump_odd_h: bool = False, shift_jump_odd_f: bool = False,
                          docking_w: bool = False, docking_h: bool = False, docking_f: bool = False,
                          docking_step_range=None):
    """Window positions of DDIM step ``step_i`` of ``basic_sample_shift_multi_windows``
    (pipeline/t2v_normal_pipeline.py:420-432 step sizes, :438-440 and :473-478 per-step starts, :483-524
    the three nested index loops with docking).  Returns ``(windows, frame_begins)``: the ordered
    (left, top) list (w outer, h inner) and the ordered frame-window begins -- the reference visits
    ``for f: for w: for h``, which is ``plan_ring(windows, ..., frame_begins=frame_begins)``'s job order.

    Kept as the reference has them: the per-axis shift is ``(size_px // loop_step) // vae`` latent cells
    (frames: ``frames // loop_step``), zero for a single window on that axis; ``shift_jump_odd_h`` moves
    the LEFT start by half the panorama width and ``shift_jump_odd_w`` the TOP start by half its height
    (:473-476, the names are crossed there); at a docking step the extra window -1 sits at 0 and the
    last window at ``size * (n - 1)``, the others stay shifted."""
    def shift(size_latent, n):
        return 0 if n == 1 else (size_latent * vae_scale_factor // loop_step) // vae_scale_factor

    odd = step_i % 2 == 1
    k = step_i % loop_step
    left0 = k * shift(latent_width, num_windows_w)
    top0 = k * shift(latent_height, num_windows_h)
    f0 = k * (0 if num_windows_f == 1 else frames // loop_step)
    assert num_windows_f == 1 or frames // loop_step > 0, f"loop_step {loop_step} > frames {frames}"
    if odd and shift_jump_odd_h and num_windows_h > 1:
        left0 += latent_width * num_windows_w // 2
    if odd and shift_jump_odd_w and num_windows_w > 1:
        top0 += latent_height * num_windows_h // 2
    if odd and shift_jump_odd_f and num_windows_f > 1:
        f0 += frames * num_windows_f // 2
    assert not (docking_w or docking_h or docking_f) or docking_step_range is not None, "docking needs docking_step_range"

    def axis(start, n, size, docking):
        dock_now = docking and step_i in docking_step_range
        out = []
        for idx in (range(-1, n) if docking else range(n)):
            b = start + idx * size
            if dock_now:
                if idx == -1:
                    b = 0
                if idx == n - 1:
                    b = size * (n - 1)
            elif idx == -1:
                continue
            out.append(b)
        return out

    lefts = axis(left0, num_windows_w, latent_width, docking_w)
    tops = axis(top0, num_windows_h, latent_height, docking_h)
    return [(l, t) for l in lefts for t in tops], axis(f0, num_windows_f, frames, docking_f)


def i2v_planar_schedule(step_i: int, *, width: int, height: int, frames: int, total_w: int, total_h: int, total_f: int,
                        num_windows_w: int, num_windows_h: int, loop_step: int, overlap_ratio_f: float,
                        loop_step_frame: int, dock_at_f: bool = False, begin_index_offset: int = 0,
                        vae_scale_factor: int = 8):
    """Window positions of DDIM step ``step_i`` of the planar I2V loop (``VC2_Pipeline_I2V_SpherePano.
    basic_sample_shift_multi_windows``, pipeline/i2v_sphere_panorama_pipeline.py:732-855): overlapping spatial
    windows at ``round(k * stride)`` plus the per-step offset (``begin_index_offset`` included), frame windows with
    optional docking.  Pixel sizes in; returns ``(windows, frame_begins)`` in latent cells / frames, the
    ``for f: for w: for h`` order of the reference as ``plan_ring``'s job order."""
    from .geometry import frame_windows, ring_windows
    windows = ring_windows(step_i, num_windows_w, num_windows_h, total_w, total_h, width=width, height=height,
                           loop_step=loop_step, vae_scale=vae_scale_factor, begin_index_offset=begin_index_offset)
    return windows, frame_windows(step_i, frames=frames, total_f=total_f, overlap=overlap_ratio_f,
                                  loop_step_frame=loop_step_frame, dock_at_f=dock_at_f)


def t2v_planar_schedule(step_i: int, *, width: int, height: int, frames: int, total_w: int, total_h: int,
                        num_windows_w: int, num_windows_h: int, num_windows_f: int, loop_step: int,
                        dock_at_h: bool = False, vae_scale_factor: int = 8):
    """Window positions of DDIM step ``step_i`` of the T2V planar loop (``VC2_Pipeline_T2V_SpherePano.
    basic_sample_shift_multi_windows``, pipeline/t2v_sphere_panorama_pipeline.py:436-535): integer strides
    ``int(size * (1 - overlap)) // vae``, per-step offsets, non-overlapping frame windows shifted by
    ``frames // loop_step`` per step, and -- ``dock_at_h`` -- extra windows docked at the top edge (first) and the
    bottom edge (last) on every step whose shift is not 0, with windows running past the bottom edge dropped.
    Pixel sizes in; returns ``(windows, frame_begins)`` in latent cells / frames (``for f: for w: for h``)."""
    vae = vae_scale_factor

    def axis(num, total, size):
        overlap = 1 - (total / size - 1) / (num - 1)
        return int(size * (1 - overlap)) // vae, int((1 - overlap) * size / loop_step) // vae

    stride_w, off_w = axis(num_windows_w, total_w, width)
    stride_h, off_h = axis(num_windows_h, total_h, height)
    phase = step_i % loop_step
    lat_h, lat_total_h = height // vae, total_h // vae
    tops = [phase * off_h + k * stride_h for k in range(num_windows_h)]
    if dock_at_h:
        if phase != 0:
            tops = [0] + tops + [lat_total_h - lat_h]
        tops = [t for t in tops if t + lat_h <= lat_total_h]
    lefts = [phase * off_w + k * stride_w for k in range(num_windows_w)]
    step_f = 0 if num_windows_f == 1 else frames // loop_step
    return [(l, t) for l in lefts for t in tops], [phase * step_f + k * frames for k in range(num_windows_f)]


def ring_denoise_step(stepper: PanoramaStepper, target, windows: Sequence, frame_begins: Sequence[int], w: int, h: int,
                      frames: int, *, model: Callable, t: int, index: int, step_i: int, schedule, cond,
                      uncond: Optional[dict], guidance_scale: float, group_size: Optional[int] = None,
                      noise_fn: Optional[Callable] = None, model_kwargs: Optional[dict] = None, cfg_batch: bool = True,
                      on_input: Optional[Callable] = None, mask: Optional[torch.Tensor] = None,
                      renoise: Optional[tuple] = None, prev_ratio: Optional[float] = None,
                      eps_fn: Optional[Callable] = None, mask_frame0: bool = False):
    """One DDIM step of the planar shift-window loops (pipeline/t2v_normal_pipeline.py:481-557; I2V twin
    pipeline/i2v_sphere_panorama_pipeline.py:812-968) over the ordered job list ``frame_begins x windows``:

        gather  ds_ring_gather (+ the mask windows + ds_renoise_mix: re_noise and the mask mix, one pass)
        UNet    ONE call per group
        merge   ds_merge_nearest_ddim: CFG + ddim_step (+ merge_prev mix) + latent / denoised / mask sets

    ``target``: ``[latent, denoised]`` reference-order tensors, or a ``RingPanoramaState`` (world_size > 1, peer
    stores).  ``index``: the scheduler index ``total_steps - i - 1`` (:611).  ``group_size`` None: the whole step is
    one group (for a T2V step without docking the windows tile the torus without overlap, so this IS the reference's
    result); an integer g: g consecutive windows of one frame window (1 = the reference's sequential order, needed
    whenever windows overlap).  I2V extras: ``mask`` = the latent-shaped denoised-mask panorama of this step (already
    zeroed, i2v :810; a RingPanoramaState carries its own); ``renoise`` = ``(c, s, ratio)`` switches the re-noise mix
    on (i2v :875-884, epsilon from ``eps_fn(step_i, job_ids, shape, device)``); ``prev_ratio`` the merge_prev mix
    (i2v :938-943); ``cond`` may be a callable ``cond(step_i, job_ids, boxes)`` -> dict with one context row per
    window (``boxes``: the (left, top, frame_begin) of each job; i2v :886-895 builds one image embedding per window).
    ``mask_frame0``: the T2V sphere pipeline's planar loop mixes with ``mask_window[0, 0, [0]]`` broadcast over the
    window (pipeline/t2v_sphere_panorama_pipeline.py:553)."""
    coef = ddim_coefficients(_get(schedule, "ddim_alphas")[index], _get(schedule, "ddim_alphas_prev")[index],
                             _get(schedule, "ddim_sigmas")[index], _get(schedule, "ddim_sqrt_one_minus_alphas")[index],
                             guidance_scale if uncond is not None else 1.0)
    sigma = float(_get(schedule, "ddim_sigmas")[index])
    nw = len(windows)
    if group_size is None:
        groups = [(list(windows), list(frame_begins), None)]
    else:
        gs = int(group_size)
        groups = [(list(windows[g0:g0 + gs]), [fb], s * nw + g0) for s, fb in enumerate(frame_begins)
                  for g0 in range(0, nw, gs)]
    is_state = not isinstance(target, (list, tuple))
    source = target if is_state else target[0]
    use_prev = prev_ratio is not None
    assert renoise is None or eps_fn is not None, "re_noise needs an epsilon source (eps_fn)"
    for wins, fbs, base in groups:
        plan = stepper.plan_ring(wins, w, h, frame_begins=fbs, frames=frames)
        job_ids = list(plan.job_ids) if base is None else [base + j for j in plan.job_ids]
        raw = mview = None
        if renoise is not None or use_prev:
            shape = (plan.n_local,) + tuple(source.shape[:2]) + (plan.f_len, plan.h, plan.w)
            rn = dict(want_raw=use_prev, want_mask=use_prev, mask=mask, mask_frame0=mask_frame0)
            if renoise is not None:
                eps = eps_fn(step_i, job_ids, shape, stepper.device)
                rn.update(eps=eps.contiguous(), c=renoise[0], s=renoise[1], ratio=renoise[2])
            x, raw, mview = stepper.gather_all(source, plan, renoise=rn)
        else:
            x = stepper.gather_all(source, plan)
        n = x.shape[0]
        xb = x.view(n * x.shape[1], *x.shape[2:])
        if on_input is not None and n > 0:
            on_input(step_i, job_ids, xb)
        if n > 0:
            c = cond
            if callable(cond):
                boxes = [(wins[j % len(wins)][0], wins[j % len(wins)][1], fbs[j // len(wins)]) for j in plan.job_ids]
                c = cond(step_i, job_ids, boxes)
            ec, eu = unet_eval(model, xb, t, c, frames, model_kwargs, uncond, cfg_batch)
            ec = ec.view(x.shape)
            eu = None if eu is None else eu.view(x.shape)
        else:
            ec, eu = x, (x if uncond is not None else None)
        noise = noise_fn(step_i, job_ids, tuple(x.shape), x.device) if (noise_fn is not None and sigma > 0) else None
        stepper.merge_denoised_ref(target, x, ec, eu, coef, plan, noise=noise, mask=None if is_state else mask,
                                   prev=raw if use_prev else None, prev_mask=mview if use_prev else None,
                                   prev_ratio=prev_ratio)


def ring_denoise_loop(stepper: PanoramaStepper, target, *, model: Callable, schedule, schedule_fn: Callable, width: int,
                      height: int, frames: int, cond, uncond: Optional[dict], guidance_scale: float = 7.5,
                      timesteps: Optional[Sequence[int]] = None, group_size=None, noise_fn: Optional[Callable] = None,
                      model_kwargs: Optional[dict] = None, cfg_batch: bool = True, on_input: Optional[Callable] = None,
                      before_step: Optional[Callable] = None, mask: Optional[torch.Tensor] = None,
                      renoise_ratio: Optional[float] = None, prev_ratio_list: Optional[Sequence[float]] = None,
                      eps_fn: Optional[Callable] = None, mask_frame0: bool = False):
    """The planar DDIM loops: ``basic_sample_shift_multi_windows`` of pipeline/t2v_normal_pipeline.py:434-561 and, with
    ``mask`` / ``renoise_ratio`` / ``prev_ratio_list`` / a callable ``cond``, of pipeline/i2v_sphere_panorama_pipeline.py:
    777-970 (stages 2 / 3 of gen_pano_360.py).  ``target``: ``[latent, denoised]`` ([B, C, F_total, H_total, W_total];
    the reference starts the denoised panorama at zero, :417) or a ``RingPanoramaState``; both are updated in place,
    ``target[1]`` is what the reference returns (:561).  ``schedule_fn(i)`` -> ``(windows, frame_begins)`` of step i
    (:func:`shift_window_schedule` / :func:`i2v_planar_schedule` / :func:`t2v_planar_schedule`); ``width`` / ``height``: the window in latent cells.
    ``timesteps`` defaults to the flipped ``ddim_timesteps`` (:300); a shorter list is the reference's
    ``use_skip_time`` tail (:303).  ``group_size``: see :func:`ring_denoise_step`; a callable ``group_size(i)`` picks
    it per step (e.g. 1 on docking steps only).  ``mask``: a latent-shaped tensor that serves as the denoised-mask
    panorama (zeroed here at every step, i2v :810; required for the re-noise / merge_prev mixes on plain tensors).
    ``mask_frame0=True`` selects the T2V sphere pipeline's planar loop (pipeline/t2v_sphere_panorama_pipeline.py:
    481-632: re-noise mix with ``mask_window[0, 0, [0]]``, no merge_prev mix).
    ``before_step(i, target)``: hook for the whole-panorama residual mix of t2v :442-470 (outside the tiling path)."""
    ddim_ts = [int(v) for v in _get(schedule, "ddim_timesteps")]
    timesteps = list(reversed(ddim_ts)) if timesteps is None else [int(v) for v in timesteps]
    total = len(timesteps)
    kwargs = dict(model_kwargs or {})
    kwargs.setdefault("clean_cond", True)                               # :581
    is_state = not isinstance(target, (list, tuple))
    step_mask = target.mask if is_state else mask
    for i, t in enumerate(timesteps):
        if before_step is not None:
            before_step(i, target)
        mixing = i < total - 1                                          # i2v :875, :938
        renoise = None
        if renoise_ratio is not None and mixing:
            rc, rs = renoise_coefficients(_get(schedule, "alphas_cumprod"), _get(schedule, "ddim_timesteps"),
                                          total - i - 2, total - i - 1)
            renoise = (rc, rs, renoise_ratio)
        prev_ratio = prev_ratio_list[i] if (prev_ratio_list is not None and mixing) else None
        if step_mask is not None:
            step_mask.zero_()                                           # i2v :810 fresh mask panorama every step
        else:
            assert renoise is None and prev_ratio is None, "the re-noise / merge_prev mixes read a mask panorama"
        windows, frame_begins = schedule_fn(i)
        ring_denoise_step(stepper, target, windows, frame_begins, width, height, frames, model=model, t=t,
                          index=total - i - 1, step_i=i, schedule=schedule, cond=cond, uncond=uncond,
                          guidance_scale=guidance_scale, group_size=group_size(i) if callable(group_size) else group_size,
                          noise_fn=noise_fn, model_kwargs=kwargs, cfg_batch=cfg_batch, on_input=on_input,
                          mask=step_mask, renoise=renoise, prev_ratio=prev_ratio, eps_fn=eps_fn, mask_frame0=mask_frame0)
    return target
