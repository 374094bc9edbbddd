"""Per-step window geometry of the reference's denoising loops, as plain Python: which
views / planar windows / frame windows a step visits and in which order.  The step driver
(`step.PanoramaStepper`) takes these lists; `bench.py` and the tests build their workloads
from them.  Pure host arithmetic, no tensors."""
from __future__ import annotations

import math
from typing import List, Tuple


def sphere_views(step_i: int = 0, phi_num: int = 6, loop_step_theta: int = 10, view_fov: int = 120
                 ) -> List[Tuple[int, int]]:
    """Ordered (phi, theta) list of one sphere step: the default ``phi_theta_dict`` of
    gen_pano_360.py:444-455 (poles once, seven latitude bands x ``phi_num`` longitudes) walked
    phi-outer / theta-inner like t2v_sphere_panorama_pipeline.py:176-187, every longitude
    shifted by ``(i % loop_step_theta) * (view_fov // loop_step_theta)`` (:169)."""
    lon = [360 * n // phi_num for n in range(phi_num)]
    bands = [(90, [0]), (-90, [0])]
    for lat in (75, 60, 45):
        bands += [(lat, lon), (-lat, lon)]
    bands.append((0, lon))
    shift = (step_i % loop_step_theta) * (view_fov // loop_step_theta)
    return [(phi, theta + shift) for phi, thetas in bands for theta in thetas]


def ring_windows(step_i: int, num_w: int, num_h: int, total_w: int, total_h: int, width: int = 512,
                 height: int = 320, loop_step: int = 16, vae_scale: int = 8, begin_index_offset: int = 0
                 ) -> List[Tuple[int, int]]:
    """Ordered (left, top) latent positions of the planar shift windows of one step
    (i2v_sphere_panorama_pipeline.py:732-752 overlaps and strides, :779-783 per-step offset incl.
    ``begin_index_offset``, :818-821 positions; w outer, h inner).  Pixel sizes in, latent positions out."""
    def axis(num, total, size):
        if num <= 1:
            return 0.0, 0
        overlap = 1 - (total / size - 1) / (num - 1)
        stride = size / vae_scale * (1 - overlap)
        offset = int((1 - overlap) * size / loop_step) // vae_scale
        return stride, offset

    stride_w, off_w = axis(num_w, total_w, width)
    stride_h, off_h = axis(num_h, total_h, height)
    phase = (step_i + begin_index_offset) % loop_step
    return [(phase * off_w + round(a * stride_w), phase * off_h + round(b * stride_h))
            for a in range(num_w) for b in range(num_h)]


def frame_windows(step_i: int, frames: int = 16, total_f: int = 64, overlap: float = 0.75,
                  loop_step_frame: int = 8, dock_at_f: bool = False) -> List[int]:
    """Frame-window begins of one step, in loop order (i2v_sphere_panorama_pipeline.py:785-855; sphere twin
    :256-325).  ``dock_at_f`` off: windows wrap modulo ``total_f`` (the temporal loop).  On: an extra window
    docked at frame 0 comes first and one at ``total_f - frames`` last -- both dropped on the steps whose
    shift is 0 --, and every window that would run past ``total_f`` is dropped (:832-855)."""
    if total_f == frames:
        return [0]
    count = math.ceil((total_f // frames - 1) / (1 - overlap)) + 1
    per_step = max(int(overlap * frames / loop_step_frame), 1)
    hop = int(frames * (1 - overlap))
    first = (step_i % loop_step_frame) * per_step
    begins = [(first + n * hop) % total_f for n in range(count)]
    if not dock_at_f:
        return begins
    docked = ([0] if first != 0 else []) + begins + ([total_f - frames] if first != 0 else [])
    return [b for b in docked if b + frames <= total_f]
