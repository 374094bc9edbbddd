"""Golden vectors from the REAL planar shift-window loop -> tests/golden/ring_loop.npz.

Runs the UNMODIFIED ``VC2_Pipeline_T2V.basic_sample_shift_multi_windows``
(pipeline/t2v_normal_pipeline.py:214-570) imported from /root/reference on CPU -- the caller of the
``RingLatent`` window path -- with the reduced-width reference UNet and the reference
scheduler of tests/golden/_ref_fake_t2v.py, a given ``init_panorama_latent``, 3 x 2 x 2 windows
(w x h x f), ``shift_jump_odd_w`` on, and ``docking_w`` at the last step so that windows overlap there
(the sequential read-after-write order matters).  Recorded: every UNet input digest + output, the
latent / denoised panoramas after every step (sha256, read from the running loop's locals by the
progress-bar stub) and at the end (arrays).  The GPU tests replay the recorded UNet outputs through
dynamicscaler_b200's ring gather / fused merge and must land on the same panoramas bit for bit.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ring_loop.py
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_fake_t2v as rf  # noqa: E402
from _ref_fake_t2v import randn, sha  # noqa: E402

with rf.quiet():
    from pipeline.t2v_normal_pipeline import VC2_Pipeline_T2V  # noqa: E402

wrapper, calls, fake, sched = rf.build()
pipe = VC2_Pipeline_T2V.__new__(VC2_Pipeline_T2V)
pipe.pretrained_t2v, pipe.scheduler = fake, sched
pipe.model_config = {"params": {"unet_config": {"params": rf.UNET}}}
pipe.vae_scale_factor = 8

CFG = dict(frames=4, h=8, w=12, nw=3, nh=2, nf=2, loop_step=4, steps=3, guidance=7.5, prompt="a planar test panorama")
DOCK_STEPS = [2]
lat0 = randn(51, 1, 4, CFG["frames"] * CFG["nf"], CFG["h"] * CFG["nh"], CFG["w"] * CFG["nw"])

per_step = []


class SnapBar:
    """Stands in for diffusers' progress bar; ``update()`` is called once per DDIM step from inside the
    loop (:555), where the two RingLatent handlers are locals of the caller."""

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):
        loc = sys._getframe(1).f_locals
        per_step.append((sha(loc["panorama_ring_latent_handler"].torch_latent.numpy()),
                         sha(loc["panorama_ring_latent_denoised_handler"].torch_latent.numpy())))


pipe.progress_bar = lambda *a, **k: SnapBar()
real_randn = torch.randn
import pipeline.scheduler as ref_sched  # noqa: E402


def zero_randn(*shape, **k):                    # ddim_step: sigma = 0 (eta 0) multiplies this draw
    if len(shape) == 1 and not isinstance(shape[0], int):
        shape = tuple(shape[0])
    return torch.zeros(shape)


ref_sched.torch.randn = zero_randn
try:
    with rf.quiet():
        videos, denoised = pipe.basic_sample_shift_multi_windows(
            prompt=CFG["prompt"], height=CFG["h"] * 8, width=CFG["w"] * 8, frames=CFG["frames"], fps=16,
            guidance_scale=CFG["guidance"], init_panorama_latent=torch.from_numpy(lat0),
            num_windows_w=CFG["nw"], num_windows_h=CFG["nh"], num_windows_f=CFG["nf"], loop_step=CFG["loop_step"],
            num_inference_steps=CFG["steps"], output_type="latent", shift_jump_odd_w=True,
            docking_w=True, docking_step_range=DOCK_STEPS,
            skip_time_step_idx=0)          # only printed (:306) when use_skip_time is off, but must be an int
finally:
    ref_sched.torch.randn = real_randn

n_calls = len(calls)
assert n_calls % 2 == 0 and len(per_step) == CFG["steps"]
out = {"cfg": np.array([CFG[k] for k in ("frames", "h", "w", "nw", "nh", "nf", "loop_step", "steps")], dtype=np.int64),
       "guidance": np.array(CFG["guidance"]), "dock_steps": np.array(DOCK_STEPS, dtype=np.int64),
       "final_denoised": denoised.numpy(), "step_latent_sha": np.stack([p[0] for p in per_step]),
       "step_denoised_sha": np.stack([p[1] for p in per_step]),
       "unet_in_sha": np.stack([c[0] for c in calls[0::2]]),
       "eps_cond": np.stack([c[2] for c in calls[0::2]]), "eps_uncond": np.stack([c[2] for c in calls[1::2]]),
       "text_emb": rf.text_embed([CFG["prompt"]]).numpy(), "uncond_emb": rf.text_embed([""]).numpy()}
assert all(np.array_equal(a[0], b[0]) for a, b in zip(calls[0::2], calls[1::2]))      # both branches see the same window
out.update(rf.schedule_arrays(sched))
np.savez_compressed(os.path.join(HERE, "ring_loop.npz"), **out)
print("ring_loop.npz:", os.path.getsize(os.path.join(HERE, "ring_loop.npz")) // 1024, "KiB;", n_calls, "UNet calls,",
      len(per_step), "steps")
