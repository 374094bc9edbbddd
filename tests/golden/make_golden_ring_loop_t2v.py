"""Golden vectors from the REAL planar T2V loop of the sphere pipeline -> tests/golden/ring_loop_t2v.npz.

Runs the UNMODIFIED ``VC2_Pipeline_T2V_SpherePano.basic_sample_shift_multi_windows``
(pipeline/t2v_sphere_panorama_pipeline.py:322-655; stages 2 / 3 of the text-to-panorama run) imported from
/root/reference on CPU: overlapping ``RingLatent`` windows on W and H with integer strides, ``dock_at_h``
windows at the top / bottom edge, 2 frame windows shifted per step, a latent-shaped mask panorama whose
``[0, 0, [0]]`` plane drives the ``re_noise`` + ``mix_latents_with_mask`` mix, with the reduced-width reference
UNet and the reference scheduler of tests/golden/_ref_fake_t2v.py.  Recorded: every UNet input digest + output,
the re_noise draw count, the latent / denoised panoramas after every step (sha256) and the returned denoised
panorama.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ring_loop_t2v.py
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
    from pipeline.t2v_sphere_panorama_pipeline import VC2_Pipeline_T2V_SpherePano  # noqa: E402
    import pipeline.scheduler as ref_sched  # noqa: E402

wrapper, calls, fake, sched = rf.build()
pipe = VC2_Pipeline_T2V_SpherePano.__new__(VC2_Pipeline_T2V_SpherePano)
pipe.pretrained_t2v, pipe.scheduler = fake, sched
pipe.model_config = {"params": {"unet_config": {"params": rf.UNET}}}
pipe.vae_scale_factor = 8

CFG = dict(frames=4, height=64, width=96, total_h=128, total_w=240, nw=3, nh=3, nf=2, loop_step=4, steps=3)
RATIO = 0.6
F, H, W = CFG["frames"] * CFG["nf"], CFG["total_h"] // 8, CFG["total_w"] // 8
lat0 = randn(95, 1, 4, F, H, W)
per_step = []


class SnapBar:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):                        # once per DDIM step, from inside the loop (:634)
        loc = sys._getframe(1).f_locals
        per_step.append((sha(loc["panorama_ring_latent_handler"].torch_latent.numpy()),
                         sha(loc["panorama_ring_latent_denoised_handler"].torch_latent.numpy())))


pipe.progress_bar = lambda *a, **k: SnapBar()
draws = {"n": 0}
real_randn, real_randn_like = torch.randn, torch.randn_like


def seeded_randn_like(x, **k):
    draws["n"] += 1
    return torch.from_numpy(randn(7000 + draws["n"], *x.shape))


def zero_randn(*shape, **k):                    # ddim_step: sigma = 0 (eta 0) multiplies this draw
    if len(shape) == 1 and not isinstance(shape[0], int):
        shape = tuple(shape[0])
    return torch.zeros(shape)


ref_sched.torch.randn_like = seeded_randn_like
ref_sched.torch.randn = zero_randn
try:
    with rf.quiet():
        videos, denoised = pipe.basic_sample_shift_multi_windows(
            prompt="a planar t2v test panorama", height=CFG["height"], width=CFG["width"], frames=CFG["frames"], fps=16,
            guidance_scale=7.5, init_panorama_latent=torch.from_numpy(lat0), total_w=CFG["total_w"],
            total_h=CFG["total_h"], num_windows_w=CFG["nw"], num_windows_h=CFG["nh"], num_windows_f=CFG["nf"],
            loop_step=CFG["loop_step"], dock_at_h=True, num_inference_steps=CFG["steps"], output_type="latent",
            merge_renoised_overlap_latent_ratio=RATIO, skip_time_step_idx=0)
finally:
    ref_sched.torch.randn_like, ref_sched.torch.randn = real_randn_like, real_randn

n_win = len(calls) // 2
assert len(calls) == 2 * n_win and len(per_step) == CFG["steps"]
assert all(np.array_equal(a[0], b[0]) for a, b in zip(calls[0::2], calls[1::2]))
out = {"cfg": np.array([CFG[k] for k in ("frames", "height", "width", "total_h", "total_w", "nw", "nh", "nf", "loop_step",
                                         "steps")], dtype=np.int64),
       "ratio": np.array(RATIO), "guidance": np.array(7.5), "final_denoised": denoised.numpy(),
       "step_latent_sha": np.stack([p[0] for p in per_step]), "step_denoised_sha": np.stack([p[1] for p in per_step]),
       "unet_in_sha": np.stack([c[0] for c in calls[0::2]]),
       "eps_cond": np.stack([c[2] for c in calls[0::2]]), "eps_uncond": np.stack([c[2] for c in calls[1::2]]),
       "renoise_draws": np.array(draws["n"]),
       "text_emb": rf.text_embed(["a planar t2v test panorama"]).numpy(), "uncond_emb": rf.text_embed([""]).numpy()}
out.update(rf.schedule_arrays(sched))
np.savez_compressed(os.path.join(HERE, "ring_loop_t2v.npz"), **out)
print("ring_loop_t2v.npz:", os.path.getsize(os.path.join(HERE, "ring_loop_t2v.npz")) // 1024, "KiB;", len(calls),
      "UNet calls,", draws["n"], "re_noise draws")
