"""Golden vectors from the REAL reference denoising loop -> tests/golden/loop.npz.

Runs the UNMODIFIED ``VC2_Pipeline_T2V_SpherePano.basic_sample_shift_shpere_panorama``
(pipeline/t2v_sphere_panorama_pipeline.py:23-310) imported from /root/reference on CPU, with
  * the reference's own ``DiffusionWrapper`` + ``UNetModel`` (lvdm/models/ddpm3d.py:696-712,
    lvdm/modules/networks/openaimodel3d.py) at REDUCED width, seeded random init (the zero-initialised
    output layers re-randomised),
  * the reference's own ``lvdm_DDIM_Scheduler`` (pipeline/scheduler.py) incl. ``re_noise`` and ``ddim_step``,
  * ``merge_renoised_overlap_latent_ratio`` set, so the re-noise + mask mix runs,
  * ~30 lines of stubs for what is not installed here (diffusers, pytorch_lightning; SURVEY §8c;
    tests/golden/_ref_stubs.py, model / scheduler set-up shared with the planar-loop generators in
    tests/golden/_ref_fake_t2v.py),
and records, per UNet call, the digest of the input window and the model output.  The GPU box has no
reference checkout: there the tests replay the recorded UNet outputs ("replay model") through
dynamicscaler_b200's fused gather / merge calls and must reproduce the final panoramas bit for bit.

The scheduler's internal RNG draws (re_noise: randn_like, ddim_step: randn) are replaced by
seeded tensors ``randn(seed)`` that the tests regenerate.

Also recorded (P9): the same windows pushed through the UNet as ONE batch with ``timesteps`` of length
b*t -- the call the batched driver makes -- next to the per-window results.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_loop.py
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_fake_t2v as rf  # noqa: E402  (reduced reference UNet, scheduler, fake pretrained_t2v, stubs)
from _ref_fake_t2v import randn, text_embed  # noqa: E402

with rf.quiet():
    import pipeline.scheduler as ref_sched  # noqa: E402
    from pipeline.t2v_sphere_panorama_pipeline import VC2_Pipeline_T2V_SpherePano  # noqa: E402

import builtins  # noqa: E402
_print = builtins.print
CTX = rf.CTX
wrapper, calls, fake, sched = rf.build()        # calls: (input sha, input, output) per UNet call
pipe = VC2_Pipeline_T2V_SpherePano.__new__(VC2_Pipeline_T2V_SpherePano)
pipe.pretrained_t2v, pipe.scheduler = fake, sched
pipe.model_config = {"params": {"unet_config": {"params": rf.UNET}}}
pipe.vae_scale_factor = 8

# ---- the run ----------------------------------------------------------------------------------
CFG = dict(F=4, H=64, W=128, h=16, w=24, steps=3, fov=120, guidance=7.5, ratio=0.6, loop_step_theta=10, prompt="a test panorama")
PHI_THETA = {90: [0], 0: [0, 120, 240], -45: [60, 300]}
lat0 = randn(41, 1, 4, CFG["F"], CFG["H"], CFG["W"])

draws = {"n": 0}
real_randn, real_randn_like = torch.randn, torch.randn_like


def seeded_randn_like(x, **k):
    draws["n"] += 1
    return torch.from_numpy(randn(5000 + draws["n"], *x.shape))


def zero_randn(*shape, **k):                    # ddim_step: sigma = 0 (eta 0) multiplies this draw, any finite value works
    if len(shape) == 1 and not isinstance(shape[0], int):
        shape = tuple(shape[0])
    return torch.zeros(shape)


out = {"cfg": np.array([CFG[k] for k in ("F", "H", "W", "h", "w", "steps", "fov")], dtype=np.int64),
       "guidance": np.array(CFG["guidance"]), "ratio": np.array(CFG["ratio"]),
       "views": np.array([(ph, th) for ph, ths in PHI_THETA.items() for th in ths], dtype=np.float64)}
builtins.print = lambda *a, **k: None
ref_sched.torch.randn_like = seeded_randn_like
ref_sched.torch.randn = zero_randn
try:
    videos, denoised = pipe.basic_sample_shift_shpere_panorama(
        prompt=CFG["prompt"], height=CFG["h"] * 8, width=CFG["w"] * 8, frames=CFG["F"], fps=16,
        guidance_scale=CFG["guidance"], init_sphere_latent=torch.from_numpy(lat0),
        equirect_width=CFG["W"] * 8, equirect_height=CFG["H"] * 8, phi_theta_dict=PHI_THETA, view_fov=CFG["fov"],
        loop_step_theta=CFG["loop_step_theta"], merge_renoised_overlap_latent_ratio=CFG["ratio"],
        num_inference_steps=CFG["steps"], output_type="latent")
finally:
    ref_sched.torch.randn_like, ref_sched.torch.randn = real_randn_like, real_randn
    builtins.print = _print

n_views = len(out["views"])
assert len(calls) == CFG["steps"] * n_views * 2, len(calls)
out["final_latent"] = videos.numpy()
out["final_denoised"] = denoised.numpy()
out["unet_in_sha"] = np.stack([c[0] for c in calls[0::2]])           # cond call of every window (uncond sees the same x)
out["unet_in_sample"] = np.stack([c[1].reshape(-1)[::97].copy() for c in calls[0::2]])
out["eps_cond"] = np.stack([c[2] for c in calls[0::2]])
out["eps_uncond"] = np.stack([c[2] for c in calls[1::2]])
out["renoise_draws"] = np.array(draws["n"])
out["text_emb"] = text_embed([CFG["prompt"]]).numpy()
out["uncond_emb"] = text_embed([""]).numpy()
# scheduler tables the loop driver needs (float64 numpy arrays in the reference)
out["ddim_timesteps"] = np.asarray(sched.ddim_timesteps)
out["alphas_cumprod"] = sched.alphas_cumprod.numpy()
for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
    out[k] = np.asarray(getattr(sched, k), dtype=np.float64)

# ---- P9: one batched UNet call (timesteps of length b*t) == the per-window calls ----------------
step0 = [c[1] for c in calls[0:2 * n_views:2]]
xb = torch.from_numpy(np.concatenate(step0, axis=0))               # [n, 4, F, h, w]
t0 = int(np.flip(sched.ddim_timesteps)[0])
ts = torch.full((xb.shape[0] * CFG["F"],), t0, dtype=torch.long)
ctx = text_embed([CFG["prompt"]]).expand(xb.shape[0], -1, -1)
with torch.no_grad():
    yb = wrapper(xb, ts, c_crossattn=[ctx], fps=16, curr_time_steps=ts, temporal_length=CFG["F"], clean_cond=True)
per = np.concatenate([c[2] for c in calls[0:2 * n_views:2]], axis=0)
out["p9_batched_vs_looped_maxabs"] = np.array(np.abs(yb.numpy() - per).max())
print("P9 batched vs looped max abs:", float(out["p9_batched_vs_looped_maxabs"]))
np.savez_compressed(os.path.join(HERE, "loop.npz"), **out)
print("loop.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "loop.npz")) // 1024, "KiB;",
      len(calls), "UNet calls,", draws["n"], "re_noise draws")
