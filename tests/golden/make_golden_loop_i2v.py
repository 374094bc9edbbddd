"""Golden vectors from the REAL I2V sphere loop -> tests/golden/loop_i2v.npz.

Runs the UNMODIFIED ``VC2_Pipeline_I2V_SpherePano.basic_sample_shift_shpere_panorama``
(pipeline/i2v_sphere_panorama_pipeline.py:32-495) imported from /root/reference on CPU: sphere views over
wrapping frame windows (``RingPanoramaLatentProxy``, the temporal loop), a latent-shaped denoised-mask
panorama, ``re_noise`` + ``mix_latents_with_mask`` on the read side, the ``merge_prev_denoised_ratio`` mix
on the write side and one ``PanoramaTensor`` condition-image view per latent view -- with the reduced-width
reference UNet / scheduler of tests/golden/_ref_fake_t2v.py and a deterministic stand-in for the CLIP
image embedder.  Recorded: every UNet input digest + output, every condition-image view digest + its
embedding, the re_noise draw count and the returned latent / denoised panoramas.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_loop_i2v.py
"""
import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_fake_t2v as rf  # noqa: E402
from _ref_fake_t2v import randn, sha  # noqa: E402

with rf.quiet():
    from pipeline.i2v_sphere_panorama_pipeline import VC2_Pipeline_I2V_SpherePano  # noqa: E402
    import pipeline.scheduler as ref_sched  # noqa: E402

wrapper, calls, fake, sched = rf.build()
IMG_TOKENS = 4
proj = torch.from_numpy(randn(77, 12, IMG_TOKENS * rf.CTX)) * 0.5
img_calls = []


def get_image_embeds(batch_imgs):
    """Stand-in for embedder + image_proj_model (lvdm/models/ddpm3d.py:689-693): [b, 3, h, w] -> [b, 4, CTX]."""
    pooled = torch.nn.functional.adaptive_avg_pool2d(batch_imgs.float(), (2, 2)).reshape(batch_imgs.shape[0], 12)
    emb = (pooled @ proj).reshape(batch_imgs.shape[0], IMG_TOKENS, rf.CTX)
    img_calls.append((sha(batch_imgs.contiguous().numpy()), tuple(batch_imgs.shape), emb.numpy().copy()))
    return emb


fake.get_image_embeds = get_image_embeds
pipe = VC2_Pipeline_I2V_SpherePano.__new__(VC2_Pipeline_I2V_SpherePano)
pipe.pretrained_t2v, pipe.scheduler = fake, sched
pipe.model_config = {"params": {"unet_config": {"params": rf.UNET}}}
pipe.vae_scale_factor = 8

CFG = dict(frames=4, total_f=8, H=16, W=32, h=8, w=12, steps=3, fov=120, loop_step_frame=2, loop_step_theta=10)
PHI_THETA = {90: [0], 0: [0, 120, 240], -45: [60, 300]}
OV_F, RATIO, PREV = 0.5, 0.6, [0.3, 0.5, 0.7]
lat0 = randn(101, 1, 4, CFG["total_f"], CFG["H"], CFG["W"])
rng = np.random.RandomState(102)
image_u8 = rng.randint(0, 256, size=(CFG["H"] * 8, CFG["W"] * 8, 3), dtype=np.uint8)
tmp = tempfile.mkdtemp()
pano_png = os.path.join(tmp, "pano.png")
Image.fromarray(image_u8).save(pano_png)

draws = {"n": 0}
real_randn, real_randn_like = torch.randn, torch.randn_like


def seeded_randn_like(x, **k):
    draws["n"] += 1
    return torch.from_numpy(randn(8000 + draws["n"], *x.shape))


def zero_randn(*shape, **k):                    # ddim_step: sigma = 0 (eta 0) multiplies this draw
    if len(shape) == 1 and not isinstance(shape[0], int):
        shape = tuple(shape[0])
    return torch.zeros(shape)


ref_sched.torch.randn_like = seeded_randn_like
ref_sched.torch.randn = zero_randn
try:
    with rf.quiet():
        final_latents, denoised = pipe.basic_sample_shift_shpere_panorama(
            prompt="an i2v sphere test panorama", img_cond_path=pano_png, height=CFG["h"] * 8, width=CFG["w"] * 8,
            frames=CFG["frames"], fps=16, guidance_scale=7.5, init_sphere_latent=torch.from_numpy(lat0),
            pano_image_path=pano_png, total_f=CFG["total_f"], dock_at_f=False, overlap_ratio_list_f=[OV_F] * CFG["steps"],
            loop_step_frame=CFG["loop_step_frame"], equirect_width=CFG["W"] * 8, equirect_height=CFG["H"] * 8,
            phi_theta_dict=PHI_THETA, view_fov=CFG["fov"], loop_step_theta=CFG["loop_step_theta"],
            merge_renoised_overlap_latent_ratio=RATIO, merge_prev_denoised_ratio_list=PREV,
            num_inference_steps=CFG["steps"], output_type="latent")
finally:
    ref_sched.torch.randn_like, ref_sched.torch.randn = real_randn_like, real_randn

n_win = len(calls) // 2
assert len(calls) == 2 * n_win and len(img_calls) == 1 + n_win
assert all(np.array_equal(a[0], b[0]) for a, b in zip(calls[0::2], calls[1::2]))
image_f32 = (torch.from_numpy(image_u8.astype(np.float32)).permute(2, 0, 1) / 255. - 0.5) * 2   # tensor_utils.py:7-16
out = {"cfg": np.array([CFG[k] for k in ("frames", "total_f", "H", "W", "h", "w", "steps", "fov", "loop_step_frame",
                                         "loop_step_theta")], dtype=np.int64),
       "views": np.array([(ph, th) for ph, ths in PHI_THETA.items() for th in ths], dtype=np.float64),
       "overlap_f": np.array(OV_F), "ratio": np.array(RATIO), "prev_ratios": np.array(PREV), "guidance": np.array(7.5),
       "image_u8": image_u8, "image_f32_sha": sha(image_f32.numpy()),
       "final_latent": final_latents.numpy(), "final_denoised": denoised.numpy(),
       "unet_in_sha": np.stack([c[0] for c in calls[0::2]]),
       "eps_cond": np.stack([c[2] for c in calls[0::2]]), "eps_uncond": np.stack([c[2] for c in calls[1::2]]),
       "img_view_sha": np.stack([c[0] for c in img_calls[1:]]), "img_emb": np.stack([c[2] for c in img_calls[1:]]),
       "renoise_draws": np.array(draws["n"]),
       "text_emb": rf.text_embed(["an i2v sphere test panorama"]).numpy(), "uncond_emb": rf.text_embed([""]).numpy()}
out.update(rf.schedule_arrays(sched))
np.savez_compressed(os.path.join(HERE, "loop_i2v.npz"), **out)
print("loop_i2v.npz:", os.path.getsize(os.path.join(HERE, "loop_i2v.npz")) // 1024, "KiB;", len(calls), "UNet calls,",
      len(img_calls) - 1, "image views,", draws["n"], "re_noise draws")
