"""Golden vectors from the REAL planar I2V loop -> tests/golden/ring_loop_i2v.npz.

Runs the UNMODIFIED ``VC2_Pipeline_I2V_SpherePano.basic_sample_shift_multi_windows``
(pipeline/i2v_sphere_panorama_pipeline.py:565-992) imported from /root/reference on CPU -- the loop behind
BASELINE configs 1 and 5: overlapping ``RingLatent`` windows on W, H and F (frame docking on), a latent-shaped
denoised-mask panorama, ``re_noise`` + ``mix_latents_with_mask`` on the read side, the
``merge_prev_denoised_ratio`` mix on the write side and one ``RingImageTensor`` condition-image window per
latent window -- with the reduced-width reference UNet and the reference scheduler of
tests/golden/_ref_fake_t2v.py and a deterministic stand-in for the CLIP image embedder.  Recorded: every UNet
input digest + output, every condition-image window digest + its embedding, the re_noise draw count, the
latent / denoised panoramas after every step (sha256) and the returned denoised panorama.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ring_loop_i2v.py
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


def get_image_embeds(img):
    """Stand-in for embedder + image_proj_model (lvdm/models/ddpm3d.py:689-693): [b, 3, h, w] -> [b, 4, CTX]."""
    pooled = torch.nn.functional.adaptive_avg_pool2d(img.float(), (2, 2)).reshape(img.shape[0], 12)
    emb = (pooled @ proj).reshape(img.shape[0], IMG_TOKENS, rf.CTX)
    img_calls.append((sha(img.numpy()), tuple(img.shape), emb.numpy().copy()))
    return emb


fake.get_image_embeds = get_image_embeds
pipe = VC2_Pipeline_I2V_SpherePano.__new__(VC2_Pipeline_I2V_SpherePano)
pipe.pretrained_t2v, pipe.scheduler = fake, sched
pipe.model_config = {"params": {"unet_config": {"params": rf.UNET}}}
pipe.vae_scale_factor = 8

CFG = dict(frames=4, height=64, width=128, total_h=128, total_w=384, total_f=8, nw=3, nh=3, loop_step=4, steps=3,
           loop_step_frame=2, begin_index_offset=1)
OV_F, RATIO, PREV = 0.5, 0.6, [0.3, 0.5, 0.7]
F, H, W = CFG["total_f"], CFG["total_h"] // 8, CFG["total_w"] // 8
lat0 = randn(81, 1, 4, F, H, W)
rng = np.random.RandomState(82)
image_u8 = rng.randint(0, 256, size=(CFG["total_h"], CFG["total_w"], 3), dtype=np.uint8)
tmp = tempfile.mkdtemp()
pano_png = os.path.join(tmp, "pano.png")
Image.fromarray(image_u8).save(pano_png)

per_step = []


class SnapBar:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self, *a):                        # once per DDIM step, from inside the loop (:970)
        loc = sys._getframe(1).f_locals
        per_step.append((sha(loc["panorama_ring_latent_handler"].torch_latent.numpy()),
                         sha(loc["panorama_ring_latent_denoised_handler"].torch_latent.numpy())))


pipe.progress_bar = lambda *a, **k: SnapBar()
draws = {"n": 0}
real_randn, real_randn_like = torch.randn, torch.randn_like


def seeded_randn_like(x, **k):
    draws["n"] += 1
    return torch.from_numpy(randn(6000 + draws["n"], *x.shape))


def zero_randn(*shape, **k):                    # ddim_step: sigma = 0 (eta 0) multiplies this draw
    if len(shape) == 1 and not isinstance(shape[0], int):
        shape = tuple(shape[0])
    return torch.zeros(shape)


ref_sched.torch.randn_like = seeded_randn_like
ref_sched.torch.randn = zero_randn
try:
    with rf.quiet():
        videos, denoised = pipe.basic_sample_shift_multi_windows(
            prompt="a planar i2v test panorama", img_cond_path=pano_png, height=CFG["height"], width=CFG["width"],
            frames=CFG["frames"], fps=16, guidance_scale=7.5, init_panorama_latent=torch.from_numpy(lat0),
            total_w=CFG["total_w"], total_h=CFG["total_h"], total_f=CFG["total_f"], num_windows_w=CFG["nw"],
            num_windows_h=CFG["nh"], num_windows_f=2, loop_step=CFG["loop_step"],
            begin_index_offset=CFG["begin_index_offset"], dock_at_f=True, overlap_ratio_list_f=[OV_F] * CFG["steps"],
            loop_step_frame=CFG["loop_step_frame"], pano_image_path=pano_png, num_inference_steps=CFG["steps"],
            output_type="latent", merge_renoised_overlap_latent_ratio=RATIO, merge_prev_denoised_ratio_list=PREV,
            skip_time_step_idx=0)
finally:
    ref_sched.torch.randn_like, ref_sched.torch.randn = real_randn_like, real_randn

n_win = len(calls) // 2
assert len(calls) == 2 * n_win and len(per_step) == CFG["steps"]
assert len(img_calls) == 1 + n_win                       # the global cond image (:629), then one window per latent window
assert all(np.array_equal(a[0], b[0]) for a, b in zip(calls[0::2], calls[1::2]))
image_f32 = (torch.from_numpy(image_u8.astype(np.float32)).permute(2, 0, 1) / 255. - 0.5) * 2   # tensor_utils.py:7-16
out = {"cfg": np.array([CFG[k] for k in ("frames", "height", "width", "total_h", "total_w", "total_f", "nw", "nh",
                                         "loop_step", "steps", "loop_step_frame", "begin_index_offset")], dtype=np.int64),
       "overlap_f": np.array(OV_F), "ratio": np.array(RATIO), "prev_ratios": np.array(PREV), "guidance": np.array(7.5),
       "image_u8": image_u8, "image_f32_sha": sha(image_f32.numpy()),
       "final_denoised": denoised.numpy(), "step_latent_sha": np.stack([p[0] for p in per_step]),
       "step_denoised_sha": np.stack([p[1] for p in per_step]),
       "unet_in_sha": np.stack([c[0] for c in calls[0::2]]),
       "eps_cond": np.stack([c[2] for c in calls[0::2]]), "eps_uncond": np.stack([c[2] for c in calls[1::2]]),
       "img_win_sha": np.stack([c[0] for c in img_calls[1:]]), "img_emb": np.stack([c[2] for c in img_calls[1:]]),
       "renoise_draws": np.array(draws["n"]),
       "text_emb": rf.text_embed(["a planar i2v test panorama"]).numpy(), "uncond_emb": rf.text_embed([""]).numpy()}
out.update(rf.schedule_arrays(sched))
np.savez_compressed(os.path.join(HERE, "ring_loop_i2v.npz"), **out)
print("ring_loop_i2v.npz:", os.path.getsize(os.path.join(HERE, "ring_loop_i2v.npz")) // 1024, "KiB;", len(calls),
      "UNet calls,", len(img_calls) - 1, "image windows,", draws["n"], "re_noise draws")
