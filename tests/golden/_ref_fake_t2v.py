"""What the golden generators of the denoising loops share: the reference's own ``DiffusionWrapper`` +
``UNetModel`` at REDUCED width with seeded random init, a recording wrapper around it, seeded text
embeddings and a fake ``pretrained_t2v`` namespace carrying only what the reference's sampling loops
touch.  Imports the reference from /root/reference: build container only."""
import builtins
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import randn  # noqa: E402

import _ref_stubs  # noqa: E402,F401  (registers the pytorch_lightning / diffusers stubs)

_print = builtins.print
builtins.print = lambda *a, **k: None          # the reference prints at import and once per window
try:
    from lvdm.models.utils_diffusion import make_beta_schedule  # noqa: E402
    from lvdm.models.ddpm3d import DiffusionWrapper  # noqa: E402
    import pipeline.scheduler as ref_sched  # noqa: E402
finally:
    builtins.print = _print

CTX = 32
UNET = dict(in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[2, 1], num_res_blocks=1,
            channel_mult=[1, 2], num_head_channels=16, transformer_depth=1, context_dim=CTX, use_linear=True,
            use_checkpoint=False, temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
            use_relative_position=False, use_causal_attention=False, temporal_length=4, addition_attention=True,
            fps_cond=True)


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def text_embed(prompts):
    out = []
    for pr in prompts:
        seed = int(hashlib.sha256(pr.encode()).hexdigest()[:8], 16)
        out.append(torch.from_numpy(randn(seed, 77, CTX)))
    return torch.stack(out)


class quiet:
    """Context manager: silence the reference's per-window prints."""

    def __enter__(self):
        builtins.print = lambda *a, **k: None

    def __exit__(self, *a):
        builtins.print = _print
        return False


def build():
    """-> (wrapper, calls, fake pretrained_t2v, scheduler).  ``calls`` receives one
    (input sha, input array, output array) per UNet call."""
    torch.manual_seed(0)
    torch.set_num_threads(1)
    wrapper = DiffusionWrapper({"target": "lvdm.modules.networks.openaimodel3d.UNetModel", "params": UNET}, "crossattn")
    g = torch.Generator().manual_seed(1)
    for name, p in wrapper.named_parameters():  # zero_module layers would make a random-init UNet output exactly 0
        if p.abs().max() == 0:
            p.data = torch.randn(p.shape, generator=g) * 0.05
    wrapper.eval()
    calls = []

    def recording_model(x, t, **kw):
        y = wrapper(x, t, **kw)
        calls.append((sha(x.numpy()), x.numpy().copy(), y.numpy().copy()))
        return y

    betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    ac = np.cumprod(1.0 - betas, axis=0)
    fake = types.SimpleNamespace(
        model=recording_model, first_stage_model=None, cond_stage_model=None, uncond_type="empty_seq",
        temporal_length=4, device=torch.device("cpu"), get_learned_conditioning=text_embed,
        num_timesteps=1000, betas=torch.tensor(betas, dtype=torch.float32),
        alphas_cumprod=torch.tensor(ac, dtype=torch.float32),
        alphas_cumprod_prev=torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32), use_scale=False)
    sched = ref_sched.lvdm_DDIM_Scheduler(fake)
    return wrapper, calls, fake, sched


def schedule_arrays(sched):
    out = {"ddim_timesteps": np.asarray(sched.ddim_timesteps), "alphas_cumprod": sched.alphas_cumprod.numpy()}
    for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
        out[k] = np.asarray(getattr(sched, k), dtype=np.float64)
    return out
