"""Golden vectors for the fused denoise epilogue (CFG combine + ddim_step), generated with the
reference's own pipeline/scheduler.py (lvdm_DDIM_Scheduler.ddim_step) imported from
/root/reference (build container only) -> tests/golden/ddim.npz.
The scheduler's internal torch.randn draws (Q6) are replaced by injected tensors so the eta > 0
case is reproducible.  Run: PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ddim.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
from conftest import randn  # noqa: E402
from lvdm.models.utils_diffusion import make_beta_schedule  # noqa: E402  (the reference)
import pipeline.scheduler as ref_sched  # noqa: E402  (the reference)

betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)   # configs/inference_t2v_512_v2.0.yaml:4-7
alphas = 1.0 - betas
ac = np.cumprod(alphas, axis=0)
model = types.SimpleNamespace(num_timesteps=1000, betas=torch.tensor(betas, dtype=torch.float32),
                              alphas_cumprod=torch.tensor(ac, dtype=torch.float32),
                              alphas_cumprod_prev=torch.tensor(np.append(1.0, ac[:-1]), dtype=torch.float32),
                              use_scale=False, device=torch.device("cpu"))
torch.set_num_threads(1)
out = {}
shape = (1, 4, 3, 6, 10)
cases = []
for n, (eta, index, g) in enumerate([(0.0, 49, 7.5), (0.0, 10, 12.0), (0.0, 0, 1.0), (0.5, 30, 7.5)]):
    sch = ref_sched.lvdm_DDIM_Scheduler(model)
    sch.make_schedule(50, ddim_eta=eta, verbose=False)
    x = torch.from_numpy(randn(100 + n, *shape))
    ec = torch.from_numpy(randn(200 + n, *shape))
    eu = torch.from_numpy(randn(300 + n, *shape))
    noise = torch.from_numpy(randn(400 + n, *shape))
    model_pred = eu + g * (ec - eu) if g != 1.0 else ec          # t2v_sphere_panorama_pipeline.py:256-261
    frames = iter([noise[:, :, [i]] for i in range(shape[2])])
    real_randn = torch.randn
    ref_sched.torch.randn = lambda *a, **k: next(frames)         # inject epsilon
    try:
        x_prev, pred_x0 = sch.ddim_step(sample=x, noise_pred=model_pred, indices=[index] * shape[2])
    finally:
        ref_sched.torch.randn = real_randn
    out[f"x_prev/{n}"] = x_prev.numpy()
    out[f"pred_x0/{n}"] = pred_x0.numpy()
    cases.append([eta, index, g, float(sch.ddim_alphas[index]), float(sch.ddim_alphas_prev[index]),
                  float(sch.ddim_sigmas[index]), float(sch.ddim_sqrt_one_minus_alphas[index])])
out["cases"] = np.array(cases, dtype=np.float64)      # eta, index, g, alpha_t, alpha_prev, sigma_t, sqrt(1-alpha_t)
out["shape"] = np.array(shape)
np.savez_compressed(os.path.join(HERE, "ddim.npz"), **out)
print("wrote ddim.npz", out["cases"])
