"""Tiny stand-in for the reference's UNet wrapper in GPU tests (the reference checkout does not travel to
the GPU box).  Same CALL SIGNATURE as ``DiffusionWrapper.forward`` with ``conditioning_key='crossattn'``
(lvdm/models/ddpm3d.py:702-712) feeding ``UNetModel.forward`` (lvdm/modules/networks/openaimodel3d.py:
657-708): ``model(x[b, C, t, h, w], timesteps, c_crossattn=[ctx[b, L, D]], fps=..., **kwargs)`` where
``timesteps`` has length b*t (one embedding per frame image, the P9 trick) or b (broadcast over t, only
valid for b == 1 in the reference).  Seeded weights, a few hundred parameters; it is NOT the UNet."""
import math

import torch
import torch.nn as nn


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class StandInUNet(nn.Module):
    def __init__(self, ctx_dim=32, ch=8, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.ch = ch
        self.inp = nn.Conv2d(4, ch, 3, padding=1)
        self.temporal = nn.Conv1d(ch, ch, 3, padding=1)
        self.time = nn.Linear(16, ch)
        self.ctx = nn.Linear(ctx_dim, ch)
        self.out = nn.Conv2d(ch, 4, 3, padding=1)
        for p in self.parameters():
            p.data = torch.randn(p.shape, generator=g) * 0.2

    def forward(self, x, timesteps, c_crossattn=None, fps=None, **kwargs):
        b, c, t, h, w = x.shape
        emb = self.time(timestep_embedding(timesteps, 16))                 # [(b t), ch] or [b, ch]
        if emb.shape[0] == b:
            emb = emb.repeat_interleave(t, dim=0)
        ctx = torch.cat(c_crossattn, dim=1).mean(dim=1)                     # [b, D]
        emb = emb + self.ctx(ctx).repeat_interleave(t, dim=0)
        y = self.inp(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w))     # "(b t) c h w", like the reference
        y = torch.tanh(y + emb[:, :, None, None])
        z = y.view(b, t, self.ch, h * w).permute(0, 3, 2, 1).reshape(b * h * w, self.ch, t)
        z = self.temporal(z).view(b, h * w, self.ch, t).permute(0, 3, 2, 1).reshape(b * t, self.ch, h, w)
        y = self.out(torch.tanh(y + z))
        return y.view(b, t, 4, h, w).permute(0, 2, 1, 3, 4).contiguous()
