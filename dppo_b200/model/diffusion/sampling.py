"""
Schedule helpers (host side).

cosine_beta_schedule / extract / make_timesteps -> /root/reference/dppo/model/diffusion/sampling.py:10-31
"""

import numpy as np
import torch


def cosine_beta_schedule(timesteps, s=0.008, dtype=torch.float32):
    """Cosine noise schedule evaluated in float64 numpy, returned as a `dtype` tensor (clipped to [0, 0.999])."""
    n = timesteps + 1
    grid = np.linspace(0, n, n)
    abar = np.cos(((grid / n) + s) / (1 + s) * np.pi * 0.5) ** 2
    abar = abar / abar[0]
    betas = 1 - (abar[1:] / abar[:-1])
    return torch.tensor(np.clip(betas, a_min=0, a_max=0.999), dtype=dtype)


def extract(table, t, x_shape):
    """table[t] reshaped to broadcast against a tensor of shape `x_shape` (batch first)."""
    return table.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def make_timesteps(batch_size, i, device):
    return torch.full((batch_size,), i, device=device, dtype=torch.long)
