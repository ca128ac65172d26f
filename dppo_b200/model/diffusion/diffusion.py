"""
Base diffusion policy (host side): owns the denoiser module and the schedule tables.

DiffusionModel -> /root/reference/dppo/model/diffusion/diffusion.py:28-196 (constructor keywords, attribute names and
table arithmetic are the reference's, so `state_dict()` and the tables are bit-identical; the tables are handed to
libdppo_b200.so as the `dppo_sched_desc` of include/dppo_b200.h).  Sampling lives in the VPG subclass, as in the
reference (the base `forward` there is not callable on its own).
"""

import logging
from collections import namedtuple

import torch
from torch import nn

from dppo_b200.model.diffusion.sampling import cosine_beta_schedule

log = logging.getLogger(__name__)
Sample = namedtuple("Sample", "trajectories chains")


class DiffusionModel(nn.Module):
    def __init__(
        self,
        network,
        horizon_steps,
        obs_dim,
        action_dim,
        network_path=None,
        device="cuda:0",
        denoised_clip_value=1.0,
        randn_clip_value=10,
        final_action_clip_value=None,
        eps_clip_value=None,
        denoising_steps=100,
        predict_epsilon=True,
        use_ddim=False,
        ddim_discretize="uniform",
        ddim_steps=None,
        **kwargs,
    ):
        super().__init__()
        self.device = device
        self.horizon_steps = horizon_steps
        self.obs_dim = obs_dim
        self.action_dim = action_dim
        self.denoising_steps = int(denoising_steps)
        self.predict_epsilon = predict_epsilon
        self.use_ddim = use_ddim
        self.ddim_steps = ddim_steps
        self.denoised_clip_value = denoised_clip_value
        self.final_action_clip_value = final_action_clip_value
        self.randn_clip_value = randn_clip_value
        self.eps_clip_value = eps_clip_value

        self.network = network.to(device)
        if network_path is not None:
            ckpt = torch.load(network_path, map_location=device, weights_only=True)
            key = "ema" if "ema" in ckpt else "model"  # "ema": supervised pre-training, "model": RL checkpoint
            self.load_state_dict(ckpt[key], strict=False)
            log.info("Loaded %s weights from %s", key, network_path)

        self._build_ddpm_tables(device)
        if use_ddim:
            if not predict_epsilon:
                raise ValueError("DDIM requires predicting epsilon")
            if ddim_discretize != "uniform":
                raise ValueError(f"unknown DDIM discretisation {ddim_discretize!r}")
            self._build_ddim_tables()

    # fp32 tensor arithmetic in the reference's order of operations (diffusion.py:98-148)
    def _build_ddpm_tables(self, device):
        self.betas = cosine_beta_schedule(self.denoising_steps).to(device)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, axis=0)
        self.alphas_cumprod_prev = torch.cat([torch.ones(1).to(device), self.alphas_cumprod[:-1]])
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.ddpm_var = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.ddpm_logvar_clipped = torch.log(torch.clamp(self.ddpm_var, min=1e-20))
        self.ddpm_mu_coef1 = self.betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.ddpm_mu_coef2 = (1.0 - self.alphas_cumprod_prev) * torch.sqrt(self.alphas) / (1.0 - self.alphas_cumprod)

    # "leading" spacing, then everything flipped so that index 0 is the noisiest step (diffusion.py:155-196)
    def _build_ddim_tables(self):
        ratio = self.denoising_steps // self.ddim_steps
        t = torch.arange(0, self.ddim_steps, device=self.device) * ratio
        a = self.alphas_cumprod[t].clone().to(torch.float32)
        a_prev = torch.cat([torch.tensor([1.0]).to(torch.float32).to(self.device), self.alphas_cumprod[t[:-1]]])
        s1m = (1.0 - a) ** 0.5
        sig = 0 * ((1 - a_prev) / (1 - a) * (1 - a / a_prev)) ** 0.5
        self.ddim_t = torch.flip(t, [0])
        self.ddim_alphas = torch.flip(a, [0])
        self.ddim_alphas_sqrt = torch.flip(torch.sqrt(a), [0])
        self.ddim_alphas_prev = torch.flip(a_prev, [0])
        self.ddim_sqrt_one_minus_alphas = torch.flip(s1m, [0])
        self.ddim_sigmas = torch.flip(sig, [0])
