"""
MLP denoiser (host-side parameter container + autograd forward for the update path).

DiffusionMLP -> /root/reference/dppo/model/diffusion/mlp_diffusion.py:174-250
(keys `time_embedding.{1,3}.*`, `cond_mlp.moduleList.*`, `mlp_mean.layers.*`)
"""

import torch
from torch import nn

from dppo_b200.model.common.mlp import MLP, ResidualMLP
from dppo_b200.model.diffusion.modules import SinusoidalPosEmb


class DiffusionMLP(nn.Module):
    def __init__(
        self,
        action_dim,
        horizon_steps,
        cond_dim,
        time_dim=16,
        mlp_dims=[256, 256],
        cond_mlp_dims=None,
        activation_type="Mish",
        out_activation_type="Identity",
        use_layernorm=False,
        residual_style=False,
    ):
        super().__init__()
        flat_action = action_dim * horizon_steps
        self.time_embedding = nn.Sequential(
            SinusoidalPosEmb(time_dim),
            nn.Linear(time_dim, time_dim * 2),
            nn.Mish(),
            nn.Linear(time_dim * 2, time_dim),
        )
        if cond_mlp_dims is not None:
            self.cond_mlp = MLP(
                [cond_dim] + list(cond_mlp_dims),
                activation_type=activation_type,
                out_activation_type="Identity",
            )
            trunk_in = time_dim + flat_action + cond_mlp_dims[-1]
        else:
            trunk_in = time_dim + flat_action + cond_dim
        trunk = ResidualMLP if residual_style else MLP
        self.mlp_mean = trunk(
            [trunk_in] + list(mlp_dims) + [flat_action],
            activation_type=activation_type,
            out_activation_type=out_activation_type,
            use_layernorm=use_layernorm,
        )
        self.time_dim = time_dim
        # static description used by the kernel-side weight packer
        self.action_dim = action_dim
        self.horizon_steps = horizon_steps
        self.cond_dim = cond_dim
        self.mlp_dims = list(mlp_dims)
        self.cond_mlp_dims = None if cond_mlp_dims is None else list(cond_mlp_dims)
        self.activation_type = activation_type
        self.out_activation_type = out_activation_type
        self.use_layernorm = use_layernorm
        self.residual_style = residual_style

    def forward(self, x, time, cond, **kwargs):
        B, Ta, Da = x.shape
        state = cond["state"].reshape(B, -1)
        if hasattr(self, "cond_mlp"):
            state = self.cond_mlp(state)
        temb = self.time_embedding(time.reshape(B, 1)).reshape(B, self.time_dim)
        out = self.mlp_mean(torch.cat([x.reshape(B, -1), temb, state], dim=-1))
        return out.reshape(B, Ta, Da)
