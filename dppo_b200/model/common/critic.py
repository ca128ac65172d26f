"""
State-value critic (host-side parameter container + autograd forward).

CriticObs -> /root/reference/dppo/model/common/critic.py:15-54  (keys `Q1.*`)
"""

from typing import Union

import torch

from dppo_b200.model.common.mlp import MLP, ResidualMLP


class CriticObs(torch.nn.Module):
    """V(s): state (B, To, Do) flattened -> (B, 1)."""

    def __init__(self, cond_dim, mlp_dims, activation_type="Mish", use_layernorm=False, residual_style=False, **kwargs):
        super().__init__()
        dims = [cond_dim] + list(mlp_dims) + [1]
        net = ResidualMLP if residual_style else MLP
        self.Q1 = net(
            dims,
            activation_type=activation_type,
            out_activation_type="Identity",
            use_layernorm=use_layernorm,
        )

    def forward(self, cond: Union[dict, torch.Tensor]):
        if isinstance(cond, dict):
            state = cond["state"]
            state = state.reshape(len(state), -1)
        else:
            state = cond
        return self.Q1(state)
