"""
Host-side parameter containers for the MLP families on the DPPO hot path.

These mirror the *interface* of the reference modules so that `state_dict()` keys, constructor keywords and
parameter-creation order (hence seeded initialisation) are identical:

  MLP                                 -> /root/reference/dppo/model/common/mlp.py:27-81
  ResidualMLP                         -> /root/reference/dppo/model/common/mlp.py:84-125
  TwoLayerPreActivationResNetLinear   -> /root/reference/dppo/model/common/mlp.py:128-154

The `forward` methods here are the autograd (update-path) implementation; the rollout path never calls them, it
runs the packed weights through the sm_100a denoise-chain kernels (dppo_b200/csrc/chain_*.cu).  Every Linear is a
SplitLinear: on CUDA its forward / dgrad / wgrad GEMMs run on the bf16 tensor cores with the 3-product split
(fp32-grade), see split_linear.py.
"""

from collections import OrderedDict

import torch
from torch import nn

from dppo_b200.model.common.split_linear import SplitLinear

_ACTIVATIONS = {
    "ReLU": nn.ReLU,
    "ELU": nn.ELU,
    "GELU": nn.GELU,
    "Tanh": nn.Tanh,
    "Mish": nn.Mish,
    "Identity": nn.Identity,
    "Softplus": nn.Softplus,
}


def make_activation(name: str) -> nn.Module:
    if name not in _ACTIVATIONS:
        raise KeyError(f"unknown activation_type {name!r}; expected one of {sorted(_ACTIVATIONS)}")
    return _ACTIVATIONS[name]()


class MLP(nn.Module):
    """Plain stack: Linear -> [LayerNorm] -> [Dropout] -> activation per stage (keys `moduleList.{i}.linear_1.*`)."""

    def __init__(
        self,
        dim_list,
        append_dim=0,
        append_layers=None,
        activation_type="Tanh",
        out_activation_type="Identity",
        use_layernorm=False,
        use_layernorm_final=False,
        dropout=0,
        use_drop_final=False,
        verbose=False,
    ):
        super().__init__()
        self.append_layers = append_layers
        self.activation_type = activation_type
        self.out_activation_type = out_activation_type
        stages = []
        n_stage = len(dim_list) - 1
        for k in range(n_stage):
            fan_in, fan_out = dim_list[k], dim_list[k + 1]
            if append_dim > 0 and k in append_layers:
                fan_in += append_dim
            last = k == n_stage - 1
            parts = OrderedDict()
            parts["linear_1"] = SplitLinear(fan_in, fan_out)
            if use_layernorm and (not last or use_layernorm_final):
                parts["norm_1"] = nn.LayerNorm(fan_out)
            if dropout > 0 and (not last or use_drop_final):
                parts["dropout_1"] = nn.Dropout(dropout)
            parts["act_1"] = make_activation(out_activation_type if last else activation_type)
            stages.append(nn.Sequential(parts))
        self.moduleList = nn.ModuleList(stages)

    def forward(self, x, append=None):
        for k, stage in enumerate(self.moduleList):
            if append is not None and k in self.append_layers:
                x = torch.cat((x, append), dim=-1)
            x = stage(x)
        return x


class TwoLayerPreActivationResNetLinear(nn.Module):
    """h + l2(act(norm2(l1(act(norm1(h)))))) with optional LayerNorm(eps=1e-6); keys `l1.*`, `l2.*`, `norm1.*`, `norm2.*`."""

    def __init__(self, hidden_dim, activation_type="Mish", use_layernorm=False, dropout=0):
        super().__init__()
        self.l1 = SplitLinear(hidden_dim, hidden_dim)
        self.l2 = SplitLinear(hidden_dim, hidden_dim)
        self.act = make_activation(activation_type)
        if use_layernorm:
            self.norm1 = nn.LayerNorm(hidden_dim, eps=1e-06)
            self.norm2 = nn.LayerNorm(hidden_dim, eps=1e-06)
        if dropout > 0:
            raise NotImplementedError("Dropout not implemented for residual MLP!")

    def forward(self, h):
        y = self.norm1(h) if hasattr(self, "norm1") else h
        y = self.l1(self.act(y))
        if hasattr(self, "norm2"):
            y = self.norm2(y)
        y = self.l2(self.act(y))
        return y + h


class ResidualMLP(nn.Module):
    """Linear(in,H) -> n pre-activation residual blocks -> Linear(H,out) [-> LayerNorm] -> out activation (keys `layers.{i}.*`)."""

    def __init__(
        self,
        dim_list,
        activation_type="Mish",
        out_activation_type="Identity",
        use_layernorm=False,
        use_layernorm_final=False,
        dropout=0,
    ):
        super().__init__()
        hidden = dim_list[1]
        n_hidden_linear = len(dim_list) - 3
        assert n_hidden_linear % 2 == 0
        self.hidden_dim = hidden
        self.activation_type = activation_type
        self.use_layernorm = use_layernorm
        seq = [SplitLinear(dim_list[0], hidden)]
        for _ in range(n_hidden_linear // 2):
            seq.append(
                TwoLayerPreActivationResNetLinear(
                    hidden_dim=hidden,
                    activation_type=activation_type,
                    use_layernorm=use_layernorm,
                    dropout=dropout,
                )
            )
        seq.append(SplitLinear(hidden, dim_list[-1]))
        if use_layernorm_final:
            seq.append(nn.LayerNorm(dim_list[-1]))
        seq.append(make_activation(out_activation_type))
        self.layers = nn.ModuleList(seq)
        self.n_blocks = n_hidden_linear // 2

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x
