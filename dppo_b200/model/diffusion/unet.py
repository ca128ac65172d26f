"""
1-D UNet denoiser with FiLM conditioning (host-side parameter container + autograd forward).

ResidualBlock1D -> /root/reference/dppo/model/diffusion/unet.py:27-118
Unet1D          -> /root/reference/dppo/model/diffusion/unet.py:121-327
(keys `time_mlp.{1,3}.*`, `down_modules.{l}.{0,1,2}.*`, `mid_modules.{0,1}.*`, `up_modules.{l}.{0,1,2}.*`,
 `final_conv.{0,1}.*`; the pixel / point-cloud variants are out of scope)
"""

import torch
from torch import nn

from dppo_b200.model.common.mlp import ResidualMLP
from dppo_b200.model.common.split_linear import SplitLinear
from dppo_b200.model.diffusion.dense_conv import DenseConv1d
from dppo_b200.model.diffusion.modules import Conv1dBlock, Downsample1d, SinusoidalPosEmb, Upsample1d


class _Col(nn.Module):
    """(B, C) -> (B, C, 1)"""

    def forward(self, x):
        return x.unsqueeze(-1)


class ResidualBlock1D(nn.Module):
    def __init__(
        self,
        in_channels,
        out_channels,
        cond_dim,
        kernel_size=5,
        n_groups=None,
        cond_predict_scale=False,
        larger_encoder=False,
        activation_type="Mish",
        groupnorm_eps=1e-5,
    ):
        super().__init__()
        mk = lambda cin: Conv1dBlock(cin, out_channels, kernel_size, n_groups=n_groups,
                                     activation_type=activation_type, eps=groupnorm_eps)
        self.blocks = nn.ModuleList([mk(in_channels), mk(out_channels)])
        if activation_type == "Mish":
            act = nn.Mish()
        elif activation_type == "ReLU":
            act = nn.ReLU()
        else:
            raise ValueError("Unknown activation type for ResidualBlock1D")
        film = out_channels * 2 if cond_predict_scale else out_channels
        self.cond_predict_scale = cond_predict_scale
        self.out_channels = out_channels
        if larger_encoder:
            self.cond_encoder = nn.Sequential(
                SplitLinear(cond_dim, film), act, SplitLinear(film, film), act, SplitLinear(film, film), _Col()
            )
        else:
            self.cond_encoder = nn.Sequential(act, SplitLinear(cond_dim, film), _Col())
        self.residual_conv = DenseConv1d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()

    def forward(self, x, cond):
        y = self.blocks[0](x)
        film = self.cond_encoder(cond)
        if self.cond_predict_scale:
            film = film.reshape(film.shape[0], 2, self.out_channels, 1)
            y = film[:, 0] * y + film[:, 1]
        else:
            y = y + film
        y = self.blocks[1](y)
        return y + self.residual_conv(x)


class Unet1D(nn.Module):
    def __init__(
        self,
        action_dim,
        cond_dim=None,
        diffusion_step_embed_dim=32,
        dim=32,
        dim_mults=(1, 2, 4, 8),
        smaller_encoder=False,
        cond_mlp_dims=None,
        kernel_size=5,
        n_groups=None,
        activation_type="Mish",
        cond_predict_scale=False,
        groupnorm_eps=1e-5,
    ):
        super().__init__()
        widths = [action_dim] + [dim * m for m in dim_mults]
        pairs = list(zip(widths[:-1], widths[1:]))
        e = diffusion_step_embed_dim
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(e), nn.Linear(e, e * 4), nn.Mish(), nn.Linear(e * 4, e))
        if cond_mlp_dims is not None:
            self.cond_mlp = ResidualMLP(
                dim_list=[cond_dim] + list(cond_mlp_dims), activation_type=activation_type, out_activation_type="Identity"
            )
            g_dim = e + cond_mlp_dims[-1]
        else:
            g_dim = e + cond_dim
        big = cond_mlp_dims is None and not smaller_encoder

        def res(cin, cout):
            return ResidualBlock1D(
                cin, cout, cond_dim=g_dim, kernel_size=kernel_size, n_groups=n_groups,
                cond_predict_scale=cond_predict_scale, larger_encoder=big, activation_type=activation_type,
                groupnorm_eps=groupnorm_eps,
            )

        top = widths[-1]
        self.mid_modules = nn.ModuleList([res(top, top), res(top, top)])
        self.down_modules = nn.ModuleList([])
        for k, (cin, cout) in enumerate(pairs):
            last = k >= len(pairs) - 1
            self.down_modules.append(
                nn.ModuleList([res(cin, cout), res(cout, cout), Downsample1d(cout) if not last else nn.Identity()])
            )
        self.up_modules = nn.ModuleList([])
        for k, (cin, cout) in enumerate(reversed(pairs[1:])):
            last = k >= len(pairs) - 1  # never true: the reference always upsamples (unet.py:223-224)
            self.up_modules.append(
                nn.ModuleList([res(cout * 2, cin), res(cin, cin), Upsample1d(cin) if not last else nn.Identity()])
            )
        self.final_conv = nn.Sequential(
            Conv1dBlock(dim, dim, kernel_size=kernel_size, n_groups=n_groups, activation_type=activation_type,
                        eps=groupnorm_eps),
            DenseConv1d(dim, action_dim, 1),
        )
        self.time_dim = e

    def forward(self, x, time, cond, **kwargs):
        B = len(x)
        h = x.permute(0, 2, 1)
        state = cond["state"].reshape(B, -1)
        if hasattr(self, "cond_mlp"):
            state = self.cond_mlp(state)
        if not torch.is_tensor(time):
            time = torch.tensor([time], dtype=torch.long, device=x.device)
        elif time.ndim == 0:
            time = time[None].to(x.device)
        g = torch.cat([self.time_mlp(time.expand(B)), state], dim=-1)
        skips = []
        for r1, r2, down in self.down_modules:
            h = r2(r1(h, g), g)
            skips.append(h)
            h = down(h)
        for m in self.mid_modules:
            h = m(h, g)
        for r1, r2, up in self.up_modules:
            h = torch.cat((h, skips.pop()), dim=1)
            h = up(r2(r1(h, g), g))
        return self.final_conv(h).permute(0, 2, 1)
