"""
ORACLE — test infrastructure only.  CPU restatement (PyTorch fp32 eager + float64 numpy) of the reference DPPO hot
path.  Nothing under dppo_b200/ imports this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may.  The product path never routes through it.

Parity status: PINNED.  The reference (enyen/dppo) ships no tests or golden vectors, so the pin is the
reference's own code executed in the build container: tests/golden/make_golden.py imports /root/reference, runs every
function below's counterpart on seeded inputs with injected noise and stores the outputs under tests/golden/*.npz;
tests/test_oracle_golden.py checks this file against those vectors (max-abs-diff 0 on CPU for the forward paths).

All arithmetic is restated from the reference, function by function (file:line cited at each definition); the tensors
are plain dict-of-tensors ("params", same keys as the reference state_dict), no nn.Module.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------------------------
@dataclass
class NetCfg:
    """Static shape of one denoiser (DiffusionMLP or Unet1D) and of the critic."""

    kind: str = "mlp"  # "mlp" | "unet"
    obs_dim: int = 11
    cond_steps: int = 1
    action_dim: int = 3
    horizon_steps: int = 4
    time_dim: int = 16
    mlp_dims: List[int] = field(default_factory=lambda: [512, 512, 512])
    cond_mlp_dims: Optional[List[int]] = None
    activation: str = "ReLU"
    use_layernorm: bool = False
    critic_dims: List[int] = field(default_factory=lambda: [256, 256, 256])
    critic_activation: str = "Mish"
    # unet only
    unet_dim: int = 64
    unet_mults: Tuple[int, ...] = (1, 2)
    unet_kernel: int = 5
    unet_groups: int = 8
    unet_cond_predict_scale: bool = True
    unet_smaller_encoder: bool = False
    unet_eps: float = 1e-5


@dataclass
class DiffCfg:
    """Diffusion / PPO hyper-parameters (keys of PPODiffusion.__init__, reference diffusion_ppo.py:25-36)."""

    denoising_steps: int = 20
    ft_denoising_steps: int = 10
    use_ddim: bool = False
    ddim_steps: Optional[int] = None
    eta: float = 1.0  # EtaFixed value (base_eta)
    denoised_clip_value: Optional[float] = 1.0
    randn_clip_value: float = 3.0
    final_action_clip_value: Optional[float] = None
    eps_clip_value: Optional[float] = None
    min_sampling_denoising_std: float = 0.1
    min_logprob_denoising_std: float = 0.1
    gamma_denoising: float = 0.99
    clip_ploss_coef: float = 0.01
    clip_ploss_coef_base: float = 0.01
    clip_ploss_coef_rate: float = 3.0
    clip_vloss_coef: Optional[float] = None
    clip_advantage_lower_quantile: float = 0.0
    clip_advantage_upper_quantile: float = 1.0
    norm_adv: bool = True


# --------------------------------------------------------------------------------------------------------------------
# a1: schedule tables
# --------------------------------------------------------------------------------------------------------------------
def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> torch.Tensor:
    """float64 numpy cosine schedule -> fp32 tensor.  reference dppo/model/diffusion/sampling.py:10-20"""
    n = timesteps + 1
    grid = np.linspace(0, n, n)
    abar = np.cos(((grid / n) + s) / (1 + s) * np.pi * 0.5) ** 2
    abar = abar / abar[0]
    betas = 1 - (abar[1:] / abar[:-1])
    return torch.tensor(np.clip(betas, a_min=0, a_max=0.999), dtype=torch.float32)


def make_tables(cfg: DiffCfg) -> Dict[str, torch.Tensor]:
    """DDPM tables reference dppo/model/diffusion/diffusion.py:98-148; DDIM tables (flipped) :155-196."""
    tb: Dict[str, torch.Tensor] = {}
    betas = cosine_beta_schedule(cfg.denoising_steps)
    alphas = 1.0 - betas
    abar = torch.cumprod(alphas, axis=0)
    abar_prev = torch.cat([torch.ones(1), abar[:-1]])
    tb["betas"] = betas
    tb["alphas_cumprod"] = abar
    tb["sqrt_recip_alphas_cumprod"] = torch.sqrt(1.0 / abar)
    tb["sqrt_recipm1_alphas_cumprod"] = torch.sqrt(1.0 / abar - 1)
    var = betas * (1.0 - abar_prev) / (1.0 - abar)
    tb["ddpm_logvar_clipped"] = torch.log(torch.clamp(var, min=1e-20))
    tb["ddpm_mu_coef1"] = betas * torch.sqrt(abar_prev) / (1.0 - abar)
    tb["ddpm_mu_coef2"] = (1.0 - abar_prev) * torch.sqrt(alphas) / (1.0 - abar)
    if cfg.use_ddim:
        ratio = cfg.denoising_steps // cfg.ddim_steps
        ddim_t = torch.arange(0, cfg.ddim_steps) * ratio
        a = abar[ddim_t].clone().to(torch.float32)
        a_prev = torch.cat([torch.tensor([1.0]).to(torch.float32), abar[ddim_t[:-1]]])
        s1m = (1.0 - a) ** 0.5
        tb["ddim_t"] = torch.flip(ddim_t, [0])
        tb["ddim_alphas"] = torch.flip(a, [0])
        tb["ddim_alphas_prev"] = torch.flip(a_prev, [0])
        tb["ddim_sqrt_one_minus_alphas"] = torch.flip(s1m, [0])
    return tb


# --------------------------------------------------------------------------------------------------------------------
# a5 / a10: networks (functional)
# --------------------------------------------------------------------------------------------------------------------
def _act(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "ReLU":
        return F.relu(x)
    if name == "Mish":
        return F.mish(x)
    if name == "Identity":
        return x
    raise KeyError(name)


def sinusoidal(t: torch.Tensor, dim: int) -> torch.Tensor:
    """reference dppo/model/diffusion/modules.py:20-27 (t is the (B,1) or (B,) tensor the caller passes)."""
    half = dim // 2
    rate = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half) * -rate)
    ph = t[:, None] * freq[None, :]
    return torch.cat((ph.sin(), ph.cos()), dim=-1)


def residual_mlp(p: Params, pre: str, n_blocks: int, act: str, ln: bool, x: torch.Tensor) -> torch.Tensor:
    """reference dppo/model/common/mlp.py:84-154: Linear, n pre-activation blocks (LN eps 1e-6), Linear."""
    h = F.linear(x, p[f"{pre}layers.0.weight"], p[f"{pre}layers.0.bias"])
    for b in range(1, n_blocks + 1):
        q = f"{pre}layers.{b}."
        y = h
        if ln:
            y = F.layer_norm(y, (y.shape[-1],), p[q + "norm1.weight"], p[q + "norm1.bias"], 1e-6)
        y = F.linear(_act(act, y), p[q + "l1.weight"], p[q + "l1.bias"])
        if ln:
            y = F.layer_norm(y, (y.shape[-1],), p[q + "norm2.weight"], p[q + "norm2.bias"], 1e-6)
        y = F.linear(_act(act, y), p[q + "l2.weight"], p[q + "l2.bias"])
        h = y + h
    k = n_blocks + 1
    return F.linear(h, p[f"{pre}layers.{k}.weight"], p[f"{pre}layers.{k}.bias"])


def time_mlp(p: Params, pre: str, dim: int, t: torch.Tensor) -> torch.Tensor:
    """Sequential(SinusoidalPosEmb, Linear, Mish, Linear); reference mlp_diffusion.py:191-196, unet.py:145-150."""
    e = sinusoidal(t, dim)
    e = F.linear(e, p[pre + "1.weight"], p[pre + "1.bias"])
    return F.linear(F.mish(e), p[pre + "3.weight"], p[pre + "3.bias"])


def diffusion_mlp(p: Params, pre: str, nc: NetCfg, x: torch.Tensor, t: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """DiffusionMLP.forward, reference dppo/model/diffusion/mlp_diffusion.py:218-250."""
    B, Ta, Da = x.shape
    s = state.reshape(B, -1)
    if nc.cond_mlp_dims is not None:  # MLP([cond]+dims), activation between, Identity out (mlp.py:27-81)
        n = len(nc.cond_mlp_dims)
        for i in range(n):
            s = F.linear(s, p[f"{pre}cond_mlp.moduleList.{i}.linear_1.weight"], p[f"{pre}cond_mlp.moduleList.{i}.linear_1.bias"])
            if i < n - 1:
                s = _act(nc.activation, s)
    temb = time_mlp(p, pre + "time_embedding.", nc.time_dim, t.reshape(B, 1)).reshape(B, nc.time_dim)
    z = torch.cat([x.reshape(B, -1), temb, s], dim=-1)
    n_blocks = (len(nc.mlp_dims) - 1) // 2
    out = residual_mlp(p, pre + "mlp_mean.", n_blocks, nc.activation, nc.use_layernorm, z)
    return out.reshape(B, Ta, Da)


def _conv_block(p: Params, pre: str, nc: NetCfg, x: torch.Tensor) -> torch.Tensor:
    """Conv1d -> GroupNorm -> Mish; reference dppo/model/diffusion/modules.py:50-95."""
    k = p[pre + "block.0.weight"].shape[-1]
    y = F.conv1d(x, p[pre + "block.0.weight"], p[pre + "block.0.bias"], padding=k // 2)
    y = F.group_norm(y, nc.unet_groups, p[pre + "block.2.weight"], p[pre + "block.2.bias"], nc.unet_eps)
    return _act(nc.activation, y)


def _res_block(p: Params, pre: str, nc: NetCfg, x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """ResidualBlock1D with FiLM; reference dppo/model/diffusion/unet.py:27-118."""
    out = _conv_block(p, pre + "blocks.0.", nc, x)
    if pre + "cond_encoder.4.weight" in p:  # larger encoder: Lin, act, Lin, act, Lin
        e = F.linear(g, p[pre + "cond_encoder.0.weight"], p[pre + "cond_encoder.0.bias"])
        e = F.linear(_act(nc.activation, e), p[pre + "cond_encoder.2.weight"], p[pre + "cond_encoder.2.bias"])
        e = F.linear(_act(nc.activation, e), p[pre + "cond_encoder.4.weight"], p[pre + "cond_encoder.4.bias"])
    else:  # act, Lin
        e = F.linear(_act(nc.activation, g), p[pre + "cond_encoder.1.weight"], p[pre + "cond_encoder.1.bias"])
    e = e[:, :, None]
    C = out.shape[1]
    if nc.unet_cond_predict_scale:
        e = e.reshape(e.shape[0], 2, C, 1)
        out = e[:, 0] * out + e[:, 1]
    else:
        out = out + e
    out = _conv_block(p, pre + "blocks.1.", nc, out)
    if pre + "residual_conv.weight" in p:
        x = F.conv1d(x, p[pre + "residual_conv.weight"], p[pre + "residual_conv.bias"])
    return out + x


def unet1d(p: Params, pre: str, nc: NetCfg, x: torch.Tensor, t: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """Unet1D.forward, reference dppo/model/diffusion/unet.py:267-327 (no cond_mlp variant)."""
    B = len(x)
    h = x.permute(0, 2, 1)
    s = state.reshape(B, -1)
    g = torch.cat([time_mlp(p, pre + "time_mlp.", nc.time_dim, t), s], dim=-1)
    n_lvl = len(nc.unet_mults)
    skips = []
    for lvl in range(n_lvl):
        q = f"{pre}down_modules.{lvl}."
        h = _res_block(p, q + "0.", nc, h, g)
        h = _res_block(p, q + "1.", nc, h, g)
        skips.append(h)
        if lvl < n_lvl - 1:
            h = F.conv1d(h, p[q + "2.conv.weight"], p[q + "2.conv.bias"], stride=2, padding=1)
    for m in range(2):
        h = _res_block(p, f"{pre}mid_modules.{m}.", nc, h, g)
    for lvl in range(n_lvl - 1):
        q = f"{pre}up_modules.{lvl}."
        h = torch.cat((h, skips.pop()), dim=1)
        h = _res_block(p, q + "0.", nc, h, g)
        h = _res_block(p, q + "1.", nc, h, g)
        h = F.conv_transpose1d(h, p[q + "2.conv.weight"], p[q + "2.conv.bias"], stride=2, padding=1)
    h = _conv_block(p, pre + "final_conv.0.", nc, h)
    h = F.conv1d(h, p[pre + "final_conv.1.weight"], p[pre + "final_conv.1.bias"])
    return h.permute(0, 2, 1)


def denoiser(p: Params, pre: str, nc: NetCfg, x, t, state):
    return unet1d(p, pre, nc, x, t, state) if nc.kind == "unet" else diffusion_mlp(p, pre, nc, x, t, state)


def critic_obs(p: Params, nc: NetCfg, state: torch.Tensor, pre: str = "critic.") -> torch.Tensor:
    """CriticObs.forward, reference dppo/model/common/critic.py:40-54 (residual style)."""
    s = state.reshape(len(state), -1)
    n_blocks = (len(nc.critic_dims) - 1) // 2
    return residual_mlp(p, pre + "Q1.", n_blocks, nc.critic_activation, False, s)


# --------------------------------------------------------------------------------------------------------------------
# a3: posterior mean / variance
# --------------------------------------------------------------------------------------------------------------------
def _pick(tab: torch.Tensor, idx: torch.Tensor, ndim: int) -> torch.Tensor:
    """gather + reshape to (B,1,1...), reference sampling.py:23-26"""
    return tab.gather(-1, idx).reshape(len(idx), *((1,) * (ndim - 1)))


def p_mean_var(
    p: Params, nc: NetCfg, dc: DiffCfg, tb, x, t, state, index=None, use_base_policy=False, deterministic=False,
    faithful_cost: bool = True,
):
    """
    VPGDiffusion.p_mean_var, reference dppo/model/diffusion/diffusion_vpg.py:139-224.
    `faithful_cost=True` evaluates the frozen actor on every row and then overwrites the fine-tuned rows, exactly like
    the reference (:148-163); False evaluates one network per row (same values, less work).
    """
    if dc.use_ddim:
        ft_mask = index >= (dc.ddim_steps - dc.ft_denoising_steps)
    else:
        ft_mask = t < dc.ft_denoising_steps
    ft_rows = torch.where(ft_mask)[0]
    ft_pre = "actor." if use_base_policy else "actor_ft."
    if faithful_cost:
        eps = denoiser(p, "actor.", nc, x, t, state)
        if len(ft_rows) > 0:
            eps[ft_rows] = denoiser(p, ft_pre, nc, x[ft_rows], t[ft_rows], state[ft_rows])
    else:
        eps = torch.empty_like(x)
        base_rows = torch.where(~ft_mask)[0]
        if len(base_rows) > 0:
            eps[base_rows] = denoiser(p, "actor.", nc, x[base_rows], t[base_rows], state[base_rows])
        if len(ft_rows) > 0:
            eps[ft_rows] = denoiser(p, ft_pre, nc, x[ft_rows], t[ft_rows], state[ft_rows])

    nd = x.ndim
    if dc.use_ddim:
        alpha = _pick(tb["ddim_alphas"], index, nd)
        alpha_prev = _pick(tb["ddim_alphas_prev"], index, nd)
        s1m = _pick(tb["ddim_sqrt_one_minus_alphas"], index, nd)
        x0 = (x - s1m * eps) / (alpha**0.5)
    else:
        x0 = _pick(tb["sqrt_recip_alphas_cumprod"], t, nd) * x - _pick(tb["sqrt_recipm1_alphas_cumprod"], t, nd) * eps
    if dc.denoised_clip_value is not None:
        x0 = x0.clamp(-dc.denoised_clip_value, dc.denoised_clip_value)
        if dc.use_ddim:
            eps = (x - alpha ** (0.5) * x0) / s1m
    if dc.use_ddim and dc.eps_clip_value is not None:
        eps = eps.clamp(-dc.eps_clip_value, dc.eps_clip_value)

    if dc.use_ddim:
        if deterministic:
            etas = torch.zeros((x.shape[0], 1, 1))
        else:
            etas = torch.full((x.shape[0], 1), float(torch.tensor(dc.eta, dtype=torch.float32))).unsqueeze(1)
        sigma = (etas * ((1 - alpha_prev) / (1 - alpha) * (1 - alpha / alpha_prev)) ** 0.5).clamp(min=1e-10)
        dir_coef = (1.0 - alpha_prev - sigma**2).clamp(min=0).sqrt()
        mu = (alpha_prev**0.5) * x0 + dir_coef * eps
        logvar = torch.log(sigma**2)
    else:
        mu = _pick(tb["ddpm_mu_coef1"], t, nd) * x0 + _pick(tb["ddpm_mu_coef2"], t, nd) * x
        logvar = _pick(tb["ddpm_logvar_clipped"], t, nd)
        etas = torch.ones_like(mu)
    return mu, logvar, etas


# --------------------------------------------------------------------------------------------------------------------
# a2 / a4: K-step chain with injected noise
# --------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def sample_chain(
    p: Params, nc: NetCfg, dc: DiffCfg, state: torch.Tensor, noise: torch.Tensor,
    deterministic=False, use_base_policy=False, min_sampling_std: Optional[float] = None, faithful_cost=True,
):
    """
    VPGDiffusion.forward, reference dppo/model/diffusion/diffusion_vpg.py:227-315.
    noise: (S+1, B, Ta, Da); slot 0 = x_T (not clipped), slot i+1 = the draw of step i *before* the +-randn_clip clamp.
    Returns (trajectories (B,Ta,Da), chains (B, ft+1, Ta, Da)).
    """
    tb = make_tables(dc)
    B = len(state)
    sig_min = dc.min_sampling_denoising_std if min_sampling_std is None else min_sampling_std
    x = noise[0].clone()
    if dc.use_ddim:
        t_all = tb["ddim_t"]
        n_eval = dc.ddim_steps
    else:
        t_all = list(reversed(range(dc.denoising_steps)))
        n_eval = dc.denoising_steps
    chain = []
    if dc.ft_denoising_steps == n_eval:
        chain.append(x)
    for i, t in enumerate(t_all):
        t_b = torch.full((B,), int(t), dtype=torch.long)
        i_b = torch.full((B,), i, dtype=torch.long)
        mean, logvar, _ = p_mean_var(p, nc, dc, tb, x, t_b, state, i_b, use_base_policy, deterministic, faithful_cost)
        std = torch.exp(0.5 * logvar)
        if dc.use_ddim:
            std = torch.zeros_like(std) if deterministic else torch.clip(std, min=sig_min)
        else:
            if deterministic and t == 0:
                std = torch.zeros_like(std)
            elif deterministic:
                std = torch.clip(std, min=1e-3)
            else:
                std = torch.clip(std, min=sig_min)
        z = noise[i + 1].clone().clamp_(-dc.randn_clip_value, dc.randn_clip_value)
        x = mean + std * z
        if dc.final_action_clip_value is not None and i == len(t_all) - 1:
            x = torch.clamp(x, -dc.final_action_clip_value, dc.final_action_clip_value)
        if (not dc.use_ddim and t <= dc.ft_denoising_steps) or (dc.use_ddim and i >= n_eval - dc.ft_denoising_steps - 1):
            chain.append(x)
    return x, torch.stack(chain, dim=1)


# --------------------------------------------------------------------------------------------------------------------
# a7 / a8: log-probabilities of stored chains
# --------------------------------------------------------------------------------------------------------------------
def _ft_schedule(dc: DiffCfg, tb):
    if dc.use_ddim:
        t_single = tb["ddim_t"][-dc.ft_denoising_steps:]
        idx_single = torch.arange(dc.ddim_steps - dc.ft_denoising_steps, dc.ddim_steps)
    else:
        t_single = torch.arange(dc.ft_denoising_steps - 1, -1, -1)
        idx_single = None
    return t_single, idx_single


def _normal_logprob(x, mean, std):
    """torch.distributions.Normal.log_prob restated: -(x-mu)^2/(2 var) - ln sigma - ln sqrt(2 pi)."""
    var = std**2
    return -((x - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))


def get_logprobs(p: Params, nc: NetCfg, dc: DiffCfg, state, chains, use_base_policy=False, faithful_cost=True):
    """VPGDiffusion.get_logprobs, reference diffusion_vpg.py:319-396. Rows are env-major, denoise-minor."""
    tb = make_tables(dc)
    ft = dc.ft_denoising_steps
    Bc = chains.shape[0]
    cond = state.unsqueeze(1).repeat(1, ft, *(1,) * (state.ndim - 1)).flatten(0, 1)
    t_single, idx_single = _ft_schedule(dc, tb)
    t_all = t_single.repeat(Bc, 1).flatten()
    idx_all = idx_single.repeat(Bc) if idx_single is not None else None
    prev = chains[:, :-1].reshape(-1, nc.horizon_steps, nc.action_dim)
    nxt = chains[:, 1:].reshape(-1, nc.horizon_steps, nc.action_dim)
    mean, logvar, _ = p_mean_var(p, nc, dc, tb, prev, t_all, cond, idx_all, use_base_policy, False, faithful_cost)
    std = torch.clip(torch.exp(0.5 * logvar), min=dc.min_logprob_denoising_std)
    return _normal_logprob(nxt, mean, std.expand_as(mean))


def get_logprobs_subsample(p, nc, dc, state, chains_prev, chains_next, denoising_inds, use_base_policy=False, faithful_cost=True):
    """VPGDiffusion.get_logprobs_subsample, reference diffusion_vpg.py:398-461. Returns (logp, etas)."""
    tb = make_tables(dc)
    t_single, idx_single = _ft_schedule(dc, tb)
    t_all = t_single[denoising_inds]
    idx_all = idx_single[denoising_inds] if idx_single is not None else None
    mean, logvar, etas = p_mean_var(p, nc, dc, tb, chains_prev, t_all, state, idx_all, use_base_policy, False, faithful_cost)
    std = torch.clip(torch.exp(0.5 * logvar), min=dc.min_logprob_denoising_std)
    return _normal_logprob(chains_next, mean, std.expand_as(mean)), etas


# --------------------------------------------------------------------------------------------------------------------
# a9: PPO loss
# --------------------------------------------------------------------------------------------------------------------
def ppo_loss(
    p: Params, nc: NetCfg, dc: DiffCfg, state, chains_prev, chains_next, denoising_inds, returns, oldvalues,
    advantages, oldlogprobs, reward_horizon: int = 4, faithful_cost=True, python_discount_loop=True,
    use_bc_loss=False, bc_noise=None,
):
    """
    PPODiffusion.loss, reference dppo/model/diffusion/diffusion_ppo.py:57-199.  `use_bc_loss` (:105-126) samples a chain
    from the BASE policy for every row (injected draws `bc_noise` (S+1, B, Ta, Da) stand in for torch.randn) and scores it
    under the fine-tuned policy.
    Returns (pg_loss, entropy_loss, v_loss, clipfrac, approx_kl, ratio_mean, bc_loss, eta_mean); the first three are
    graph tensors when `p` holds leaf tensors with requires_grad.
    `python_discount_loop=True` builds the denoising discount with the reference's per-row Python loop (:138-143) so
    the CPU baseline pays the same host cost; False uses a table lookup (same values).
    """
    ft = dc.ft_denoising_steps
    newlp, etas = get_logprobs_subsample(p, nc, dc, state, chains_prev, chains_next, denoising_inds, False, faithful_cost)
    entropy_loss = -etas.mean()
    newlp = newlp.clamp(min=-5, max=2)[:, :reward_horizon, :].mean(dim=(-1, -2)).view(-1)
    oldlp = oldlogprobs.clamp(min=-5, max=2)[:, :reward_horizon, :].mean(dim=(-1, -2)).view(-1)
    bc_loss = 0
    if use_bc_loss:  # diffusion_ppo.py:105-126
        with torch.no_grad():
            _, bc_chains = sample_chain(p, nc, dc, state, bc_noise, deterministic=False, use_base_policy=True, faithful_cost=faithful_cost)
        bc_lp = get_logprobs(p, nc, dc, state, bc_chains, use_base_policy=False, faithful_cost=faithful_cost)
        bc_loss = -bc_lp.clamp(min=-5, max=2).mean(dim=(-1, -2)).view(-1).mean()
    adv = advantages
    if dc.norm_adv:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    lo = torch.quantile(adv, dc.clip_advantage_lower_quantile)
    hi = torch.quantile(adv, dc.clip_advantage_upper_quantile)
    adv = adv.clamp(min=lo, max=hi)
    if python_discount_loop:
        disc = torch.tensor([dc.gamma_denoising ** (ft - i - 1) for i in denoising_inds])
    else:
        disc = torch.tensor([dc.gamma_denoising ** (ft - i - 1) for i in range(ft)])[denoising_inds]
    adv = adv * disc
    logratio = newlp - oldlp
    ratio = logratio.exp()
    tt = denoising_inds.float() / (ft - 1)
    if ft > 1:
        clip = dc.clip_ploss_coef_base + (dc.clip_ploss_coef - dc.clip_ploss_coef_base) * (
            torch.exp(dc.clip_ploss_coef_rate * tt) - 1
        ) / (math.exp(dc.clip_ploss_coef_rate) - 1)
    else:
        clip = tt
    with torch.no_grad():
        approx_kl = ((ratio - 1) - logratio).mean()
        clipfrac = ((ratio - 1.0).abs() > clip).float().mean().item()
    pg = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    newv = critic_obs(p, nc, state).view(-1)
    if dc.clip_vloss_coef is not None:
        vc = oldvalues + torch.clamp(newv - oldvalues, -dc.clip_vloss_coef, dc.clip_vloss_coef)
        v_loss = 0.5 * torch.max((newv - returns) ** 2, (vc - returns) ** 2).mean()
    else:
        v_loss = 0.5 * ((newv - returns) ** 2).mean()
    return pg, entropy_loss, v_loss, clipfrac, approx_kl.item(), ratio.mean().item(), bc_loss, etas.mean().item()


# --------------------------------------------------------------------------------------------------------------------
# a11: GAE, a12: minibatch indexing, n2: running reward scaling  (float64 numpy, as the reference agent)
# --------------------------------------------------------------------------------------------------------------------
def gae(reward, terminated, values, next_value, gamma, lam, reward_scale_const=1.0):
    """
    Reverse scan of reference dppo/agent/finetune/train_ppo_diffusion_agent.py:255-279.
    reward / terminated / values: (n_steps, E) float64; next_value: (E,) critic value of the post-rollout observation.
    Returns (advantages, returns) float64.
    """
    n = reward.shape[0]
    adv = np.zeros_like(reward)
    last = 0
    for t in reversed(range(n)):
        nxt = next_value.reshape(1, -1) if t == n - 1 else values[t + 1]
        live = 1.0 - terminated[t]
        delta = reward[t] * reward_scale_const + gamma * nxt * live - values[t]
        adv[t] = last = delta + gamma * lam * live * last
    return adv, adv + values


def minibatch_indices(perm: torch.Tensor, batch: int, batch_size: int, ft: int):
    """(b, d) rows of minibatch `batch`; reference train_ppo_diffusion_agent.py:312-320 (unravel over (N, ft))."""
    sel = perm[batch * batch_size:(batch + 1) * batch_size]
    return sel // ft, sel % ft


class RunningRewardScaler:
    """
    Running return-variance reward scaling, reference dppo/util/reward_scaling.py:12-87 (float64 numpy, one scalar
    RunningMeanStd with initial count 1e-4; note the reference divides the merged M2 by (count - 1)).
    reward, first: (E, n_steps).  State (ret, mean, var, count) persists across iterations.
    """

    def __init__(self, num_envs, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        self.mean, self.var, self.count = 0.0, 1.0, 1e-4
        self.ret = np.zeros(num_envs)
        self.cliprew, self.gamma, self.epsilon = cliprew, gamma, epsilon

    def __call__(self, reward, first):
        rets = np.zeros_like(reward)
        carry = self.ret
        for t in range(reward.shape[1]):  # reward_scaling.py:76-87
            carry = rets[:, t] = reward[:, t] + (1 - first[:, t]) * self.gamma * carry
        self.ret = rets[:, -1]
        flat = rets.reshape(-1)
        b_mean, b_var, b_n = np.mean(flat, axis=0), np.var(flat, axis=0), flat.shape[0]
        delta = b_mean - self.mean  # reward_scaling.py:31-40
        tot = self.count + b_n
        self.mean = self.mean + delta * b_n / tot
        m2 = self.var * self.count + b_var * b_n + delta**2 * self.count * b_n / tot
        self.var = m2 / (tot - 1)
        self.count = tot
        return np.clip(reward / np.sqrt(self.var + self.epsilon), -self.cliprew, self.cliprew)
