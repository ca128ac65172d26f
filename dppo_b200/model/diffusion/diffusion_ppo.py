"""
DPPO: PPO on the two-level (environment x denoising) MDP.

PPODiffusion -> /root/reference/dppo/model/diffusion/diffusion_ppo.py:24-199.  `loss` keeps the reference signature
and 8-tuple return; its body is: actor_ft / critic forward (autograd), then ONE fused kernel (dppo_ppo_loss_rows) that
evaluates the Gaussian log-probabilities, ratio, per-step clip, policy / value losses and diagnostics and emits the
closed-form gradients w.r.t. the two network outputs, which autograd then carries into the parameters.
`loss_gathered` is the same computation with the minibatch gathers fused into the kernel (dppo_ppo_loss_fwd_bwd); the
B200 agent uses it so the rollout buffers are indexed in place.
"""

import math
from typing import Optional

import torch

from dppo_b200.model.diffusion.diffusion_vpg import VPGDiffusion


class _FusedLoss(torch.autograd.Function):
    """(eps, vpred) -> (pg_loss, v_loss); gradients were produced by the forward kernel."""

    @staticmethod
    def forward(ctx, eps, vpred, grad_eps, grad_v, scalars):
        ctx.save_for_backward(grad_eps, grad_v)
        ctx.shapes = (eps.shape, vpred.shape)
        return scalars[0].clone(), scalars[1].clone()

    @staticmethod
    def backward(ctx, g_pg, g_v):
        grad_eps, grad_v = ctx.saved_tensors
        return (g_pg * grad_eps).reshape(ctx.shapes[0]), (g_v * grad_v).reshape(ctx.shapes[1]), None, None, None


class PPODiffusion(VPGDiffusion):
    def __init__(
        self,
        gamma_denoising: float,
        clip_ploss_coef: float,
        clip_ploss_coef_base: float = 1e-3,
        clip_ploss_coef_rate: float = 3,
        clip_vloss_coef: Optional[float] = None,
        clip_advantage_lower_quantile: float = 0,
        clip_advantage_upper_quantile: float = 1,
        norm_adv: bool = True,
        **kwargs,
    ):
        super().__init__(**kwargs)
        self.norm_adv = norm_adv
        self.clip_ploss_coef = clip_ploss_coef
        self.clip_ploss_coef_base = clip_ploss_coef_base
        self.clip_ploss_coef_rate = clip_ploss_coef_rate
        self.clip_vloss_coef = clip_vloss_coef
        self.gamma_denoising = gamma_denoising
        self.clip_advantage_lower_quantile = clip_advantage_lower_quantile
        self.clip_advantage_upper_quantile = clip_advantage_upper_quantile

    # ------------------------------------------------------------------ helpers
    def _adv_bounds(self, advantages):
        """Clamp bounds in normalised-advantage space; quantiles 0 / 1 (every YAML) are min / max, i.e. a no-op."""
        lo_q, hi_q = self.clip_advantage_lower_quantile, self.clip_advantage_upper_quantile
        if lo_q <= 0 and hi_q >= 1:
            return -math.inf, math.inf
        adv = advantages.float()
        if self.norm_adv:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        return float(torch.quantile(adv, lo_q)), float(torch.quantile(adv, hi_q))

    def _bc_loss(self, obs):
        samples = self.forward(cond=obs, deterministic=False, return_chain=True, use_base_policy=True)
        bc = self.get_logprobs(obs, samples.chains, get_ent=False, use_base_policy=False)
        return -bc.clamp(min=-5, max=2).mean(dim=(-1, -2)).view(-1).mean()

    def _eta_value(self):
        """EtaFixed's value, read from the device once (the reference calls .item() per loss call, eta.py:40)."""
        if not (self.use_ddim and hasattr(self, "eta")):
            return 1.0
        key = tuple(p._version for p in self.eta.parameters())
        if getattr(self, "_eta_cache", (None, None))[0] != key:
            self._eta_cache = (key, float(self.eta.value()))
        return self._eta_cache[1]

    def _finish(self, eps, vpred, grad_eps, grad_v, scalars, obs, use_bc_loss, scalars_out=None):
        pg_loss, v_loss = _FusedLoss.apply(eps, vpred, grad_eps, grad_v, scalars)
        bc_loss = self._bc_loss(obs) if use_bc_loss else 0
        eta_mean = self._eta_value()
        entropy_loss = torch.full((), -eta_mean, device=eps.device)
        if scalars_out is not None:
            # multi-rank caller: the partial means stay on the device and ride the gradient all-reduce
            scalars_out.copy_(scalars)
            return pg_loss, entropy_loss, v_loss, scalars[3], scalars[2], scalars[4], bc_loss, eta_mean
        host = scalars.tolist()  # the one device->host read of the call (the reference does four .item()s)
        return pg_loss, entropy_loss, v_loss, host[3], host[2], host[4], bc_loss, eta_mean

    # ------------------------------------------------------------------ reference signature
    def loss(self, obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
             use_bc_loss=False, reward_horizon=4):
        """Returns (pg_loss, entropy_loss, v_loss, clipfrac, approx_kl, ratio, bc_loss, eta)."""
        eng = self.engine()
        eps = self.actor_ft(chains_prev, self._ft_timesteps(denoising_inds), cond=obs)
        vpred = self.critic(obs).view(-1)
        lo, hi = self._adv_bounds(advantages)
        hp = eng.make_hp(self, reward_horizon, lo, hi)
        grad_eps, grad_v, scalars = eng.loss_rows(hp, chains_prev, chains_next, oldlogprobs, returns, oldvalues,
                                                  advantages, denoising_inds, eps.detach(), vpred.detach())
        return self._finish(eps, vpred, grad_eps, grad_v, scalars, obs, use_bc_loss)

    # ------------------------------------------------------------------ fused-gather variant (B200 agent)
    def loss_gathered(self, obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k, inds_all, row_begin=0,
                      row_count=None, use_bc_loss=False, reward_horizon=4, scalars_out=None):
        """
        Rollout buffers stay in place: obs_k (N, To, Do), chains_k (N, ft+1, Ta, Da), logprobs_k (N, ft, Ta, Da),
        returns_k / values_k / advantages_k (N,), inds_all = the minibatch's flat indices into (N, ft)
        (reference train_ppo_diffusion_agent.py:316-327).  This rank evaluates rows [row_begin, row_begin+row_count)
        and divides by len(inds_all), so summing ranks' losses / gradients gives the single-process result.
        `scalars_out` (8 floats, device): the kernel's partial means are copied there and NOT read back to the host
        (the caller all-reduces them together with the gradients); elements 3-5 of the return are then device scalars.
        """
        eng = self.engine(sync=False)  # the update reads the nn.Parameters, not the packed rollout copies
        ft = self.ft_denoising_steps
        row_count = inds_all.numel() - row_begin if row_count is None else row_count
        mine = inds_all[row_begin:row_begin + row_count]
        b, d = mine // ft, mine % ft
        obs = {"state": obs_k[b]}
        eps = self.actor_ft(chains_k[b, d], self._ft_timesteps(d), cond=obs)
        vpred = self.critic(obs).view(-1)
        lo, hi = self._adv_bounds(advantages_k[inds_all // ft])
        hp = eng.make_hp(self, reward_horizon, lo, hi)
        grad_eps, grad_v, scalars = eng.loss_gathered(hp, chains_k, logprobs_k, returns_k, values_k, advantages_k,
                                                      inds_all, row_begin, eps.detach(), vpred.detach())
        return self._finish(eps, vpred, grad_eps, grad_v, scalars, obs, use_bc_loss, scalars_out)
