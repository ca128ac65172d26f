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
import os
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

    bc_noise = None  # test hook: (S+1, B, Ta, Da) draws injected into the BC branch's sampling call (parity tests)

    def _bc_loss(self, obs):
        samples = self.forward(cond=obs, deterministic=False, return_chain=True, use_base_policy=True, noise=self.bc_noise)
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

    # ------------------------------------------------------------------ tensor-core update path
    def fused_update_reason(self):
        """None when actor_ft / critic forward + backward run on the hand-written tcgen05 kernels (dppo_update_*), else
        the reason the torch-autograd path is used (Unet1D actors, non-residual critics, DPPO_B200_UPDATE=autograd)."""
        if os.environ.get("DPPO_B200_UPDATE", "fused") == "autograd":
            return "DPPO_B200_UPDATE=autograd"
        from dppo_b200.update_engine import unsupported_reason

        return unsupported_reason(self)

    def update_plan(self, rows):
        """The dppo_update workspace for minibatches of up to `rows` rows (rebuilt when it has to grow or the kernel
        context changed, e.g. after the fine-tuning window was annealed)."""
        from dppo_b200.update_engine import UpdatePlan

        eng = self.engine(sync=False)
        plan = getattr(self, "_update_plan", None)
        if plan is None or plan.engine is not eng or plan.max_rows < rows:
            self._update_plan = plan = UpdatePlan(self, rows)
        return plan

    @torch.no_grad()
    def values(self, cond):
        """critic(cond) -> (rows,) for the value pass over the rollout buffer and the bootstrap value (reference
        train_ppo_diffusion_agent.py:197-206, 259-263).  On the tensor-core update path this runs the critic through the
        same hand-written row GEMMs as the minibatch forward (dppo_update_values), in chunks of the plan's workspace;
        otherwise it is the critic module's own forward."""
        state = cond["state"]
        rows = state.shape[0]
        if rows == 0 or not state.is_cuda or self.fused_update_reason() is not None:
            return self.critic(cond).view(-1)
        from dppo_b200.update_engine import _c32

        flat = _c32(state, (rows, -1))
        plan = getattr(self, "_update_plan", None)
        if plan is None or plan.engine is not self.engine(sync=False):
            plan = self.update_plan(min(rows, 16384))
        plan.bind_model(self)
        out = torch.empty(rows, dtype=torch.float32, device=flat.device)
        for s in range(0, rows, plan.max_rows):
            e = min(rows, s + plan.max_rows)
            plan.values(flat[s:e], out[s:e])
        return out

    def _loss_fused(self, obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
                    reward_horizon):
        from dppo_b200.engine import _mlp_param_list
        from dppo_b200.update_engine import FusedLoss, _c32, critic_param_list

        eng = self.engine(sync=False)
        B = denoising_inds.shape[0]
        plan = self.update_plan(B)
        lo, hi = self._adv_bounds(advantages)
        hp = eng.make_hp(self, reward_horizon, lo, hi)
        tensors = dict(obs=_c32(obs["state"], (B, -1)), chains=_c32(chains_prev, (B, -1)), x_next=_c32(chains_next, (B, -1)),
                       old_logprobs=_c32(oldlogprobs, (B, -1)), returns=_c32(returns, (B,)), old_values=_c32(oldvalues, (B,)),
                       advantages=_c32(advantages, (B,)), denoising_inds=denoising_inds.detach().contiguous().to(torch.int64))
        ap, cp = _mlp_param_list(self.actor_ft), critic_param_list(self.critic)
        pg_loss, v_loss, scalars = FusedLoss.apply(self, plan, hp, tensors, len(ap), *ap, *cp)
        eta_mean = self._eta_value()
        entropy_loss = torch.full((), -eta_mean, device=pg_loss.device)
        host = scalars.tolist()  # the one device->host read of the call (the reference does four .item()s)
        return pg_loss, entropy_loss, v_loss, host[3], host[2], host[4], 0, eta_mean

    def update_minibatch(self, obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k, inds_all, row_begin=0,
                         row_count=None, reward_horizon=4, vf_coef=0.5, with_actor=True, scalars_out=None, actor_event=None):
        """
        One PPO minibatch (slice) entirely inside libdppo_b200: gather -> actor_ft / critic forward -> fused loss ->
        backward of pg_loss + vf_coef * v_loss, gradients ACCUMULATED into the parameters' .grad tensors (the views of
        the agent's flat all-reduce buffer; the caller zeroes them).  Arguments as loss_gathered.  `scalars_out`
        (8 floats, device) receives the partial means [pg, v, kl, clipfrac, ratio, -, adv mean, adv std]; nothing is
        read back to the host.  `actor_event`: a recorded torch.cuda.Event the library records again behind the actor
        backward (multi-GPU callers start the actor segment's all-reduce there, FlatGradBuffer.allreduce_split).
        reference: train_ppo_diffusion_agent.py:316-364.
        """
        eng = self.engine(sync=False)
        row_count = inds_all.numel() - row_begin if row_count is None else row_count
        plan = self.update_plan(row_count)
        plan.bind_model(self)
        lo, hi = self._adv_bounds(advantages_k[inds_all // self.ft_denoising_steps]) if (
            self.clip_advantage_lower_quantile > 0 or self.clip_advantage_upper_quantile < 1) else (-math.inf, math.inf)
        hp = eng.make_hp(self, reward_horizon, lo, hi)
        from dppo_b200.update_engine import UpdatePlan

        for t in (obs_k, chains_k, logprobs_k, returns_k, values_k, advantages_k):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("rollout buffers must be contiguous fp32")
        batch = UpdatePlan._batch(row_count, inds_all.numel(), row_begin, obs=obs_k, chains=chains_k, old_logprobs=logprobs_k,
                                  returns=returns_k, old_values=values_k, advantages=advantages_k, inds_all=inds_all)
        scalars = scalars_out if scalars_out is not None else torch.empty(8, dtype=torch.float32, device=obs_k.device)
        plan.set_actor_event(actor_event if with_actor else None)
        plan.minibatch(batch, hp, vf_coef, with_actor, scalars, eng._ws)
        return scalars

    # ------------------------------------------------------------------ reference signature
    def loss(self, obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages, oldlogprobs,
             use_bc_loss=False, reward_horizon=4):
        """Returns (pg_loss, entropy_loss, v_loss, clipfrac, approx_kl, ratio, bc_loss, eta)."""
        if not use_bc_loss and self.fused_update_reason() is None:
            return self._loss_fused(obs, chains_prev, chains_next, denoising_inds, returns, oldvalues, advantages,
                                    oldlogprobs, reward_horizon)
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
