"""
Policy-gradient diffusion policy: frozen base denoiser + fine-tuned copy + critic, with the K-step sampler and the
chain log-probabilities running in libdppo_b200.so.

VPGDiffusion -> /root/reference/dppo/model/diffusion/diffusion_vpg.py:27-461.  Method names, arguments, return shapes
and `state_dict()` keys (`network.*`, `actor.*`, `actor_ft.*`, `critic.*`, `eta.*`) are the reference's; the bodies
call the C ABI (include/dppo_b200.h) instead of ~113 eager ops per denoising step:

  forward                 -> dppo_sample_chain      (one persistent tcgen05 kernel for all envs x all S steps)
  get_logprobs (no grad)  -> dppo_chain_logprobs    (same kernel, teacher-forced over the ft window)
  get_logprobs_subsample  -> actor_ft (autograd) + dppo_logprob_rows (closed-form derivative kept for backward)

There is no CPU or eager fallback for these methods.
"""

import copy
import logging
import os

import torch
from torch.nn.modules import module as _nn_module

from dppo_b200.engine import ChainEngine
from dppo_b200.model.diffusion.diffusion import DiffusionModel, Sample

log = logging.getLogger(__name__)


class _LogProbRows(torch.autograd.Function):
    """logp(eps) of dppo_logprob_rows with d logp / d eps = (x_next - mu)/sigma^2 * d mu/d eps from the same kernel."""

    @staticmethod
    def forward(ctx, eps, engine, x_prev, x_next, denoising_inds):
        shape = eps.shape
        logp, fac = engine.logprob_rows(eps.detach(), x_prev, x_next, denoising_inds, want_grad_factor=True)
        ctx.save_for_backward(fac)
        ctx.shape = shape
        return logp.reshape(shape)

    @staticmethod
    def backward(ctx, grad_out):
        (fac,) = ctx.saved_tensors
        return (grad_out.reshape(fac.shape) * fac).reshape(ctx.shape), None, None, None, None


class VPGDiffusion(DiffusionModel):
    def __init__(
        self,
        actor,
        critic,
        ft_denoising_steps,
        ft_denoising_steps_d=0,
        ft_denoising_steps_t=0,
        network_path=None,
        min_sampling_denoising_std=0.1,
        min_logprob_denoising_std=0.1,
        eta=None,
        learn_eta=False,
        engine_precision=None,
        **kwargs,
    ):
        super().__init__(network=actor, network_path=network_path, **kwargs)
        if ft_denoising_steps > self.denoising_steps or (self.use_ddim and ft_denoising_steps > self.ddim_steps):
            raise ValueError("ft_denoising_steps exceeds the number of denoising steps")
        if learn_eta and not self.use_ddim:
            raise ValueError("Cannot learn eta with DDPM.")
        self.ft_denoising_steps = ft_denoising_steps
        self.ft_denoising_steps_d = ft_denoising_steps_d
        self.ft_denoising_steps_t = ft_denoising_steps_t
        self.ft_denoising_steps_cnt = 0
        self.min_sampling_denoising_std = min_sampling_denoising_std
        self.min_logprob_denoising_std = min_logprob_denoising_std
        self.learn_eta = learn_eta
        if eta is not None:
            self.eta = eta.to(self.device)
            if not learn_eta:
                for p in self.eta.parameters():
                    p.requires_grad = False
        self.actor = self.network
        self.actor_ft = copy.deepcopy(self.actor)
        for p in self.actor.parameters():
            p.requires_grad = False
        self.critic = critic.to(self.device)
        if network_path is not None:
            ckpt = torch.load(network_path, map_location=self.device, weights_only=True)
            if "ema" not in ckpt:
                self.load_state_dict(ckpt["model"], strict=False)
        # "split3" (3 x bf16 hi/lo MMAs, fp32-level parity) or "bf16" (single pass, fast mode)
        self.engine_precision = engine_precision or os.environ.get("DPPO_B200_PRECISION", "split3")
        self._engine = None
        self._rng_offset = 0

    # ------------------------------------------------------------------ engine plumbing
    def engine(self, sync=True):
        """The kernel context (created on first use; rebuilt when the fine-tuning window is annealed).  `sync` refreshes
        the packed weight copies the chain kernels read (not needed by the loss kernels)."""
        eng = self._engine
        if eng is None or eng.ft != self.ft_denoising_steps:
            eng = self._engine = ChainEngine(self, precision=self.engine_precision)
        if sync:
            mods = self._modules  # plain dict lookups: nn.Module.__getattr__ costs ~1 us per sub-module
            actor, actor_ft = mods["actor"], mods["actor_ft"]
            if not eng.weights_current(actor, actor_ft):
                eng.sync_weights(0, actor)
                eng.sync_weights(1, actor_ft)
                eng.mark_current(actor, actor_ft)
        return eng

    # ------------------------------------------------------------------ annealing (reference :102-136)
    def step(self):
        if type(self.min_sampling_denoising_std) is not float:
            self.min_sampling_denoising_std.step()
        self.ft_denoising_steps_cnt += 1
        if (
            self.ft_denoising_steps_d > 0
            and self.ft_denoising_steps_t > 0
            and self.ft_denoising_steps_cnt % self.ft_denoising_steps_t == 0
        ):
            self.ft_denoising_steps = max(0, self.ft_denoising_steps - self.ft_denoising_steps_d)
            self.actor = self.actor_ft
            self.actor_ft = copy.deepcopy(self.actor)
            for p in self.actor.parameters():
                p.requires_grad = False
            log.info("Annealed fine-tuning denoising steps to %d", self.ft_denoising_steps)

    def get_min_sampling_denoising_std(self):
        if type(self.min_sampling_denoising_std) is float:
            return self.min_sampling_denoising_std
        return self.min_sampling_denoising_std()

    def __call__(self, *args, **kwargs):
        # A rollout decision of a handful of envs is latency-bound end to end (0.145 ms of kernel): nn.Module.__call__
        # spends ~3 us on hook bookkeeping before it reaches forward().  With no hook registered anywhere it IS forward().
        if (self._forward_hooks or self._forward_pre_hooks or self._backward_hooks or self._backward_pre_hooks
                or _nn_module._global_forward_hooks or _nn_module._global_forward_pre_hooks
                or _nn_module._global_backward_hooks or _nn_module._global_backward_pre_hooks):
            return super().__call__(*args, **kwargs)
        return self.forward(*args, **kwargs)

    # ------------------------------------------------------------------ sampling (reference :227-315)
    def forward(self, cond, deterministic=False, return_chain=True, use_base_policy=False, noise=None, env_offset=0,
                out_trajectories=None, out_chains=None):
        """
        cond["state"]: (B, To, Do).  Returns Sample(trajectories (B, Ta, Da), chains (B, ft+1, Ta, Da)).
        With cond["state"] on the HOST and no `out_trajectories`, the results come back on the host too: views into
        a ring of page-locked buffers the kernel stored into directly (valid for the next 3 calls), complete on return -
        the reference's obs.to(device) -> model(cond) -> .cpu().numpy() as one call without copy launches; an
        `out_chains` tensor (device-resident rollout buffer slice or pinned) then receives the chains instead of the ring.
        `out_trajectories` / `out_chains`: optional preallocated float32 tensors of those shapes, on the device or in
        PINNED host memory - the kernel then stores straight into them (over PCIe for pinned memory, overlapped with the
        chain instead of a copy after it) and they are what Sample holds; cond["state"] may be pinned host memory too.
        `noise` (S+1, B, Ta, Da) injects the draws the reference takes from torch.randn / randn_like (parity tests);
        without it the kernel draws Philox normals seeded from torch's generator.  `env_offset` = global index of
        row 0 (env-sharded ranks then draw what one process would draw for the same envs).
        """
        state = cond["state"]
        if state.is_cuda or noise is not None or out_trajectories is not None:
            return self._forward_device(state, deterministic, return_chain, use_base_policy, noise, env_offset,
                                        out_trajectories, out_chains)
        # host observations in -> host results out: one library call per decision (ChainEngine.sample_host).  This is the
        # latency path of small env counts, so it touches no torch op (nothing autograd could record) and keeps the
        # per-call Python work to the weight-cache check and one ctypes call.
        eng = self.engine()
        offset = self._rng_offset + 1
        self.__dict__["_rng_offset"] = offset  # not through nn.Module.__setattr__ (2 us of isinstance checks)
        min_std = self.min_sampling_denoising_std
        if type(min_std) is not float:
            min_std = float(min_std())
        return Sample(*eng.sample_host(state, torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, offset, env_offset, deterministic,
                                       use_base_policy, min_std, return_chain, out_chains))

    @torch.no_grad()
    def _forward_device(self, state, deterministic, return_chain, use_base_policy, noise, env_offset, out_trajectories,
                        out_chains):
        eng = self.engine()
        B = state.shape[0]
        seed = offset = 0
        if noise is None:
            seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
            self._rng_offset += 1
            offset = self._rng_offset
        if not (state.is_cuda or state.is_pinned()):
            state = state.to(self.device)
        traj, chain = eng.sample(
            state, noise=noise, seed=seed, offset=offset, env_offset=env_offset, deterministic=deterministic,
            use_base_policy=use_base_policy, min_sampling_std=float(self.get_min_sampling_denoising_std()),
            return_chain=return_chain, out_traj=out_trajectories, out_chain=out_chains if return_chain else None,
        )
        traj = traj.view(B, self.horizon_steps, self.action_dim)
        if chain is not None:
            chain = chain.view(B, self.ft_denoising_steps + 1, self.horizon_steps, self.action_dim)
        return Sample(traj, chain)

    # ------------------------------------------------------------------ log-probabilities (reference :319-461)
    def _ft_timesteps(self, denoising_inds):
        if self.use_ddim:
            t_single = self.ddim_t[-self.ft_denoising_steps:]
        else:
            t_single = torch.arange(start=self.ft_denoising_steps - 1, end=-1, step=-1, device=self.device)
        return t_single[denoising_inds]

    def get_logprobs(self, cond, chains, get_ent=False, use_base_policy=False):
        """chains (B, ft+1, Ta, Da) -> log-probs (B*ft, Ta, Da), rows env-major / denoise-minor."""
        if get_ent:
            raise NotImplementedError("entropy of the chain is constant for fixed eta; the reference never requests it here")
        B, ft = chains.shape[0], self.ft_denoising_steps
        actor = self.actor if use_base_policy else self.actor_ft
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in actor.parameters())
        if not needs_grad:
            logp = self.engine().chain_logprobs(cond["state"], chains, use_base_policy=use_base_policy)
            return logp.view(B * ft, self.horizon_steps, self.action_dim)
        # differentiable variant (BC regulariser inside the loss, reference diffusion_ppo.py:105-126)
        state = cond["state"].unsqueeze(1).repeat(1, ft, *(1,) * (cond["state"].ndim - 1)).flatten(0, 1)
        dinds = torch.arange(ft, device=chains.device).repeat(B)
        prev = chains[:, :-1].reshape(B * ft, self.horizon_steps, self.action_dim)
        nxt = chains[:, 1:].reshape(B * ft, self.horizon_steps, self.action_dim)
        return self.get_logprobs_subsample({"state": state}, prev, nxt, dinds, use_base_policy=use_base_policy)

    def get_logprobs_subsample(self, cond, chains_prev, chains_next, denoising_inds, get_ent=False, use_base_policy=False):
        """One (prev, next) pair per row with its denoising index -> (B, Ta, Da) [, eta (B, 1, 1)]."""
        eng = self.engine()
        actor = self.actor if use_base_policy else self.actor_ft
        eps = actor(chains_prev, self._ft_timesteps(denoising_inds), cond=cond)
        logp = _LogProbRows.apply(eps, eng, chains_prev, chains_next, denoising_inds)
        if get_ent:
            if self.use_ddim:
                etas = self.eta(cond).unsqueeze(1)
            else:
                etas = torch.ones_like(logp)
            return logp, etas
        return logp
