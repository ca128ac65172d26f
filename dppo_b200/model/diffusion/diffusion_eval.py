"""
Evaluation-time diffusion policy: frozen base denoiser for the early denoising steps, fine-tuned denoiser for the last
`ft_denoising_steps`, sampled by the same persistent chain kernels as the rollout (libdppo_b200 `dppo_sample_chain`).

DiffusionEval -> /root/reference/dppo/model/diffusion/diffusion_eval.py:19-150.  The constructor keywords and the
checkpoint convention are the reference's: `network_path` holds {"model": state_dict} with keys `network.*`, `actor.*`,
`actor_ft.*` (what TrainAgent.save_model writes, dppo/agent/finetune/train_agent.py:125-135); a pre-training checkpoint
without `actor.*` keys is accepted with ft_denoising_steps = 0 (`network.*` weights).  `forward(cond, deterministic)`
returns Sample(trajectories, None) like DiffusionModel.forward (diffusion.py:262-314): the noise level is 0 for DDIM and
max(sigma, 1e-3) (0 at t = 0) for DDPM whatever `deterministic` says; `deterministic` only selects eta = 0 for DDIM,
which is the mode the evaluation agent uses (eval_diffusion_agent.py:58).  There is no CPU or eager fallback.
"""

import copy
import logging

import torch

from dppo_b200.engine import ChainEngine
from dppo_b200.model.diffusion.diffusion import DiffusionModel, Sample

log = logging.getLogger(__name__)


class DiffusionEval(DiffusionModel):
    def __init__(self, network_path, ft_denoising_steps, use_ddim=False, engine_precision="split3", **kwargs):
        super().__init__(use_ddim=use_ddim, network_path=None, **kwargs)  # the base class must not load the checkpoint
        self.ft_denoising_steps = ft_denoising_steps
        self.min_logprob_denoising_std = 0.1  # unused by the sampler; the kernel context wants a value
        model_sd = torch.load(network_path, map_location=self.device, weights_only=True)["model"]

        def sub(prefix):
            return {k[len(prefix):]: v for k, v in model_sd.items() if k.startswith(prefix)}

        self.actor = self.network
        base = sub("actor.")
        if base:
            self.actor.load_state_dict(base, strict=True)
            self.actor_ft = copy.deepcopy(self.network)
            self.actor_ft.load_state_dict(sub("actor_ft."), strict=True)
            log.info("Loaded base and fine-tuned policy weights from %s", network_path)
        else:
            if ft_denoising_steps != 0:
                raise ValueError("If no base policy weights are found, ft_denoising_steps must be 0")
            self.actor.load_state_dict(sub("network."), strict=True)
            log.info("Actor weights not found in %s. Using pre-trained weights!", network_path)
        self.engine_precision = engine_precision
        self._engine = None
        self._rng_offset = 0

    def engine(self):
        if self._engine is None:
            self._engine = ChainEngine(self, precision=self.engine_precision)
        self._engine.sync_weights(0, self.actor)
        if hasattr(self, "actor_ft"):
            self._engine.sync_weights(1, self.actor_ft)
        return self._engine

    @torch.no_grad()
    def forward(self, cond, deterministic=True, noise=None):
        """cond["state"]: (B, To, Do) -> Sample(trajectories (B, Ta, Da), None).  `noise` (S+1, B, Ta, Da): test hook."""
        if self.use_ddim and not deterministic:
            raise NotImplementedError("DiffusionEval with DDIM samples with eta = 0 (deterministic=True), as the evaluation agent does")
        eng = self.engine()
        state = cond["state"]
        B = state.shape[0]
        seed = offset = 0
        if noise is None:
            seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
            self._rng_offset += 1
            offset = self._rng_offset
            if not state.is_cuda:
                # host observations in -> host actions out, one library call (ChainEngine.sample_host)
                traj, _ = eng.sample_host(state, seed, offset, 0, True, not hasattr(self, "actor_ft"), 0.0, False)
                return Sample(traj, None)
        traj, _ = eng.sample(state.to(self.device), noise=noise, seed=seed, offset=offset, deterministic=True,
                             use_base_policy=not hasattr(self, "actor_ft"), min_sampling_std=0.0, return_chain=False)
        return Sample(traj.view(B, self.horizon_steps, self.action_dim), None)
