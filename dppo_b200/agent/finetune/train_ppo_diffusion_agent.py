"""
DPPO fine-tuning loop, B200-native.

TrainPPODiffusionAgent -> /root/reference/dppo/agent/finetune/train_ppo_diffusion_agent.py:21-483 (+ the constructor
chain train_ppo_agent.py:16-89, train_agent.py:21-120).  Same config keys, same iteration structure
(eval / train iterations, rollout of n_steps decisions, old log-probs + values, reward scaling, GAE, update_epochs x
minibatches with the KL early stop, LR schedules, model.step(), checkpoints with the reference's state_dict keys).
What changed is where the data lives and what executes the hot path:

  rollout     model(cond) = ONE persistent sm_100a kernel per decision (dppo_sample_chain); chains are written into a
              device-resident (n_steps, E, ft+1, Ta, Da) fp32 buffer - only the action chunk goes back to the host for
              the simulator (the reference copies actions AND chains to float64 numpy every step, :117-145).
  prologue    get_logprobs (dppo_chain_logprobs) and the critic run on the device buffers in chunks of
              logprob_batch_size; GAE is dppo_gae_f64 (float64, one thread per env) instead of a host numpy loop.
  update      the minibatch gathers are fused into dppo_ppo_loss_fwd_bwd (indices from the unchanged torch.randperm);
              one device->host read per minibatch (the reference does 4 .item()s + Bmb element reads).
  multi-GPU   envs are sharded over ranks, rollout buffers all-gathered once per iteration, every minibatch is split
              into contiguous per-rank slices and gradients + diagnostics are summed by ONE all-reduce of a flat buffer
              (dppo_b200/distributed.py); W ranks reproduce the single-process update.

There is no CPU fallback: the model's methods raise without a CUDA device.
"""

import logging
import math
import os
import pickle
import random
import time

import numpy as np
import torch

from dppo_b200 import distributed as D
from dppo_b200 import engine as E_
from dppo_b200.agent.finetune.graphed import MinibatchStep
from dppo_b200.optim import FlatAdamW
from dppo_b200.util.config import instantiate
from dppo_b200.util.reward_scaling import RunningRewardScaler, RunningRewardScalerCUDA  # noqa: F401

log = logging.getLogger(__name__)


def cosine_warmup_lr(step, first_cycle_steps, max_lr, min_lr, warmup_steps):
    """Learning rate after `step` >= 1 calls of CosineAnnealingWarmupRestarts.step() (reference dppo/util/scheduler.py:
    _LRScheduler's constructor already takes one step, so the scheduler sits at step_in_cycle = 0 when the agent starts
    and at step_in_cycle = n after n calls; cycle_mult = gamma = 1): linear warm-up from min_lr, then cosine to min_lr,
    restarting every first_cycle_steps.  Before the first call the optimiser sits at min_lr (init_lr()), step = 0."""
    if step <= 0:
        return min_lr
    s = step % first_cycle_steps
    if s < warmup_steps:
        return (max_lr - min_lr) * s / warmup_steps + min_lr
    return min_lr + (max_lr - min_lr) * (1 + math.cos(math.pi * (s - warmup_steps) / (first_cycle_steps - warmup_steps))) / 2


class TrainPPODiffusionAgent:
    def __init__(self, cfg, venv=None):
        self.cfg = cfg
        self.device = cfg.device
        self.seed = cfg.get("seed", 42)
        random.seed(self.seed)
        np.random.seed(self.seed)
        torch.manual_seed(self.seed)
        self.rank, self.world = D.world()

        # env shard of this rank (reference: one process owns all envs, train_agent.py:43-66)
        self.n_envs_global = cfg.env.n_envs
        self.env_begin, self.env_end = D.env_shard(self.n_envs_global, self.rank, self.world)
        self.n_envs = self.env_end - self.env_begin
        self.n_cond_step, self.obs_dim, self.action_dim = cfg.cond_steps, cfg.obs_dim, cfg.action_dim
        self.act_steps, self.horizon_steps = cfg.act_steps, cfg.horizon_steps
        self.max_episode_steps = cfg.env.max_episode_steps
        self.reset_at_iteration = cfg.env.get("reset_at_iteration", True)
        if venv is None:
            from dppo_b200.env.synthetic import SyntheticVecEnv

            venv = SyntheticVecEnv(self.n_envs, self.obs_dim, self.action_dim, self.n_cond_step, self.act_steps,
                                   self.max_episode_steps, seed=self.seed, env_offset=self.env_begin)
        self.venv = venv
        self.best_reward_threshold_for_success = cfg.env.get("best_reward_threshold_for_success", 0)

        self.batch_size = cfg.train.batch_size
        self.model = instantiate(cfg.model)
        self.itr = 0
        self.n_train_itr, self.val_freq = cfg.train.n_train_itr, cfg.train.val_freq
        self.force_train = cfg.train.get("force_train", False)
        self.n_steps = cfg.train.n_steps
        self.max_grad_norm = cfg.train.get("max_grad_norm", None)
        self.logdir = cfg.logdir
        self.checkpoint_dir = os.path.join(self.logdir, "checkpoint")
        self.result_path = os.path.join(self.logdir, "result.pkl")
        if self.rank == 0:
            os.makedirs(self.checkpoint_dir, exist_ok=True)
        self.log_freq = cfg.train.get("log_freq", 1)
        self.save_model_freq = cfg.train.save_model_freq

        # PPO hyper-parameters (train_ppo_agent.py:18-89)
        self.logprob_batch_size = cfg.train.get("logprob_batch_size", 10000)
        self.gamma = cfg.train.gamma
        self.n_critic_warmup_itr = cfg.train.n_critic_warmup_itr
        # torch.optim.AdamW's update rule as one fused kernel per network over flat parameter / gradient segments
        self.actor_optimizer = FlatAdamW(self.model.actor_ft.parameters(), lr=cfg.train.actor_lr,
                                         weight_decay=cfg.train.actor_weight_decay)
        self.critic_optimizer = FlatAdamW(self.model.critic.parameters(), lr=cfg.train.critic_lr,
                                          weight_decay=cfg.train.critic_weight_decay)
        self._sched = {"actor": 0, "critic": 0}  # scheduler.step() calls so far (the reference's step_in_cycle)
        self._apply_lr()
        self.gae_lambda = cfg.train.get("gae_lambda", 0.95)
        self.target_kl = cfg.train.target_kl
        self.update_epochs = cfg.train.update_epochs
        self.ent_coef = cfg.train.get("ent_coef", 0)
        self.vf_coef = cfg.train.get("vf_coef", 0)
        self.reward_scale_running = cfg.train.reward_scale_running
        if self.reward_scale_running:
            # device-resident statistics feeding the GAE kernel directly (host mirror: RunningRewardScaler)
            self.running_reward_scaler = RunningRewardScalerCUDA(self.n_envs, self.device)
        self.reward_scale_const = cfg.train.get("reward_scale_const", 1)
        self.use_bc_loss = cfg.train.get("use_bc_loss", False)
        self.bc_loss_coeff = cfg.train.get("bc_loss_coeff", 0)
        self.reward_horizon = cfg.get("reward_horizon", self.act_steps)
        self.cuda_graph_update = cfg.train.get("cuda_graph_update", True)
        self.host_rollout = cfg.train.get("host_rollout", True)  # one host-buffer library call per decision (rollout())
        self._pinned_obs = None
        if self.model.learn_eta:
            raise NotImplementedError("learned eta is outside the hot path (no YAML enables it)")

        # gradients of both networks + 8 diagnostics in ONE flat buffer -> one all-reduce per minibatch
        self.grads = D.FlatGradBuffer([list(self.model.actor_ft.parameters()), list(self.model.critic.parameters())])
        # parity-test hooks (tests/test_gpu_agent.py, agent-level golden): "noise" -> callable(E) returning the (S+1, E, Ta, Da)
        # draws of one decision (instead of in-kernel Philox), "perm" -> callable(n) returning the minibatch permutation
        # (instead of torch.randperm on the device: a recorded CPU permutation stream cannot be reproduced by the CUDA
        # generator).  Empty in production.
        self.test_hooks = {}
        self.last_history = None  # per-minibatch diagnostics [pg, v, kl, clipfrac, ratio, ...] of the last update()
        self._actor_ft_id = id(self.model.actor_ft)
        self._actor_frozen = False
        self.timings = {}

    # ------------------------------------------------------------------ schedules
    def _apply_lr(self):
        t = self.cfg.train
        for name, opt, lr, sc in (("actor", self.actor_optimizer, t.actor_lr, t.actor_lr_scheduler),
                                  ("critic", self.critic_optimizer, t.critic_lr, t.critic_lr_scheduler)):
            v = cosine_warmup_lr(self._sched[name], sc.first_cycle_steps, lr, sc.min_lr, sc.warmup_steps)
            for g in opt.param_groups:
                g["lr"] = v

    # ------------------------------------------------------------------ checkpoints (train_agent.py:125-145)
    def save_model(self):
        if self.rank != 0:
            return
        path = os.path.join(self.checkpoint_dir, f"state_{self.itr}.pt")
        torch.save({"itr": self.itr, "model": self.model.state_dict()}, path)
        log.info("Saved model to %s", path)

    def load(self, itr):
        data = torch.load(os.path.join(self.checkpoint_dir, f"state_{itr}.pt"), weights_only=True)
        self.itr = data["itr"]
        self.model.load_state_dict(data["model"])

    def reset_env_all(self, options_venv=None):
        obs = self.venv.reset_arg(options_list=options_venv or [{} for _ in range(self.n_envs)])
        if isinstance(obs, list):
            obs = {k: np.stack([o[k] for o in obs]) for k in obs[0]}
        return obs

    # ------------------------------------------------------------------ the pieces of one iteration
    def rollout(self, prev_obs_venv, eval_mode, firsts_trajs):
        """n_steps decisions for this rank's envs.  Returns device buffers + host reward / terminated arrays."""
        dev, n, E = self.device, self.n_steps, self.n_envs
        ft = self.model.ft_denoising_steps
        obs_buf = torch.empty((n, E, self.n_cond_step, self.obs_dim), dtype=torch.float32, device=dev)
        chains_buf = torch.empty((n, E, ft + 1, self.horizon_steps, self.action_dim), dtype=torch.float32, device=dev)
        reward_trajs, terminated_trajs = np.zeros((n, E)), np.zeros((n, E))
        pinned_act = torch.empty((E, self.horizon_steps, self.action_dim), dtype=torch.float32).pin_memory()
        # Host-buffer decisions (default): the observations of the whole rollout live in ONE page-locked buffer the
        # simulator's arrays are copied into; every decision is one library call that reads its slice from there, stores
        # the chains into the device-resident rollout buffer and the action chunk into page-locked memory and returns when
        # they are there (dppo_sample_chain_host) - no copy launch, no torch op per step; the observations follow in one
        # H2D copy after the loop.  With injected noise (parity hook) or `host_rollout: False` every step runs the device
        # call with explicit copies instead.
        host_rollout = self.host_rollout and "noise" not in self.test_hooks
        shape = (n if host_rollout else 1, E, self.n_cond_step, self.obs_dim)
        if self._pinned_obs is None or tuple(self._pinned_obs.shape) != shape:  # page-locking tens of MB costs milliseconds
            self._pinned_obs = torch.empty(shape, dtype=torch.float32).pin_memory()
        pinned_obs_all = self._pinned_obs
        obs_np = pinned_obs_all.numpy()
        env_steps = 0
        for step in range(n):
            if host_rollout:
                np.copyto(obs_np[step], prev_obs_venv["state"], casting="same_kind")
                out = self.model(cond={"state": pinned_obs_all[step]}, deterministic=eval_mode, return_chain=True,
                                 env_offset=self.env_begin, out_chains=chains_buf[step])
                action_venv = out.trajectories.numpy()[:, : self.act_steps]
            else:
                np.copyto(obs_np[0], prev_obs_venv["state"], casting="same_kind")
                obs_buf[step].copy_(pinned_obs_all[0], non_blocking=True)
                # the kernel stores the chains straight into the device-resident rollout buffer and the action chunk
                # straight into pinned host memory: no copy launches after it
                noise = self.test_hooks["noise"](E).to(dev) if "noise" in self.test_hooks else None
                self.model(cond={"state": obs_buf[step]}, deterministic=eval_mode, return_chain=True, env_offset=self.env_begin,
                           out_trajectories=pinned_act, out_chains=chains_buf[step], noise=noise)
                torch.cuda.current_stream().synchronize()  # the simulator needs the action chunk on the host
                action_venv = pinned_act.numpy()[:, : self.act_steps]
            obs_venv, reward_venv, terminated_venv, truncated_venv, _ = self.venv.step(action_venv)
            done_venv = terminated_venv | truncated_venv
            reward_trajs[step], terminated_trajs[step] = reward_venv, terminated_venv
            firsts_trajs[step + 1] = done_venv
            prev_obs_venv = obs_venv
            env_steps += E * self.act_steps if not eval_mode else 0
        if host_rollout:
            obs_buf.copy_(pinned_obs_all, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the pinned buffer is refilled by the next rollout
        return obs_buf, chains_buf, reward_trajs, terminated_trajs, prev_obs_venv, done_venv, env_steps

    @torch.no_grad()
    def prologue(self, obs_buf, chains_buf, reward_trajs, terminated_trajs, firsts_trajs, last_obs):
        """Old log-probs, values, reward scaling and GAE (reference :197-279) on this rank's envs."""
        n, E = self.n_steps, self.n_envs
        ft = self.model.ft_denoising_steps
        obs_k = obs_buf.view(n * E, self.n_cond_step, self.obs_dim)
        chains_k = chains_buf.view(n * E, ft + 1, self.horizon_steps, self.action_dim)
        values = torch.empty(n * E, dtype=torch.float32, device=self.device)
        logprobs = torch.empty((n * E, ft, self.horizon_steps, self.action_dim), dtype=torch.float32, device=self.device)
        for s in range(0, n * E, self.logprob_batch_size):
            e = min(n * E, s + self.logprob_batch_size)
            values[s:e] = self.model.values({"state": obs_k[s:e]})
            logprobs[s:e] = self.model.get_logprobs({"state": obs_k[s:e]}, chains_k[s:e]).view(e - s, ft, self.horizon_steps,
                                                                                             self.action_dim)
        f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)  # noqa: E731
        reward_dev = f64(reward_trajs)
        if self.reward_scale_running:
            reward_dev = self.running_reward_scaler(reward=reward_dev, first=f64(firsts_trajs[:-1]))
        next_value = self.model.values({"state": torch.from_numpy(np.ascontiguousarray(last_obs["state"], dtype=np.float32))
                                        .to(self.device)})
        adv, ret = E_.gae(reward_dev, f64(terminated_trajs), values.view(n, E).double(), next_value.double(),
                          self.gamma, self.gae_lambda, self.reward_scale_const)
        return values.view(n, E), logprobs.view(n, E, ft, self.horizon_steps, self.action_dim), adv.float(), ret.float()

    def update(self, obs_buf, chains_buf, logprobs, values, adv, ret):
        """update_epochs x minibatches over the GLOBAL buffers (reference :305-383)."""
        Eg, n = self.n_envs_global, self.n_steps
        ft = self.model.ft_denoising_steps
        # every rank gets the reference's (step, env)-ordered *_k arrays
        obs_k = D.gather_env_dim(obs_buf, Eg).view(n * Eg, self.n_cond_step, self.obs_dim)
        chains_k = D.gather_env_dim(chains_buf, Eg).view(n * Eg, ft + 1, self.horizon_steps, self.action_dim)
        logprobs_k = D.gather_env_dim(logprobs, Eg).view(n * Eg, ft, self.horizon_steps, self.action_dim)
        values_k = D.gather_env_dim(values, Eg).reshape(-1)
        adv_k = D.gather_env_dim(adv, Eg).reshape(-1)
        ret_k = D.gather_env_dim(ret, Eg).reshape(-1)
        total_steps = n * Eg * ft
        num_batch = max(1, total_steps // self.batch_size)  # the tail rows of each permutation are skipped, as in the reference
        lo, hi = D.minibatch_slice(min(self.batch_size, total_steps), self.rank, self.world)
        eta_mean = self.model._eta_value()
        ent_const = -eta_mean  # the entropy term of a fixed-eta policy is a constant (diffusion_ppo.py:183-187)

        fused = not self.use_bc_loss and self.model.fused_update_reason() is None
        with_actor = self.itr >= self.n_critic_warmup_itr and not self._actor_frozen

        # multi-GPU: the actor's gradient segment is reduced next to the critic backward (two collectives, see
        # FlatGradBuffer.allreduce_split); DPPO_B200_OVERLAP=0 keeps the single all-reduce behind the whole backward
        overlap = self.grads.overlap_setup() if (fused and with_actor and os.environ.get("DPPO_B200_OVERLAP", "1") == "1") else None

        def fwd_bwd(inds_b):
            """zero grads -> actor_ft / critic forward -> fused loss kernel -> backward -> gradient all-reduce (no host sync)"""
            self.grads.zero()
            if fused:
                # the whole minibatch inside libdppo_b200 (tcgen05 forward / dgrad / wgrad kernels, dppo_update_minibatch):
                # gradients are accumulated straight into the views of the flat all-reduce buffer
                self.model.update_minibatch(obs_k, chains_k, logprobs_k, ret_k, values_k, adv_k, inds_b, row_begin=lo,
                                            row_count=hi - lo, reward_horizon=self.reward_horizon, vf_coef=self.vf_coef,
                                            with_actor=with_actor, scalars_out=self.grads.scalars,
                                            actor_event=overlap[0] if overlap else None)
                if overlap:
                    self.grads.allreduce_split()
                else:
                    self.grads.allreduce()
                return None
            res = self.model.loss_gathered(obs_k, chains_k, logprobs_k, ret_k, values_k, adv_k, inds_b, row_begin=lo,
                                           row_count=hi - lo, use_bc_loss=self.use_bc_loss,
                                           reward_horizon=self.reward_horizon, scalars_out=self.grads.scalars)
            pg_loss, entropy_loss, v_loss, bc_loss = res[0], res[1], res[2], res[6]
            loss = pg_loss + entropy_loss * self.ent_coef + v_loss * self.vf_coef + bc_loss * self.bc_loss_coeff
            loss.backward()
            self.grads.allreduce()  # gradients + [pg, v, kl, clipfrac, ratio] partial means, one collective
            return bc_loss

        # KL early stop without a host round trip per minibatch (reference :376-382 reads approx_kl on the host after every
        # minibatch): MinibatchStep chains forward / backward / all-reduce, both AdamW steps and dppo_kl_check on the stream;
        # the kernel keeps the diagnostics of minibatch k in history[k] and raises a device flag once approx_kl > target_kl,
        # after which the optimiser launches of LATER minibatches do nothing.  What is applied is exactly what the reference
        # applies (the minibatch that trips the test included), while the host only polls.  The unit is replayed as one CUDA
        # graph when the iteration has enough minibatches to amortise the capture (the rollout buffers get new addresses
        # every iteration, so the graph is per iteration).
        n_max = self.update_epochs * num_batch
        use_target = self.target_kl is not None
        quantile_clip = self.model.clip_advantage_lower_quantile > 0 or self.model.clip_advantage_upper_quantile < 1
        use_graph = (self.cuda_graph_update and not self.use_bc_loss and not quantile_clip and n_max >= 16
                     and total_steps >= self.batch_size)
        step_fn = MinibatchStep(fwd_bwd, self.grads, self.actor_optimizer, self.critic_optimizer, with_actor,
                                self.max_grad_norm, self.target_kl, n_max, self.batch_size, self.device, use_graph=use_graph,
                                world=self.world)
        if use_graph and step_fn.graphed is None:
            log.warning("CUDA-graph capture of the PPO minibatch failed (%s); running eagerly", step_fn.graph_error)
            self.cuda_graph_update = False
        kl_state, history = step_fn.kl_state, step_fn.history
        bc_last, launched, stopped = 0.0, 0, False
        # The host looks at the flag with a fixed lag of LAG minibatches (it waits for the copy issued LAG launches ago, never
        # for the current one): the GPU always has work queued, and every rank takes the same decision at the same minibatch
        # (the all-reduce inside a minibatch needs all ranks to launch the same number of them).
        LAG = 2
        ring = [(torch.zeros(4, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(LAG + 1)]
        for update_epoch in range(self.update_epochs):
            if "perm" in self.test_hooks:
                inds_k = self.test_hooks["perm"](total_steps).to(self.device)
            else:
                inds_k = D.broadcast_permutation(total_steps, self.device)
            in_epoch = 0
            for batch in range(num_batch):
                inds_b = inds_k[batch * self.batch_size:(batch + 1) * self.batch_size]
                bc = step_fn(inds_b)
                bc_last = 0.0 if bc is None else float(bc)
                launched += 1
                in_epoch += 1
                if use_target:
                    buf, ev = ring[launched % (LAG + 1)]
                    buf.copy_(kl_state, non_blocking=True)
                    ev.record()
                    if in_epoch > LAG:
                        old_buf, old_ev = ring[(launched - LAG) % (LAG + 1)]
                        old_ev.synchronize()
                        if int(old_buf[0]):
                            stopped = True
                            break
            if stopped:
                break
            if use_target:  # epoch boundary: exact test, the next permutation is drawn only if the reference would draw it
                if int(kl_state[0].item()):
                    stopped = True
                    break
        torch.cuda.current_stream().synchronize()
        st = kl_state.tolist()
        applied = st[1] + 1 if st[0] else launched  # minibatches whose optimiser steps took effect
        flag_break = bool(st[0])
        hist = history[:applied].cpu().numpy()
        self.last_history = hist
        s = hist[-1]
        stats = dict(pg_loss=float(s[0]), v_loss=float(s[1]), approx_kl=float(s[2]), clipfrac=float(s[3]), ratio=float(s[4]),
                     bc_loss=bc_last, eta=eta_mean,
                     loss=float(s[0]) + ent_const * self.ent_coef + float(s[1]) * self.vf_coef + bc_last * self.bc_loss_coeff)
        clipfracs = [float(v) for v in hist[:, 3]]
        step_fn.finish()  # version counters: the chain kernels repack actor_ft at the next rollout
        y_pred, y_true = values_k.cpu().numpy(), ret_k.cpu().numpy()
        var_y = np.var(y_true)
        stats["explained_var"] = np.nan if var_y == 0 else 1 - np.var(y_true - y_pred) / var_y
        stats["clipfrac"] = float(np.mean(clipfracs))
        stats["minibatches"] = len(clipfracs)
        return stats

    # ------------------------------------------------------------------ main loop (reference :47-483)
    def run(self):
        run_results, cnt_train_step = [], 0
        done_venv = np.zeros((1, self.n_envs))
        prev_obs_venv = None
        t_start = time.perf_counter()
        while self.itr < self.n_train_itr:
            eval_mode = self.itr % self.val_freq == 0 and not self.force_train
            self.model.eval() if eval_mode else self.model.train()
            firsts_trajs = np.zeros((self.n_steps + 1, self.n_envs))
            # the reference resets when reset_at_iteration, in eval mode, or right after an eval iteration; its
            # `last_itr_eval` is overwritten before use (:66), so "eval_mode" covers both; a run that never resets
            # needs an initial reset, which the reference leaves unbound
            if self.reset_at_iteration or eval_mode or prev_obs_venv is None:
                prev_obs_venv = self.reset_env_all()
                firsts_trajs[0] = 1
            else:
                firsts_trajs[0] = done_venv
            t0 = time.perf_counter()
            obs_buf, chains_buf, reward_trajs, terminated_trajs, prev_obs_venv, done_venv, env_steps = self.rollout(
                prev_obs_venv, eval_mode, firsts_trajs)
            torch.cuda.synchronize()
            if self.model.engine(sync=False).nonfinite():  # the reference documents NaN observations from IsaacGym (README.md:184)
                log.warning("itr %d: the sampler produced non-finite actions (NaN / Inf observations or diverged weights)", self.itr)
            t_roll = time.perf_counter() - t0
            cnt_train_step += env_steps * (self.n_envs_global // max(1, self.n_envs)) if self.world > 1 else env_steps

            # episode statistics: episodes that start and finish inside the iteration (:153-193)
            ep_rewards = []
            for e in range(self.n_envs):
                marks = np.where(firsts_trajs[:, e] == 1)[0]
                for a, b in zip(marks[:-1], marks[1:]):
                    if b - a > 1:
                        ep_rewards.append(reward_trajs[a:b, e])
            avg_episode_reward = float(np.mean([r.sum() for r in ep_rewards])) if ep_rewards else 0.0
            avg_best_reward = float(np.mean([r.max() / self.act_steps for r in ep_rewards])) if ep_rewards else 0.0
            success_rate = float(np.mean([r.max() / self.act_steps >= self.best_reward_threshold_for_success
                                          for r in ep_rewards])) if ep_rewards else 0.0

            stats, t_pro, t_upd = None, 0.0, 0.0
            if not eval_mode:
                t0 = time.perf_counter()
                values, logprobs, adv, ret = self.prologue(obs_buf, chains_buf, reward_trajs, terminated_trajs, firsts_trajs,
                                                           prev_obs_venv)
                torch.cuda.synchronize()
                t_pro = time.perf_counter() - t0
                t0 = time.perf_counter()
                stats = self.update(obs_buf, chains_buf, logprobs, values, adv, ret)
                torch.cuda.synchronize()
                t_upd = time.perf_counter() - t0

            if self.itr >= self.n_critic_warmup_itr:
                self._sched["actor"] += 1
            self._sched["critic"] += 1
            self._apply_lr()
            self.model.step()
            if id(self.model.actor_ft) != self._actor_ft_id:
                # ft_denoising_steps was annealed (reference diffusion_vpg.py:102-127): actor_ft is a fresh copy the
                # reference's optimiser never learns about - its parameter list still holds the old tensors, now the frozen
                # base policy, whose gradients stay None, so actor training silently stops.  Same here: no further actor
                # steps (and no weight decay / momentum on the base policy); the gradient buffer follows the new tensors.
                log.warning("ft_denoising_steps annealed to %d: like the reference, the actor optimiser is not rebuilt - "
                            "actor_ft is no longer updated", self.model.ft_denoising_steps)
                self._actor_ft_id, self._actor_frozen = id(self.model.actor_ft), True
                self.grads = D.FlatGradBuffer([list(self.model.actor_ft.parameters()), list(self.model.critic.parameters())])
            if self.itr % self.save_model_freq == 0 or self.itr == self.n_train_itr - 1:
                self.save_model()
            rec = {"itr": self.itr, "step": cnt_train_step, "time": time.perf_counter() - t_start,
                   "t_rollout": t_roll, "t_prologue": t_pro, "t_update": t_upd}
            if eval_mode:
                rec.update(eval_success_rate=success_rate, eval_episode_reward=avg_episode_reward,
                           eval_best_reward=avg_best_reward)
            else:
                rec.update(train_episode_reward=avg_episode_reward, **stats)
            run_results.append(rec)
            if self.rank == 0 and self.itr % self.log_freq == 0:
                log.info("%s", rec)
                with open(self.result_path, "wb") as f:
                    pickle.dump(run_results, f)
            self.itr += 1
        return run_results
