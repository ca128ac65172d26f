"""
Evaluation loop for a pre-trained / DPPO-fine-tuned diffusion policy.

EvalAgent / EvalDiffusionAgent -> /root/reference/dppo/agent/eval/eval_agent.py:20-125,
/root/reference/dppo/agent/eval/eval_diffusion_agent.py:17-145.  Same config keys (`n_steps`, `env.*`, `model`, `act_steps`
...), same episode accounting (only episodes that start and finish inside the run count; best reward = max step reward /
act_steps; success = best reward >= threshold) and the same `eval.npz` fields.  Actions come from
`self.model(cond=..., deterministic=True)`, i.e. the chain kernel; only the action chunk leaves the GPU each step.
Simulators are out of scope: without a `venv` argument the synthetic vector env with the `make_async` contract is used.
"""

import logging
import os
import random
import time

import numpy as np
import torch

from dppo_b200.util.config import instantiate

log = logging.getLogger(__name__)


class EvalDiffusionAgent:
    def __init__(self, cfg, venv=None):
        self.cfg = cfg
        self.device = torch.device(cfg.device)
        self.seed = cfg.get("seed", 42)
        random.seed(self.seed)
        np.random.seed(self.seed)
        torch.manual_seed(self.seed)
        self.n_envs = cfg.env.n_envs
        self.n_cond_step, self.obs_dim, self.action_dim = cfg.cond_steps, cfg.obs_dim, cfg.action_dim
        self.act_steps, self.horizon_steps = cfg.act_steps, cfg.horizon_steps
        self.max_episode_steps = cfg.env.max_episode_steps
        if venv is None:
            from dppo_b200.env.synthetic import SyntheticVecEnv

            venv = SyntheticVecEnv(self.n_envs, self.obs_dim, self.action_dim, self.n_cond_step, self.act_steps,
                                   self.max_episode_steps, seed=self.seed)
        self.venv = venv
        self.venv.seed([self.seed + i for i in range(self.n_envs)])
        specific = cfg.env.get("specific", None)
        self.furniture_sparse_reward = bool(specific.get("sparse_reward", False)) if specific else False
        self.model = instantiate(cfg.model)
        self.n_steps = cfg.n_steps
        pairs = getattr(self.venv, "pairs_to_assemble", None)
        self.best_reward_threshold_for_success = len(pairs) if pairs is not None else cfg.env.best_reward_threshold_for_success
        self.logdir = cfg.logdir
        os.makedirs(self.logdir, exist_ok=True)
        self.result_path = os.path.join(self.logdir, "eval.npz")

    def reset_env_all(self):
        obs = self.venv.reset_arg(options_list=[{} for _ in range(self.n_envs)])
        if isinstance(obs, list):
            obs = {k: np.stack([o[k] for o in obs]) for k in obs[0]}
        return obs

    def run(self):
        t_start = time.perf_counter()
        self.model.eval()
        firsts = np.zeros((self.n_steps + 1, self.n_envs))
        rewards = np.zeros((self.n_steps, self.n_envs))
        obs = self.reset_env_all()
        firsts[0] = 1
        for step in range(self.n_steps):
            # host observations in, host actions out: one library call per decision (dppo_sample_chain_host), the action
            # chunk is in page-locked host memory when it returns
            state = torch.from_numpy(np.ascontiguousarray(obs["state"], dtype=np.float32))
            samples = self.model(cond={"state": state}, deterministic=True)
            obs, reward, terminated, truncated, _ = self.venv.step(samples.trajectories.numpy()[:, : self.act_steps])
            rewards[step] = reward
            firsts[step + 1] = terminated | truncated
        # episodes that start and finish inside the run
        episodes = []
        for e in range(self.n_envs):
            marks = np.where(firsts[:, e] == 1)[0]
            episodes += [rewards[a:b, e] for a, b in zip(marks[:-1], marks[1:]) if b - a > 1]
        if episodes:
            ep_reward = np.array([r.sum() for r in episodes])
            best = ep_reward if self.furniture_sparse_reward else np.array([r.max() / self.act_steps for r in episodes])
            res = dict(num_episode=len(episodes), eval_success_rate=float(np.mean(best >= self.best_reward_threshold_for_success)),
                       eval_episode_reward=float(ep_reward.mean()), eval_best_reward=float(best.mean()))
        else:
            log.info("[WARNING] No episode completed within the iteration!")
            res = dict(num_episode=0, eval_success_rate=0.0, eval_episode_reward=0.0, eval_best_reward=0.0)
        res["time"] = time.perf_counter() - t_start
        log.info("eval: num episode %4d | success rate %8.4f | avg episode reward %8.4f | avg best reward %8.4f",
                 res["num_episode"], res["eval_success_rate"], res["eval_episode_reward"], res["eval_best_reward"])
        np.savez(self.result_path, **res)
        return res
