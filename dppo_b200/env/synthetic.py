"""
Synthetic vector-env stub with the contract the agent loop expects from `make_async(...)` (reference
dppo/env/gym_utils/__init__.py and the MultiStep wrapper, dppo/env/gym_utils/wrapper/multi_step.py:82-192):

    reset_arg(options_list)              -> {"state": (E, To, Do) float32}
    step(action (E, act_steps, Da))      -> ({"state": (E, To, Do) f32}, reward (E,), terminated (E,) bool,
                                             truncated (E,) bool, info list[dict])
    seed(list)

MuJoCo / IsaacGym are not available offline: observations are fresh U(-1, 1) draws (normalised
observations live in [-1, 1], mujoco_locomotion_lowdim.py:57-58), rewards N(0, 1), terminated ~ Bernoulli(p_term),
truncation after max_episode_steps / act_steps decisions; a finished env restarts within the step (reset_within_step).
"""

import numpy as np


class SyntheticVecEnv:
    def __init__(self, n_envs, obs_dim, action_dim, cond_steps=1, act_steps=4, max_episode_steps=1000, p_term=0.01,
                 seed=0, env_offset=0):
        self.n_envs, self.obs_dim, self.action_dim = n_envs, obs_dim, action_dim
        self.cond_steps, self.act_steps = cond_steps, act_steps
        self.max_decisions = max(1, max_episode_steps // act_steps)
        self.p_term = p_term
        self.env_offset = env_offset
        self.seed([seed + env_offset + i for i in range(n_envs)])

    def seed(self, seeds):
        # one stream per env so that an env-sharded run draws what the single-process run draws for the same envs
        self.rngs = [np.random.default_rng(int(s)) for s in seeds]
        self.t = np.zeros(self.n_envs, dtype=np.int64)

    def _obs(self):
        return np.stack([r.uniform(-1, 1, (self.cond_steps, self.obs_dim)) for r in self.rngs]).astype(np.float32)

    def reset_arg(self, options_list=None):
        self.t[:] = 0
        return {"state": self._obs()}

    def step(self, action):
        assert action.shape == (self.n_envs, self.act_steps, self.action_dim), action.shape
        self.t += 1
        reward = np.array([r.standard_normal() for r in self.rngs])
        terminated = np.array([r.uniform() < self.p_term for r in self.rngs])
        truncated = self.t >= self.max_decisions
        self.t[terminated | truncated] = 0
        return {"state": self._obs()}, reward, terminated, truncated, [{} for _ in range(self.n_envs)]
