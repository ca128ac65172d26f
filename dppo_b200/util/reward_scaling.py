"""
Running reward scaling: rewards are divided by the standard deviation of a rolling discounted sum of rewards.

RunningRewardScaler -> /root/reference/dppo/util/reward_scaling.py:42-87 (state: per-env running return `ret`, scalar
running mean / var / count; float64 on the host like the reference - (E, n_steps) values once per iteration, the step
that feeds GAE).  State is exposed through state_dict() so a resumed run continues the statistics.
"""

import numpy as np


class RunningRewardScaler:
    def __init__(self, num_envs, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        self.ret = np.zeros(num_envs)
        self.mean, self.var, self.count = 0.0, 1.0, 1e-4
        self.cliprew, self.gamma, self.epsilon = cliprew, gamma, epsilon

    def __call__(self, reward, first):
        """reward, first: (E, n_steps) float64.  Returns the scaled rewards, same shape."""
        E, n = reward.shape
        rets = np.zeros_like(reward)
        prev = self.ret
        for t in range(n):
            prev = rets[:, t] = reward[:, t] + (1.0 - first[:, t]) * self.gamma * prev
        self.ret = rets[:, -1]
        flat = rets.reshape(-1)
        b_mean, b_var, b_n = flat.mean(), flat.var(), flat.shape[0]
        delta, tot = b_mean - self.mean, self.count + b_n
        m2 = self.var * self.count + b_var * b_n + delta ** 2 * self.count * b_n / tot
        self.mean = self.mean + delta * b_n / tot
        self.var = m2 / (tot - 1)
        self.count = tot
        return np.clip(reward / np.sqrt(self.var + self.epsilon), -self.cliprew, self.cliprew)

    def state_dict(self):
        return dict(ret=self.ret.copy(), mean=self.mean, var=self.var, count=self.count)

    def load_state_dict(self, s):
        self.ret, self.mean, self.var, self.count = np.array(s["ret"]), s["mean"], s["var"], s["count"]
