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


class RunningRewardScalerCUDA:
    """
    The same statistics on the device (libdppo_b200 `dppo_reward_scale_f64`): rewards / firsts arrive as float64 CUDA
    tensors of shape (n_steps, E) - the layout of the GAE kernel the scaled rewards feed - and never return to the host.
    With torch.distributed initialised (env-sharded ranks) the two batch sums are all-reduced, so every rank holds the
    single-process running mean / variance.  state_dict() is interchangeable with RunningRewardScaler's.
    """

    def __init__(self, num_envs, device, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        import torch

        from dppo_b200 import _lib

        self._torch, self._lib_mod, self.lib = torch, _lib, _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RunningRewardScalerCUDA needs a CUDA device; the host mirror is RunningRewardScaler")
        self.ret = torch.zeros(num_envs, dtype=torch.float64, device=self.device)
        self.stats = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device=self.device)  # mean, var, count
        self._ws = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.cliprew, self.gamma, self.epsilon = cliprew, gamma, epsilon

    def __call__(self, reward, first):
        """reward, first: (n_steps, E) float64 CUDA tensors.  Returns the scaled rewards, same shape, on the device."""
        torch, L = self._torch, self._lib_mod
        import torch.distributed as dist

        n, E = reward.shape
        reward, first = reward.contiguous().to(torch.float64), first.contiguous().to(torch.float64)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        n_global = torch.tensor([n * E], dtype=torch.int64, device=self.device)
        if world > 1:
            dist.all_reduce(n_global)
        n_global = int(n_global.item()) if world > 1 else n * E
        rets, out = torch.empty_like(reward), torch.empty_like(reward)
        for phase in range(3):
            L.check(self.lib.dppo_reward_scale_f64(L.ptr(reward), L.ptr(first), n, E, n_global, self.gamma, self.epsilon,
                                                   self.cliprew, L.ptr(self.ret), L.ptr(self.stats), L.ptr(rets),
                                                   L.ptr(self._ws), L.ptr(out), phase, L.stream_ptr()),
                    "dppo_reward_scale_f64")
            if world > 1 and phase < 2:
                dist.all_reduce(self._ws[phase:phase + 1])
        return out

    def state_dict(self):
        m, v, c = self.stats.tolist()
        return dict(ret=self.ret.cpu().numpy().copy(), mean=m, var=v, count=c)

    def load_state_dict(self, s):
        torch = self._torch
        self.ret.copy_(torch.as_tensor(np.asarray(s["ret"]), dtype=torch.float64))
        self.stats.copy_(torch.tensor([s["mean"], s["var"], s["count"]], dtype=torch.float64))
