"""
DDIM eta (host side). Only the fixed variant is on the hot path (every PPO YAML sets learn_eta: False).

EtaFixed -> /root/reference/dppo/model/diffusion/eta.py:12-40
"""

import torch


class EtaFixed(torch.nn.Module):
    def __init__(self, base_eta=0.5, min_eta=0.1, max_eta=1.0, **kwargs):
        super().__init__()
        self.eta_logit = torch.nn.Parameter(torch.ones(1))
        self.min = min_eta
        self.max = max_eta
        self.eta_logit.data = torch.atanh(torch.tensor([2 * (base_eta - min_eta) / (max_eta - min_eta) - 1]))

    def value(self) -> torch.Tensor:
        """0-d fp32 tensor, same arithmetic as the reference (tanh -> affine map to [min, max])."""
        return (0.5 * (torch.tanh(self.eta_logit) + 1) * (self.max - self.min) + self.min).reshape(())

    def __call__(self, cond):
        ref = cond["state"] if "state" in cond else cond["rgb"]
        return torch.full((len(ref), 1), self.value().item()).to(ref.device)
