"""
Fused AdamW over flat parameter / gradient segments (libdppo_b200 `dppo_adamw_flat`).

Replaces the two `torch.optim.AdamW(...).step()` calls (+ the optional `clip_grad_norm_`) of the reference's minibatch
loop (/root/reference/dppo/agent/finetune/train_ppo_diffusion_agent.py:360-373; optimisers built at
train_ppo_agent.py:34-53) by ONE kernel per network: the parameters are re-homed into one flat fp32 buffer (their
`.data` become views, names / shapes / state_dict unchanged), the gradients already live in the flat all-reduce buffer
(dppo_b200.distributed.FlatGradBuffer), and exp_avg / exp_avg_sq are flat buffers of the same length.
"""

import ctypes as C

import torch

from dppo_b200 import _lib


class FlatAdamW:
    """AdamW(lr, betas, eps, weight_decay) with torch's update rule; `param_groups[0]["lr"]` is honoured every step."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdamW: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW runs on CUDA devices only; there is no CPU path")
        self.lib = _lib.load()
        # every tensor starts on a 16-byte boundary (float4 loads in the kernels that read the parameters in place);
        # FlatGradBuffer lays the gradients out with the same padding, so one index addresses p, g, m and v
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        self.n = n
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        with torch.no_grad():
            for p in self.params:
                if p.dtype != torch.float32:
                    raise RuntimeError("FlatAdamW needs fp32 parameters")
                k = p.numel()
                self.flat[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + k].view_as(p)
                torch.autograd.graph.increment_version(p)  # derived weight caches key on the version counter
                o += (k + 3) // 4 * 4
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self._ws = torch.zeros(2, dtype=torch.float64, device=dev)
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        # the step count lives on the device ([0] = count, [2..3] = bias corrections of the current step) so that a step
        # captured in a CUDA graph advances its own bias correction on every replay (dppo_adamw_flat_dev)
        self._state = torch.zeros(4, dtype=torch.int32, device=dev)
        self._grad_flat = None

    @property
    def step_count(self):
        return int(self._state[0].item())

    @step_count.setter
    def step_count(self, v):
        self._state[0] = int(v)

    def _grads(self):
        """Device address of the contiguous gradient segment (views handed out by FlatGradBuffer); the layout is verified
        once per buffer.  Only the address is cached - no tensor that would keep a (possibly NCCL-registered) gradient
        buffer alive behind FlatGradBuffer.release()."""
        first = self.params[0].grad
        if first is None:
            raise RuntimeError("FlatAdamW.step: parameters have no gradients")
        if self._grad_flat == first.data_ptr():
            return self._grad_flat
        ptr = first.data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != ptr or not p.grad.is_contiguous():
                raise RuntimeError("FlatAdamW needs gradients laid out contiguously in parameter order (FlatGradBuffer)")
            ptr += (p.numel() + 3) // 4 * 16
        self._grad_flat = first.data_ptr()
        return self._grad_flat

    @torch.no_grad()
    def step(self, max_grad_norm=None, stop_flag=None, bump=True):
        """One AdamW step (stream-ordered, no host synchronisation, CUDA-graph capturable).  `stop_flag`: optional device
        int32 tensor; once it is non-zero the launch is a no-op and the step count does not advance (KL early stop).
        `bump=False` leaves the parameters' version counters alone (the caller bumps them after a graph replay)."""
        g = self._grads()
        if g % 16:
            raise RuntimeError("FlatAdamW: gradient segment must be 16-byte aligned")
        hp = self.param_groups[0]
        _lib.check(
            self.lib.dppo_adamw_flat_dev(_lib.ptr(self.flat), C.c_void_p(g), _lib.ptr(self.exp_avg),
                                         _lib.ptr(self.exp_avg_sq), self.n, float(hp["lr"]), float(hp["betas"][0]),
                                         float(hp["betas"][1]), float(hp["eps"]), float(hp["weight_decay"]),
                                         _lib.ptr(self._state), _lib.ptr(stop_flag),
                                         -1.0 if max_grad_norm is None else float(max_grad_norm), _lib.ptr(self._ws),
                                         _lib.stream_ptr()),
            "dppo_adamw_flat_dev")
        if bump:
            self.bump_versions()

    def bump_versions(self):
        """The packed weight caches (chain kernels) key on Tensor._version: mark the parameters as changed."""
        for p in self.params:
            torch.autograd.graph.increment_version(p)

    def zero_grad(self, set_to_none=False):
        for p in self.params:
            if p.grad is not None:
                p.grad.zero_()

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(),
                    param_groups=[dict(g) for g in self.param_groups])

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, h in zip(self.param_groups, sd["param_groups"]):
            g.update(h)
