"""
nn.Linear whose CUDA forward / backward GEMMs run on the bf16 tensor cores at fp32-grade precision.

The update path of DPPO (PPODiffusion.loss -> get_logprobs_subsample -> actor_ft, CriticObs.forward, loss.backward();
/root/reference/dppo/model/diffusion/diffusion_vpg.py:398-461, /root/reference/dppo/model/common/critic.py:40-54,
/root/reference/dppo/agent/finetune/train_ppo_diffusion_agent.py:360-364) is 84 % fp32 SIMT GEMMs under torch eager
(measured: 5.5 of 6.5 ms of GPU time per 50 000-row minibatch of cfg2).  Here every Linear is evaluated with the same
3-product split the chain kernel uses, x w ~= x_hi w_hi + x_hi w_lo + x_lo w_hi (bf16 halves, fp32 accumulation):
`dppo_split3_pack` (libdppo_b200) writes [hi | hi | lo] / [hi | lo | hi] operand rows in one pass, and ONE bf16 GEMM
over the 3K-long rows (a plain library GEMM: torch.mm(..., out_dtype=float32) -> cuBLASLt) is the whole product.
forward: y = [x | 1] [W | b]^T (the bias rides along as one more K column);  backward: dx = dy W (same trick on dy and
W^T), [dW | db] = dy^T [x | 1] (three small-output GEMMs on the hi / lo column blocks of the packed operands).  Parameters, names and state_dict are nn.Linear's.
CPU tensors are refused unless the host-logic tests switch CPU_TEST_HOOK on (then plain F.linear).
"""

import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from dppo_b200 import _lib

ENABLED = True  # module-level switch (tests compare both paths)
# Test hook, OFF in production: evaluating these modules on CPU tensors (stock torch kernels) is something only the host
# logic tests do (tests/conftest.py switches it on); a product run that ends up with CPU tensors here fails loudly.
CPU_TEST_HOOK = False


def require_cuda(x, who):
    if not x.is_cuda and not CPU_TEST_HOOK:
        raise RuntimeError(f"{who}: CPU tensors - dppo_b200 has no CPU path (tests enable split_linear.CPU_TEST_HOOK explicitly)")


def _pack(x2d, pattern, ones=False, extra=None):
    """fp32 [M, K] (unit column stride) -> bf16 [M, 3 * Kp]; `ones` / `extra` append one more column before the padding"""
    M, K = x2d.shape
    if x2d.stride(1) != 1:
        x2d = x2d.contiguous()
    mode = 1 if ones else (2 if extra is not None else 0)
    Kp = (K + (1 if mode else 0) + 7) // 8 * 8
    out = torch.empty((M, 3 * Kp), dtype=torch.bfloat16, device=x2d.device)
    _lib.check(_lib.load().dppo_split3_pack(C.c_void_p(x2d.data_ptr()), M, K, x2d.stride(0), mode, _lib.ptr(extra),
                                            _lib.ptr(out), pattern, _lib.stream_ptr()), "dppo_split3_pack")
    return out, Kp


class _Split3Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        # the bias is one more K column: [x | 1] . [W | b]^T
        xp, Kp = _pack(x2, 0, ones=bias is not None)
        wp, _ = _pack(weight, 1, extra=None if bias is None else bias.contiguous())
        y = torch.mm(xp, wp.t(), out_dtype=torch.float32)
        ctx.save_for_backward(xp, weight)
        ctx.meta = (shape, Kp, bias is not None)
        return y.view(*shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        xp, weight = ctx.saved_tensors
        shape, Kp, has_bias = ctx.meta
        N, K = weight.shape
        gy2 = gy.reshape(-1, N)
        gyp, Np = _pack(gy2, 0)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            wtp, _ = _pack(weight.t(), 1)  # [K, 3 Np]
            gx = torch.mm(gyp, wtp.t(), out_dtype=torch.float32).view(shape)
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gyh, gyl = gyp[:, :Np], gyp[:, 2 * Np:]
            xh, xl = xp[:, :Kp], xp[:, 2 * Kp:]
            gwb = torch.mm(gyh.t(), xh, out_dtype=torch.float32)
            gwb += torch.mm(gyh.t(), xl, out_dtype=torch.float32)
            gwb += torch.mm(gyl.t(), xh, out_dtype=torch.float32)
            gw = gwb[:N, :K]
            if has_bias:
                gb = gwb[:N, K]  # the ones column of [x | 1]: column sums of dy
        return gx, gw, gb


class SplitLinear(nn.Linear):
    """Drop-in nn.Linear (same parameters / state_dict); CUDA fp32 inputs take the split-3 tensor-core path."""

    def forward(self, x):
        require_cuda(x, "SplitLinear")
        # (a 1-wide value head is a GEMV: not worth three operand passes, and an odd leading dimension suits no tensor-core tile)
        if ENABLED and x.is_cuda and x.dtype == torch.float32 and self.weight.dtype == torch.float32 and self.out_features >= 8:
            return _Split3Linear.apply(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)
