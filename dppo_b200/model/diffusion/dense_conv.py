"""
Conv1d / ConvTranspose1d over a short action horizon evaluated as the dense map they are, on the tensor cores.

Update-path counterpart of csrc/unet_plan.cu: for Ta = 4 a Conv1d(k = 5, pad = 2) is a 14/16-dense block-Toeplitz matrix
[C_out * T_out, C_in * T_in].  cuDNN runs these tiny convolutions as fp32 kernels (cfg5: 30 ms per 10 000-row minibatch,
45x the per-row time of the MLP configs); here the dense matrix is gathered from the conv weight (differentiable: its
gradient scatters back onto the taps), and the product runs through the same 3-product bf16-split tensor-core GEMM as
every Linear (model/common/split_linear.py), forward, dgrad and wgrad.  Modules keep nn.Conv1d / nn.ConvTranspose1d
parameters, names and state_dict (reference modules: /root/reference/dppo/model/diffusion/modules.py:30-95,
unet.py:27-118); on CPU tensors, or for sequences longer than MAX_LEN, they are the stock convolutions.
"""

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from dppo_b200.model.common import split_linear as SL

MAX_LEN = 16          # longest sequence lowered (the dense matrix grows with T^2)
MAX_DENSE = 1 << 21   # largest dense matrix (elements)


def _lowering(kind, c_in, c_out, ks, stride, pad, t_in):
    """index / mask arrays of the dense matrix: D[co * T_out + to, ci * T_in + ti] = W_flat[idx] * mask"""
    if kind == "conv":
        t_out = (t_in + 2 * pad - ks) // stride + 1
    else:
        t_out = (t_in - 1) * stride - 2 * pad + ks
    co, to, ci, ti = np.meshgrid(np.arange(c_out), np.arange(t_out), np.arange(c_in), np.arange(t_in), indexing="ij")
    if kind == "conv":   # y[co, to] = sum W[co, ci, q] x[ci, to * stride + q - pad]
        q = ti - to * stride + pad
        flat = (co * c_in + ci) * ks + q
    else:                # ConvTranspose1d, weight [C_in, C_out, ks]: y[co, ti * stride + q - pad] += W[ci, co, q] x[ci, ti]
        q = to - ti * stride + pad
        flat = (ci * c_out + co) * ks + q
    ok = (q >= 0) & (q < ks)
    idx = np.where(ok, flat, 0).reshape(c_out * t_out, c_in * t_in)
    mask = ok.reshape(c_out * t_out, c_in * t_in)
    # inverse map for the backward: the dense positions every weight tap appears at (at most max(T) of them), padded with
    # the index of an appended zero -> the weight gradient is a GATHER + row sum instead of an atomic scatter
    n_w, n_d = c_in * c_out * ks, idx.size
    pos = np.nonzero(mask.reshape(-1))[0]
    taps = idx.reshape(-1)[pos]
    order = np.argsort(taps, kind="stable")
    taps, pos = taps[order], pos[order]
    counts = np.bincount(taps, minlength=n_w)
    width = int(counts.max()) if counts.size else 1
    inv = np.full((n_w, width), n_d, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    inv[taps, np.arange(taps.size) - start[taps]] = pos
    return t_out, idx.astype(np.int64), mask.astype(np.float32), inv


class _GatherDense(torch.autograd.Function):
    """dense = W_flat[idx] * mask; backward gathers the dense gradient back onto the taps (no atomics)."""

    @staticmethod
    def forward(ctx, weight, idx, mask, inv):
        ctx.save_for_backward(inv)
        ctx.wshape = weight.shape
        return weight.reshape(-1)[idx] * mask

    @staticmethod
    def backward(ctx, g):
        (inv,) = ctx.saved_tensors
        gp = torch.cat([g.reshape(-1), g.new_zeros(1)])
        return gp[inv].sum(-1).view(ctx.wshape), None, None, None


class _DenseMixin:
    def _dense_forward(self, x, kind):
        B, c_in, t_in = x.shape
        c_out = self.out_channels
        key = (t_in, x.device)
        cache = self.__dict__.setdefault("_lowerings", {})
        if key not in cache:
            t_out, idx, mask, inv = _lowering(kind, c_in, c_out, self.kernel_size[0], self.stride[0], self.padding[0], t_in)
            cache[key] = (t_out, torch.from_numpy(idx).to(x.device), torch.from_numpy(mask).to(x.device),
                          torch.from_numpy(inv).to(x.device))
        t_out, idx, mask, inv = cache[key]
        dense = _GatherDense.apply(self.weight, idx, mask, inv)
        bias = None if self.bias is None else self.bias.repeat_interleave(t_out)
        y = SL._Split3Linear.apply(x.reshape(B, c_in * t_in), dense, bias)
        return y.view(B, c_out, t_out)

    def _lowerable(self, x):
        if not (SL.ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[-1] <= MAX_LEN):
            return False
        if self.groups != 1 or self.dilation[0] != 1 or isinstance(self.padding, str):
            return False
        return self.in_channels * x.shape[-1] * self.out_channels * x.shape[-1] * 2 <= MAX_DENSE


class DenseConv1d(_DenseMixin, nn.Conv1d):
    def forward(self, x):
        if self._lowerable(x) and self.out_channels * ((x.shape[-1] + 2 * self.padding[0] - self.kernel_size[0]) // self.stride[0] + 1) >= 8:
            return self._dense_forward(x, "conv")
        return F.conv1d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class DenseConvTranspose1d(_DenseMixin, nn.ConvTranspose1d):
    def forward(self, x):
        if self._lowerable(x) and self.output_padding[0] == 0:
            return self._dense_forward(x, "convT")
        return F.conv_transpose1d(x, self.weight, self.bias, self.stride, self.padding, self.output_padding, self.groups,
                                  self.dilation)
