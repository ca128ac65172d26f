"""
Host-side wrapper of `dppo_update_*` (include/dppo_b200.h): the PPO minibatch of a DiffusionMLP actor_ft and a
residual-MLP critic on hand-written tcgen05 kernels (csrc/update_gemm.cu, csrc/update_plan.cu) instead of torch autograd.

Replaces what the reference runs per minibatch through autograd: get_logprobs_subsample -> actor_ft forward
(/root/reference/dppo/model/diffusion/diffusion_vpg.py:398-461), CriticObs.forward (dppo/model/common/critic.py:40-54) and
loss.backward() (dppo/agent/finetune/train_ppo_diffusion_agent.py:360-364).  There is no CPU path.
"""

import ctypes as C

import torch

from dppo_b200 import _lib
from dppo_b200.engine import _mlp_param_list, mlp_desc_of


def critic_param_list(critic):
    """Parameters of a residual-style CriticObs in the order dppo_update_bind expects."""
    q = critic.Q1
    layers = q.layers
    nb = q.n_blocks
    ps = [layers[0].weight, layers[0].bias]
    for b in range(1, nb + 1):
        blk = layers[b]
        ps += [blk.l1.weight, blk.l1.bias, blk.l2.weight, blk.l2.bias]
        if hasattr(blk, "norm1"):
            ps += [blk.norm1.weight, blk.norm1.bias, blk.norm2.weight, blk.norm2.bias]
    ps += [layers[nb + 1].weight, layers[nb + 1].bias]
    return ps


def unsupported_reason(model):
    """None when the tensor-core update path covers this model, else why not (the caller keeps the autograd path)."""
    if type(model.actor_ft).__name__ != "DiffusionMLP":
        return "actor is not a DiffusionMLP"
    q = getattr(model.critic, "Q1", None)
    if q is None or type(q).__name__ != "ResidualMLP":
        return "critic is not a residual MLP"
    if not isinstance(q.layers[-1], torch.nn.Identity) or isinstance(q.layers[-2], torch.nn.LayerNorm):
        return "critic output stage"
    if q.activation_type not in ("ReLU", "Mish"):
        return f"critic activation {q.activation_type}"
    if q.layers[q.n_blocks + 1].out_features != 1:
        return "critic output width"
    if getattr(model, "learn_eta", False):
        return "learned eta"
    if model.ft_denoising_steps < 1:
        return "no fine-tuned denoising step"
    # the geometry limits of dppo_update_create (csrc/update_plan.cu: 64-wide K chunks and output tiles, LayerNorm
    # epilogues written for Mish) - such models keep the autograd route instead of failing at the first minibatch
    desc = mlp_desc_of(model.actor_ft)
    widths = {"actor hidden": desc.hidden_dim, "critic hidden": q.hidden_dim}
    if desc.cond_hidden:
        widths["cond_mlp hidden"], widths["cond_mlp output"] = desc.cond_hidden, desc.cond_out
    for what, n in widths.items():
        if n % 64:
            return f"{what} width {n} is not a multiple of 64"
    if (desc.use_layernorm and desc.activation != _lib.ACT_MISH) or (q.use_layernorm and q.activation_type != "Mish"):
        return "LayerNorm with an activation other than Mish"
    return None


class UpdatePlan:
    """Workspace + kernel program for minibatches of up to `max_rows` rows of one model."""

    def __init__(self, model, max_rows):
        why = unsupported_reason(model)
        if why is not None:
            raise NotImplementedError(f"tensor-core update path: {why}")
        self.lib = _lib.load()
        self.engine = model.engine(sync=False)
        self.max_rows = int(max_rows)
        self.D = model.horizon_steps * model.action_dim
        q = model.critic.Q1
        d = _lib.ResMlpDesc()
        d.in_dim, d.hidden_dim, d.n_blocks = q.layers[0].in_features, q.hidden_dim, q.n_blocks
        d.out_dim, d.activation, d.use_layernorm = 1, {"ReLU": _lib.ACT_RELU, "Mish": _lib.ACT_MISH}[q.activation_type], int(q.use_layernorm)
        self.handle = C.c_void_p()
        _lib.check(self.lib.dppo_update_create(C.byref(self.handle), self.engine.ctx, C.byref(d), self.max_rows), "dppo_update_create")
        self._sig = None
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dppo_update_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ binding
    def bind(self, actor_params, actor_grads, critic_params, critic_grads):
        """Device pointers of the fp32 parameters and of the tensors their gradients are accumulated into."""
        sig = tuple(t.data_ptr() for t in list(actor_params) + list(actor_grads) + list(critic_params) + list(critic_grads))
        if sig == self._sig:
            return
        for t in list(actor_params) + list(actor_grads) + list(critic_params) + list(critic_grads):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("dppo_update_bind needs contiguous fp32 CUDA tensors")
        arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        ap, ag, cp, cg = arr(actor_params), arr(actor_grads), arr(critic_params), arr(critic_grads)
        _lib.check(self.lib.dppo_update_bind(self.handle, ap, ag, len(actor_params), cp, cg, len(critic_params)), "dppo_update_bind")
        self._sig = sig
        self._keep = (list(actor_params), list(actor_grads), list(critic_params), list(critic_grads))

    def bind_model(self, model):
        """Parameters of model.actor_ft / model.critic with their .grad tensors (e.g. views of a FlatGradBuffer)."""
        ap, cp = _mlp_param_list(model.actor_ft), critic_param_list(model.critic)
        for p in ap + cp:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        self.bind(ap, [p.grad for p in ap], cp, [p.grad for p in cp])

    # ------------------------------------------------------------------ calls
    @staticmethod
    def _batch(n_rows, global_rows, row_begin=0, **t):
        b = _lib.UpdateBatch()
        for k in ("obs", "chains", "x_next", "old_logprobs", "returns", "old_values", "advantages", "inds_all", "denoising_inds"):
            v = t.get(k)
            setattr(b, k, None if v is None else v.data_ptr())
        b.row_begin, b.n_rows, b.global_rows = int(row_begin), int(n_rows), int(global_rows)
        return b

    def forward(self, batch, eps_out=None, vpred_out=None):
        _lib.check(self.lib.dppo_update_forward(self.handle, C.byref(batch), _lib.ptr(eps_out), _lib.ptr(vpred_out),
                                                _lib.stream_ptr()), "dppo_update_forward")

    def set_actor_event(self, event):
        """`event`: a torch.cuda.Event that has been recorded once (so that it owns a cudaEvent_t), or None.  The backward
        records it behind the actor backward (see dppo_update_set_actor_event)."""
        handle = None if event is None else C.c_void_p(event.cuda_event)
        _lib.check(self.lib.dppo_update_set_actor_event(self.handle, handle), "dppo_update_set_actor_event")
        self._actor_event = event  # keep it alive

    def values(self, obs, out=None):
        """critic(obs) for (rows, cond_dim) observation rows, rows <= max_rows (bind_model first)."""
        rows = obs.shape[0]
        out = torch.empty(rows, dtype=torch.float32, device=obs.device) if out is None else out
        _lib.check(self.lib.dppo_update_values(self.handle, _lib.ptr(obs), rows, _lib.ptr(out), _lib.stream_ptr()),
                   "dppo_update_values")
        return out

    def backward(self, grad_eps, grad_v, scale_pg=None, scale_v=None, vf_coef=1.0, with_actor=True, with_critic=True):
        _lib.check(self.lib.dppo_update_backward(self.handle, _lib.ptr(grad_eps), _lib.ptr(grad_v), _lib.ptr(scale_pg),
                                                 _lib.ptr(scale_v), float(vf_coef), int(with_actor), int(with_critic),
                                                 _lib.stream_ptr()), "dppo_update_backward")

    def minibatch(self, batch, hp, vf_coef, with_actor, scalars, workspace):
        _lib.check(self.lib.dppo_update_minibatch(self.handle, C.byref(batch), C.byref(hp), float(vf_coef), int(with_actor),
                                                  _lib.ptr(scalars), _lib.ptr(workspace), _lib.stream_ptr()),
                   "dppo_update_minibatch")


def _c32(t, shape=None):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    if shape is not None:
        t = t.reshape(shape)
    return t.contiguous()


class FusedLoss(torch.autograd.Function):
    """PPODiffusion.loss on the tensor-core update path, as an autograd node: (parameters) -> (pg_loss, v_loss).

    forward: dppo_update_forward + dppo_ppo_loss_rows; backward: dppo_update_backward with autograd's incoming factors as
    device scalars, gradients returned per parameter tensor (autograd accumulates them into .grad)."""

    @staticmethod
    def forward(ctx, model, plan, hp, tensors, n_actor, *params):
        eng = plan.engine
        B, D = tensors["denoising_inds"].shape[0], plan.D
        dev = tensors["chains"].device
        flat = torch.zeros(sum((p.numel() + 3) // 4 * 4 for p in params), dtype=torch.float32, device=dev)
        views, o = [], 0
        for p in params:
            views.append(flat[o:o + p.numel()].view_as(p))
            o += (p.numel() + 3) // 4 * 4
        data = [p.detach() for p in params]
        plan.bind(data[:n_actor], views[:n_actor], data[n_actor:], views[n_actor:])
        eps = torch.empty((B, D), dtype=torch.float32, device=dev)
        vpred = torch.empty(B, dtype=torch.float32, device=dev)
        batch = UpdatePlan._batch(B, B, **tensors)
        plan.forward(batch, eps, vpred)
        grad_eps, grad_v, scalars = eng.loss_rows(hp, tensors["chains"], tensors["x_next"], tensors["old_logprobs"],
                                                  tensors["returns"], tensors["old_values"], tensors["advantages"],
                                                  tensors["denoising_inds"], eps, vpred)
        ctx.plan, ctx.views, ctx.flat = plan, views, flat
        ctx.save_for_backward(grad_eps, grad_v)
        out_scalars = scalars.clone()
        ctx.mark_non_differentiable(out_scalars)
        return scalars[0].clone(), scalars[1].clone(), out_scalars

    @staticmethod
    def backward(ctx, g_pg, g_v, _):
        grad_eps, grad_v = ctx.saved_tensors
        dev = grad_eps.device
        one = lambda g: torch.zeros((), device=dev) if g is None else g.detach().float().reshape(()).contiguous()  # noqa: E731
        ctx.flat.zero_()
        ctx.plan.backward(grad_eps, grad_v, scale_pg=one(g_pg), scale_v=one(g_v))
        return (None, None, None, None, None) + tuple(ctx.views)
