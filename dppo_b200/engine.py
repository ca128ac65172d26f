"""
Host-side wrapper of one libdppo_b200 context: schedule upload, weight-cache coherence, kernel calls.

The packed (swizzled bf16 hi/lo) weight tiles are derived caches of the nn.Parameters; they are refreshed whenever a
parameter's `_version` (or identity) changes, i.e. after every optimiser step on actor_ft and after
VPGDiffusion.step() swaps actor <- actor_ft (reference dppo/model/diffusion/diffusion_vpg.py:123-127).
"""

import ctypes as C
import math
import operator

import torch

from dppo_b200 import _lib


_VERSION = operator.attrgetter("_version")
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda dev: torch.cuda.current_stream(dev).cuda_stream)
_CUR_DEVICE = torch.cuda.current_device


def _mlp_param_list(net):
    """Parameters of a DiffusionMLP in the order dppo_pack_mlp expects (include/dppo_b200.h)."""
    ps = [net.time_embedding[1].weight, net.time_embedding[1].bias, net.time_embedding[3].weight, net.time_embedding[3].bias]
    if hasattr(net, "cond_mlp"):
        for stage in net.cond_mlp.moduleList:
            ps += [stage.linear_1.weight, stage.linear_1.bias]
    layers = net.mlp_mean.layers
    ps += [layers[0].weight, layers[0].bias]
    nb = net.mlp_mean.n_blocks
    for b in range(1, nb + 1):
        blk = layers[b]
        ps += [blk.l1.weight, blk.l1.bias, blk.l2.weight, blk.l2.bias]
        if hasattr(blk, "norm1"):
            ps += [blk.norm1.weight, blk.norm1.bias, blk.norm2.weight, blk.norm2.bias]
    ps += [layers[nb + 1].weight, layers[nb + 1].bias]
    return ps


def mlp_desc_of(net):
    """dppo_mlp_desc of a dppo_b200.model.diffusion.mlp_diffusion.DiffusionMLP; raises for geometries outside the kernel."""
    if type(net).__name__ != "DiffusionMLP":
        raise NotImplementedError(f"the sm_100a chain kernel covers DiffusionMLP; got {type(net).__name__}")
    if not net.residual_style:
        raise NotImplementedError("chain kernel: only residual_style=True trunks (all DPPO fine-tuning configs)")
    if net.out_activation_type != "Identity":
        raise NotImplementedError("chain kernel: out_activation_type must be Identity")
    dims = net.mlp_dims
    if len(set(dims)) != 1 or (len(dims) - 1) % 2:
        raise NotImplementedError(f"chain kernel: mlp_dims {dims} is not a residual trunk of equal widths")
    act = {"ReLU": _lib.ACT_RELU, "Mish": _lib.ACT_MISH}.get(net.activation_type)
    if act is None:
        raise NotImplementedError(f"chain kernel: activation {net.activation_type!r} (ReLU and Mish are built)")
    d = _lib.MlpDesc()
    d.cond_dim, d.action_dim, d.horizon_steps, d.time_dim = net.cond_dim, net.action_dim, net.horizon_steps, net.time_dim
    d.hidden_dim, d.n_blocks, d.activation, d.use_layernorm = dims[0], (len(dims) - 1) // 2, act, int(net.use_layernorm)
    if net.cond_mlp_dims is not None:
        if len(net.cond_mlp_dims) != 2:
            raise NotImplementedError("chain kernel: cond_mlp must have two layers")
        d.cond_hidden, d.cond_out = net.cond_mlp_dims
    return d


def _unet_param_list(net):
    """Parameters of a Unet1D in the order dppo_pack_unet expects (include/dppo_b200.h): execution order."""
    ps = [net.time_mlp[1].weight, net.time_mlp[1].bias, net.time_mlp[3].weight, net.time_mlp[3].bias]

    def res(blk):
        out = []
        for cb in blk.blocks:
            out += [cb.block[0].weight, cb.block[0].bias, cb.block[2].weight, cb.block[2].bias]
        for m in blk.cond_encoder:
            if isinstance(m, torch.nn.Linear):
                out += [m.weight, m.bias]
        if isinstance(blk.residual_conv, torch.nn.Conv1d):
            out += [blk.residual_conv.weight, blk.residual_conv.bias]
        return out

    for r1, r2, down in net.down_modules:
        ps += res(r1) + res(r2)
        if hasattr(down, "conv"):
            ps += [down.conv.weight, down.conv.bias]
    for m in net.mid_modules:
        ps += res(m)
    for r1, r2, up in net.up_modules:
        ps += res(r1) + res(r2)
        if hasattr(up, "conv"):
            ps += [up.conv.weight, up.conv.bias]
    fc = net.final_conv
    ps += [fc[0].block[0].weight, fc[0].block[0].bias, fc[0].block[2].weight, fc[0].block[2].bias, fc[1].weight, fc[1].bias]
    return ps


def unet_desc_of(net, horizon_steps):
    """dppo_unet_desc of a dppo_b200.model.diffusion.unet.Unet1D; raises for variants outside the kernel."""
    if hasattr(net, "cond_mlp"):
        raise NotImplementedError("unet chain kernel: cond_mlp_dims is not supported (no fine-tuning YAML sets it)")
    blk = net.mid_modules[0]
    conv0 = blk.blocks[0].block[0]
    norm0 = blk.blocks[0].block[2]
    if not isinstance(norm0, torch.nn.GroupNorm):
        raise NotImplementedError("unet chain kernel: n_groups must be set (GroupNorm)")
    act_mod = blk.blocks[0].block[4]
    act = {"ReLU": _lib.ACT_RELU, "Mish": _lib.ACT_MISH}.get(type(act_mod).__name__)
    if act is None:
        raise NotImplementedError(f"unet chain kernel: activation {type(act_mod).__name__}")
    d = _lib.UnetDesc()
    first = net.down_modules[0][0]
    d.action_dim = first.blocks[0].block[0].in_channels
    d.horizon_steps = horizon_steps
    d.time_dim = net.time_dim
    n_lin = [m for m in blk.cond_encoder if isinstance(m, torch.nn.Linear)]
    d.larger_encoder = int(len(n_lin) == 3)
    d.cond_dim = n_lin[0].in_features - net.time_dim
    widths = [lvl[0].out_channels for lvl in net.down_modules]
    if len(widths) > _lib.UNET_MAX_LEVELS:
        raise NotImplementedError(f"unet chain kernel: at most {_lib.UNET_MAX_LEVELS} levels")
    d.dim = widths[0]
    d.n_levels = len(widths)
    for i, w in enumerate(widths):
        if w % widths[0]:
            raise NotImplementedError("unet chain kernel: level widths must be multiples of dim")
        d.dim_mults[i] = w // widths[0]
    d.kernel_size = conv0.kernel_size[0]
    d.n_groups = norm0.num_groups
    d.activation = act
    d.cond_predict_scale = int(blk.cond_predict_scale)
    d.groupnorm_eps = float(norm0.eps)
    return d


class ChainEngine:
    """One context per (model, device, precision)."""

    def __init__(self, model, precision="split3"):
        self.lib = _lib.load()
        dev = torch.device(model.device)
        if dev.type != "cuda":
            raise RuntimeError("dppo_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self.device = dev
        self.precision = precision
        self.ft = int(model.ft_denoising_steps)
        self.D = model.horizon_steps * model.action_dim
        self.S = int(model.ddim_steps) if model.use_ddim else int(model.denoising_steps)
        self.is_unet = type(model.actor).__name__ == "Unet1D"
        desc = unet_desc_of(model.actor, model.horizon_steps) if self.is_unet else mlp_desc_of(model.actor)
        sd = _lib.SchedDesc()
        sd.denoising_steps, sd.ft_denoising_steps = model.denoising_steps, self.ft
        sd.use_ddim, sd.ddim_steps = int(model.use_ddim), int(model.ddim_steps or 0)
        eta = 1.0
        if model.use_ddim:
            if getattr(model, "learn_eta", False):
                raise NotImplementedError("learned eta is outside the hot path (no YAML enables it)")
            eta = float(model.eta.value().item()) if hasattr(model, "eta") else 1.0
        sd.eta = eta
        opt = lambda v: -1.0 if v is None else float(v)  # noqa: E731
        sd.denoised_clip_value, sd.randn_clip_value = opt(model.denoised_clip_value), float(model.randn_clip_value)
        sd.final_action_clip_value, sd.eps_clip_value = opt(model.final_action_clip_value), opt(model.eps_clip_value)
        sd.min_logprob_denoising_std = float(model.min_logprob_denoising_std)
        keep = []

        def host_f32(t):
            arr = (C.c_float * t.numel())(*t.detach().float().cpu().tolist())
            keep.append(arr)
            return C.cast(arr, C.POINTER(C.c_float))

        for name in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "ddpm_mu_coef1", "ddpm_mu_coef2",
                     "ddpm_logvar_clipped"):
            setattr(sd, name, host_f32(getattr(model, name)))
        if model.use_ddim:
            ts = model.ddim_t.cpu().tolist()
            arr = (C.c_int32 * len(ts))(*ts)
            keep.append(arr)
            sd.ddim_t = C.cast(arr, C.POINTER(C.c_int32))
            for name in ("ddim_alphas", "ddim_alphas_prev", "ddim_sqrt_one_minus_alphas"):
                setattr(sd, name, host_f32(getattr(model, name)))
        self.ctx = C.c_void_p()
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        create = self.lib.dppo_ctx_create_unet if self.is_unet else self.lib.dppo_ctx_create
        _lib.check(create(C.byref(self.ctx), C.byref(desc), C.byref(sd), _lib.PRECISIONS[precision], idx), "dppo_ctx_create")
        self._packed = {0: None, 1: None}
        self._plists = {}
        self._current = None  # (actor, actor_ft, their parameters, signature, split) as of the last full synchronisation
        self._host_io = {}  # batch size -> [next slot, ring of (trajectories, chains, traj ptr, chain ptr)] (sample_host)
        self._sample_host_fn = self.lib.dppo_sample_chain_host
        self.Ta, self.Da = int(model.horizon_steps), int(model.action_dim)
        self.cond_numel = int(desc.cond_dim)
        self._ws = torch.zeros(32, dtype=torch.float64, device=dev)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.dppo_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def set_launch_shape(self, tile_envs=0, cluster=0):
        """Tuning / test hook: force the chain kernel's tile size and cluster size (0 = cost model)."""
        self.lib.dppo_debug_set_shape.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _lib.check(self.lib.dppo_debug_set_shape(self.ctx, int(tile_envs), int(cluster)), "dppo_debug_set_shape")

    # ------------------------------------------------------------------ weights
    def sync_weights(self, which, net):
        # hot path of every rollout decision: the parameter list is cached per module object and the signature is the
        # tensors' version counters (every in-place update bumps them; FlatAdamW bumps them explicitly) plus the first
        # tensor's address (catches a wholesale re-homing of the storage)
        cached = self._plists.get(which)
        if cached is None or cached[0] is not net:
            cached = (net, _unet_param_list(net) if self.is_unet else _mlp_param_list(net))
            self._plists[which] = cached
            self._packed[which] = None
        ps = cached[1]
        sig = (ps[0].data_ptr(), ps[-1].data_ptr(), *map(_VERSION, ps))
        if self._packed[which] == sig:
            return False
        for p in ps:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise RuntimeError("dppo_pack_mlp needs contiguous fp32 CUDA parameters")
        arr = (C.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
        pack = self.lib.dppo_pack_unet if self.is_unet else self.lib.dppo_pack_mlp
        _lib.check(pack(self.ctx, which, arr, len(ps), _lib.stream_ptr()), "dppo_pack_unet" if self.is_unet else "dppo_pack_mlp")
        self._packed[which] = sig
        return True

    def weights_current(self, actor, actor_ft):
        """Per-decision fast path of the cache check: True when neither network (module identity, storage address of the
        first parameter, every parameter's version counter) changed since `mark_current`."""
        c = self._current
        if c is None or c[0] is not actor or c[1] is not actor_ft:
            return False
        ps = c[2]
        return c[3] == (ps[0].data_ptr(), ps[c[4]].data_ptr(), *map(_VERSION, ps))

    def mark_current(self, actor, actor_ft):
        """Record the state both packed copies were just synchronised with (after sync_weights(0, ...) and (1, ...))."""
        pa, pf = self._plists[0][1], self._plists[1][1]
        ps = pa + pf
        self._current = (actor, actor_ft, ps, (ps[0].data_ptr(), ps[len(pa)].data_ptr(), *map(_VERSION, ps)), len(pa))

    # ------------------------------------------------------------------ rollout
    def sample(self, state, noise=None, seed=0, offset=0, env_offset=0, deterministic=False, use_base_policy=False,
               min_sampling_std=0.1, return_chain=True, out_traj=None, out_chain=None):
        """`state` may be a pinned host tensor (read once by the kernel prologue); `out_traj` (E, D) / `out_chain`
        (E, ft+1, D) may be preallocated CUDA or PINNED HOST tensors the kernel writes into directly (zero-copy)."""
        E = state.shape[0]
        if state.dtype != torch.float32 or not state.is_contiguous():
            state = state.contiguous().float()
        state = state.view(E, math.prod(state.shape[1:]))
        if not (state.is_cuda or state.is_pinned()):
            state = state.to(self.device)
        traj = out_traj if out_traj is not None else torch.empty((E, self.D), dtype=torch.float32, device=self.device)
        chain = out_chain if out_chain is not None else (
            torch.empty((E, self.ft + 1, self.D), dtype=torch.float32, device=self.device) if return_chain else None)
        for name, t, shape in (("out_traj", out_traj, E * self.D), ("out_chain", out_chain, E * (self.ft + 1) * self.D)):
            if t is not None and (t.dtype != torch.float32 or t.numel() != shape):
                raise RuntimeError(f"{name}: expected {shape} float32 elements")
        if E == 0:  # nothing to launch (empty tensors have no storage to point at)
            return traj, chain
        if noise is not None:
            noise = noise.reshape(self.S + 1, E, self.D).contiguous().float()
        _lib.check(
            self.lib.dppo_sample_chain(self.ctx, _lib.ptr_dev_or_pinned(state), E, _lib.ptr(noise), seed, offset, env_offset,
                                       int(deterministic), int(use_base_policy), float(min_sampling_std),
                                       _lib.ptr_dev_or_pinned(traj), _lib.ptr_dev_or_pinned(chain), _lib.stream_ptr()),
            "dppo_sample_chain")
        return traj, chain

    HOST_RING = 4            # result buffers per batch size of sample_host
    PIN_CHECK_BYTES = 65536  # below this a staging memcpy inside the library is cheaper than asking whether `state` is pinned

    def sample_host(self, state, seed=0, offset=0, env_offset=0, deterministic=False, use_base_policy=False,
                    min_sampling_std=0.1, return_chain=True, chain_out=None):
        """Host observations in, host results out, one call (dppo_sample_chain_host): `state` is a CPU tensor (E, ...) -
        pageable (staged by the library) or pinned (read in place); the kernel stores trajectories (E, Ta, Da) and chains
        (E, ft+1, Ta, Da) straight into page-locked memory and the call returns when they are there.  The returned tensors
        are views into a ring of HOST_RING page-locked buffers per batch size: they stay valid for the next
        HOST_RING - 1 calls (the rollout loop consumes them within the step, reference train_ppo_diffusion_agent.py:112-122).
        `chain_out`: optional contiguous float32 tensor (E, ft+1, Ta, Da) in device-accessible memory - a slice of a
        device-resident rollout buffer or a pinned host tensor - the chains are stored there instead of in the ring."""
        E = state.shape[0]
        io = self._host_io.get(E)
        if io is None:
            io = self._host_io[E] = self._host_ring(E)
        if state.dtype is not torch.float32 or not state.is_contiguous():
            state = state.contiguous().float()
        n = state.numel()
        if n != E * self.cond_numel:
            raise RuntimeError(f"state: expected {E} x {self.cond_numel} elements, got {tuple(state.shape)}")
        slot = io[0]
        io[0] = slot + 1 if slot + 1 < self.HOST_RING else 0
        traj, chain, traj_p, chain_p = io[1][slot]
        if not return_chain:
            chain = chain_p = None
        elif chain_out is not None:
            if (chain_out.dtype is not torch.float32 or chain_out.numel() != chain.numel() or not chain_out.is_contiguous()
                    or not (chain_out.is_cuda or chain_out.is_pinned())):
                raise RuntimeError(f"chain_out: expected a contiguous float32 CUDA / pinned tensor of {tuple(chain.shape)}")
            chain, chain_p = chain_out.view(chain.shape), chain_out.data_ptr()
        if E == 0:
            return traj, chain
        flags = _lib.HOST_OUT_PINNED
        if n * 4 > self.PIN_CHECK_BYTES and state.is_pinned():
            flags |= _lib.HOST_STATE_PINNED
        rc = self._sample_host_fn(self.ctx, state.data_ptr(), E, seed, offset, env_offset, deterministic, use_base_policy,
                                  min_sampling_std, traj_p, chain_p, flags, _RAW_STREAM(_CUR_DEVICE()))
        if rc:
            _lib.check(rc, "dppo_sample_chain_host")
        return traj, chain

    def _host_ring(self, E):
        ring = []
        for _ in range(self.HOST_RING):
            t = torch.empty((E, self.Ta, self.Da), dtype=torch.float32).pin_memory()
            c = torch.empty((E, self.ft + 1, self.Ta, self.Da), dtype=torch.float32).pin_memory()
            ring.append((t, c, t.data_ptr(), c.data_ptr()))
        return [0, ring]

    def nonfinite(self, reset=True):
        """True when a sampling launch since the last reset produced a NaN / Inf action element (synchronises the stream)."""
        flag = C.c_int(0)
        _lib.check(self.lib.dppo_sample_nonfinite(self.ctx, C.byref(flag), int(reset), _lib.stream_ptr()), "dppo_sample_nonfinite")
        return bool(flag.value)

    def chain_logprobs(self, state, chains, use_base_policy=False):
        B = chains.shape[0]
        state = state.reshape(B, -1).contiguous().float()
        chains = chains.reshape(B, self.ft + 1, self.D).contiguous().float()
        logp = torch.empty((B * self.ft, self.D), dtype=torch.float32, device=chains.device)
        if B == 0 or self.ft == 0:
            return logp
        _lib.check(self.lib.dppo_chain_logprobs(self.ctx, _lib.ptr(state), _lib.ptr(chains), B, int(use_base_policy),
                                                _lib.ptr(logp), _lib.stream_ptr()), "dppo_chain_logprobs")
        return logp

    # ------------------------------------------------------------------ update (elementwise halves)
    def logprob_rows(self, eps, x_prev, x_next, denoising_inds, want_grad_factor=False):
        B = eps.shape[0]
        eps, x_prev, x_next = (t.reshape(B, self.D).contiguous().float() for t in (eps, x_prev, x_next))
        dinds = denoising_inds.contiguous().to(torch.int64)
        logp = torch.empty_like(eps)
        fac = torch.empty_like(eps) if want_grad_factor else None
        _lib.check(self.lib.dppo_logprob_rows(self.ctx, _lib.ptr(eps), _lib.ptr(x_prev), _lib.ptr(x_next), _lib.ptr(dinds),
                                              B, _lib.ptr(logp), _lib.ptr(fac), _lib.stream_ptr()), "dppo_logprob_rows")
        return logp, fac

    def make_hp(self, model, reward_horizon, adv_lo=-math.inf, adv_hi=math.inf):
        hp = _lib.LossHp()
        hp.ft_denoising_steps, hp.horizon_steps, hp.action_dim = self.ft, model.horizon_steps, model.action_dim
        hp.reward_horizon, hp.norm_adv = int(reward_horizon), int(model.norm_adv)
        hp.gamma_denoising = float(model.gamma_denoising)
        hp.clip_ploss_coef, hp.clip_ploss_coef_base = float(model.clip_ploss_coef), float(model.clip_ploss_coef_base)
        hp.clip_ploss_coef_rate = float(model.clip_ploss_coef_rate)
        hp.clip_vloss_coef = -1.0 if model.clip_vloss_coef is None else float(model.clip_vloss_coef)
        hp.adv_clip_lo, hp.adv_clip_hi = adv_lo, adv_hi
        return hp

    def loss_rows(self, hp, x_prev, x_next, old_lp, returns, old_values, advantages, denoising_inds, eps, vpred):
        """dppo_ppo_loss_rows: inputs already gathered per row. Returns (grad_eps, grad_v, scalars[8])."""
        B = eps.shape[0]
        f = lambda t: t.reshape(B, -1).contiguous().float()  # noqa: E731
        x_prev, x_next, old_lp, eps = f(x_prev), f(x_next), f(old_lp), f(eps)
        g = lambda t: t.reshape(B).contiguous().float()  # noqa: E731
        returns, old_values, advantages, vpred = g(returns), g(old_values), g(advantages), g(vpred)
        dinds = denoising_inds.contiguous().to(torch.int64)
        grad_eps, grad_v = torch.empty_like(eps), torch.empty_like(vpred)
        scalars = torch.empty(8, dtype=torch.float32, device=eps.device)
        _lib.check(
            self.lib.dppo_ppo_loss_rows(self.ctx, _lib.ptr(x_prev), _lib.ptr(x_next), _lib.ptr(old_lp), _lib.ptr(returns),
                                        _lib.ptr(old_values), _lib.ptr(advantages), _lib.ptr(dinds), _lib.ptr(eps),
                                        _lib.ptr(vpred), B, C.byref(hp), _lib.ptr(grad_eps), _lib.ptr(grad_v),
                                        _lib.ptr(scalars), _lib.ptr(self._ws), _lib.stream_ptr()),
            "dppo_ppo_loss_rows")
        return grad_eps, grad_v, scalars

    def loss_gathered(self, hp, chains_k, logprobs_k, returns_k, values_k, advantages_k, inds_all, row_begin, eps, vpred):
        """dppo_ppo_loss_fwd_bwd: gathers fused into the kernel; this rank owns rows [row_begin, row_begin+len(eps))."""
        n = eps.shape[0]
        eps = eps.reshape(n, self.D).contiguous().float()
        vpred = vpred.reshape(n).contiguous().float()
        grad_eps, grad_v = torch.empty_like(eps), torch.empty_like(vpred)
        scalars = torch.empty(8, dtype=torch.float32, device=eps.device)
        for t in (chains_k, logprobs_k, returns_k, values_k, advantages_k):
            if t.dtype != torch.float32:
                raise RuntimeError("rollout buffers must be fp32")
        _lib.check(
            self.lib.dppo_ppo_loss_fwd_bwd(self.ctx, _lib.ptr(chains_k), _lib.ptr(logprobs_k), _lib.ptr(returns_k),
                                           _lib.ptr(values_k), _lib.ptr(advantages_k), _lib.ptr(inds_all), int(row_begin),
                                           _lib.ptr(eps), _lib.ptr(vpred), n, int(inds_all.numel()), C.byref(hp),
                                           _lib.ptr(grad_eps), _lib.ptr(grad_v), _lib.ptr(scalars), _lib.ptr(self._ws),
                                           _lib.stream_ptr()),
            "dppo_ppo_loss_fwd_bwd")
        return grad_eps, grad_v, scalars


def gae(reward, terminated, values, next_value, gamma, gae_lambda, reward_scale_const=1.0):
    """dppo_gae_f64 on float64 CUDA tensors of shape (n_steps, n_envs); returns (advantages, returns)."""
    lib = _lib.load()
    n_steps, E = reward.shape
    cast = lambda t: t.contiguous().to(torch.float64)  # noqa: E731
    reward, terminated, values, next_value = cast(reward), cast(terminated), cast(values), cast(next_value).reshape(-1)
    adv, ret = torch.empty_like(reward), torch.empty_like(reward)
    _lib.check(lib.dppo_gae_f64(_lib.ptr(reward), _lib.ptr(terminated), _lib.ptr(values), _lib.ptr(next_value), n_steps, E,
                                float(gamma), float(gae_lambda), float(reward_scale_const), _lib.ptr(adv), _lib.ptr(ret),
                                _lib.stream_ptr()), "dppo_gae_f64")
    return adv, ret
