"""
ctypes binding of libdppo_b200.so (C ABI in include/dppo_b200.h).

There is no CPU fallback: every call raises if the library is missing or a kernel launch fails.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdppo_b200.so")

ACT_RELU, ACT_MISH = 0, 1
NET_ACTOR, NET_ACTOR_FT = 0, 1
PRECISION_SPLIT3, PRECISION_BF16 = 0, 1
PRECISIONS = {"split3": PRECISION_SPLIT3, "bf16": PRECISION_BF16}

# every symbol include/dppo_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "dppo_last_error", "dppo_version", "dppo_ctx_create", "dppo_ctx_destroy", "dppo_pack_mlp", "dppo_sample_chain",
    "dppo_chain_logprobs", "dppo_logprob_rows", "dppo_ppo_loss_fwd_bwd", "dppo_ppo_loss_rows", "dppo_gae_f64",
    "dppo_selftest_umma",
]


class MlpDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "cond_dim", "action_dim", "horizon_steps", "time_dim", "hidden_dim", "n_blocks", "activation",
        "use_layernorm", "cond_hidden", "cond_out")]


class SchedDesc(C.Structure):
    _fields_ = [
        ("denoising_steps", C.c_int32), ("ft_denoising_steps", C.c_int32), ("use_ddim", C.c_int32),
        ("ddim_steps", C.c_int32), ("eta", C.c_float), ("denoised_clip_value", C.c_float),
        ("randn_clip_value", C.c_float), ("final_action_clip_value", C.c_float), ("eps_clip_value", C.c_float),
        ("min_logprob_denoising_std", C.c_float),
        ("sqrt_recip_alphas_cumprod", C.POINTER(C.c_float)), ("sqrt_recipm1_alphas_cumprod", C.POINTER(C.c_float)),
        ("ddpm_mu_coef1", C.POINTER(C.c_float)), ("ddpm_mu_coef2", C.POINTER(C.c_float)),
        ("ddpm_logvar_clipped", C.POINTER(C.c_float)), ("ddim_t", C.POINTER(C.c_int32)),
        ("ddim_alphas", C.POINTER(C.c_float)), ("ddim_alphas_prev", C.POINTER(C.c_float)),
        ("ddim_sqrt_one_minus_alphas", C.POINTER(C.c_float)),
    ]


class LossHp(C.Structure):
    _fields_ = [
        ("ft_denoising_steps", C.c_int32), ("horizon_steps", C.c_int32), ("action_dim", C.c_int32),
        ("reward_horizon", C.c_int32), ("norm_adv", C.c_int32), ("gamma_denoising", C.c_float),
        ("clip_ploss_coef", C.c_float), ("clip_ploss_coef_base", C.c_float), ("clip_ploss_coef_rate", C.c_float),
        ("clip_vloss_coef", C.c_float), ("adv_clip_lo", C.c_float), ("adv_clip_hi", C.c_float),
    ]


_lib = None


def load(build_if_missing=True):
    """Return the loaded CDLL (building it in-tree first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} not built; run `python -m dppo_b200.build`")
        from dppo_b200.build import build

        build()
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double
    lib.dppo_last_error.restype = C.c_char_p
    lib.dppo_last_error.argtypes = []
    lib.dppo_version.restype = i32
    lib.dppo_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(MlpDesc), C.POINTER(SchedDesc), i32, i32]
    lib.dppo_ctx_destroy.argtypes = [vp]
    lib.dppo_pack_mlp.argtypes = [vp, i32, C.POINTER(vp), i32, vp]
    lib.dppo_sample_chain.argtypes = [vp, vp, i32, vp, u64, u64, i64, i32, i32, f32, vp, vp, vp]
    lib.dppo_chain_logprobs.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.dppo_logprob_rows.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.dppo_ppo_loss_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, C.POINTER(LossHp), vp, vp,
                                          vp, vp, vp]
    lib.dppo_ppo_loss_rows.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(LossHp), vp, vp, vp, vp, vp]
    lib.dppo_gae_f64.argtypes = [vp, vp, vp, vp, i32, i32, f64, f64, f64, vp, vp, vp]
    lib.dppo_selftest_umma.argtypes = [vp, vp, vp, vp, i32, i32, u64, C.c_uint32, vp]
    for name in EXPORTS:
        if name not in ("dppo_last_error",):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dppo_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("dppo_b200 kernels need CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise RuntimeError("dppo_b200 kernels need contiguous tensors")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
