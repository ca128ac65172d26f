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
HOST_STATE_PINNED, HOST_OUT_PINNED = 1, 2  # dppo_sample_chain_host flags

# every symbol include/dppo_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "dppo_last_error", "dppo_version", "dppo_ctx_create", "dppo_ctx_destroy", "dppo_pack_mlp", "dppo_ctx_create_unet",
    "dppo_pack_unet", "dppo_unet_param_count", "dppo_sample_chain", "dppo_sample_chain_host", "dppo_sample_nonfinite",
    "dppo_chain_logprobs", "dppo_logprob_rows", "dppo_ppo_loss_fwd_bwd", "dppo_ppo_loss_rows", "dppo_gae_f64", "dppo_split3_pack", "dppo_reward_scale_f64", "dppo_adamw_flat",
    "dppo_update_create", "dppo_update_destroy", "dppo_update_bind", "dppo_update_forward", "dppo_update_backward",
    "dppo_update_minibatch", "dppo_update_values", "dppo_update_set_actor_event", "dppo_update_buffers", "dppo_memset_zero", "dppo_adamw_flat_dev", "dppo_kl_check",
]


class MlpDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "cond_dim", "action_dim", "horizon_steps", "time_dim", "hidden_dim", "n_blocks", "activation",
        "use_layernorm", "cond_hidden", "cond_out")]


UNET_MAX_LEVELS = 4


class UnetDesc(C.Structure):
    _fields_ = [
        ("cond_dim", C.c_int32), ("action_dim", C.c_int32), ("horizon_steps", C.c_int32), ("time_dim", C.c_int32),
        ("dim", C.c_int32), ("n_levels", C.c_int32), ("dim_mults", C.c_int32 * UNET_MAX_LEVELS),
        ("kernel_size", C.c_int32), ("n_groups", C.c_int32), ("activation", C.c_int32),
        ("cond_predict_scale", C.c_int32), ("larger_encoder", C.c_int32), ("groupnorm_eps", C.c_float),
    ]


class UGemm(C.Structure):
    """csrc/unet_plan.h UGemm (host-only inspection of the lowered Unet program)."""
    _fields_ = [("tile_off", C.c_uint32), ("mt", C.c_uint16), ("kc", C.c_uint16), ("src_chunk", C.c_uint16 * 2),
                ("src_n", C.c_uint16 * 2), ("acc_tile", C.c_uint16), ("pad", C.c_uint16)]


class ULayer(C.Structure):
    """csrc/unet_plan.h ULayer."""
    _fields_ = [("n_gemm", C.c_int32), ("g", UGemm * 2)] + [(n, C.c_int32) for n in (
        "kind", "acc_tile", "mt", "nf", "bias_off", "bias_tstride", "gn_size", "gamma_off", "beta_off")] + [
        ("gn_eps", C.c_float)] + [(n, C.c_int32) for n in (
            "act", "film", "film_c", "film_tshift", "res", "res_acc_tile", "res_bias_off", "res_chunk", "dst_chunk", "track",
            "wait_chunk", "wait_tiles")]


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


class ResMlpDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "hidden_dim", "n_blocks", "out_dim", "activation", "use_layernorm")]


class UpdateBatch(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "chains", "x_next", "old_logprobs", "returns", "old_values", "advantages",
                                          "inds_all", "denoising_inds")] + [
        (n, C.c_int32) for n in ("row_begin", "n_rows", "global_rows")]


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
    lib.dppo_ctx_create_unet.argtypes = [C.POINTER(vp), C.POINTER(UnetDesc), C.POINTER(SchedDesc), i32, i32]
    lib.dppo_pack_unet.argtypes = [vp, i32, C.POINTER(vp), i32, vp]
    lib.dppo_unet_param_count.argtypes = [C.POINTER(UnetDesc)]
    # host-only inspection of the lowered Unet program (tests; not part of the public header)
    lib.dppo_unet_plan_create.argtypes = [C.POINTER(UnetDesc), i32, i32, C.POINTER(vp)]
    lib.dppo_unet_plan_destroy.argtypes = [vp]
    lib.dppo_unet_plan_info.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.dppo_unet_plan_layers.argtypes = [vp, vp]
    lib.dppo_unet_plan_dense.argtypes = [vp, i32, C.POINTER(vp), vp]
    lib.dppo_unet_plan_side.argtypes = [vp, C.POINTER(vp), vp]
    for name in ("dppo_unet_plan_create", "dppo_unet_plan_destroy", "dppo_unet_plan_info", "dppo_unet_plan_layers",
                 "dppo_unet_plan_dense", "dppo_unet_plan_side"):
        getattr(lib, name).restype = i32
    lib.dppo_sample_chain.argtypes = [vp, vp, i32, vp, u64, u64, i64, i32, i32, f32, vp, vp, vp]
    lib.dppo_sample_chain_host.argtypes = [vp, vp, i32, u64, u64, i64, i32, i32, f32, vp, vp, i32, vp]
    lib.dppo_sample_nonfinite.argtypes = [vp, C.POINTER(C.c_int), i32, vp]
    lib.dppo_chain_logprobs.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.dppo_logprob_rows.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.dppo_ppo_loss_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, C.POINTER(LossHp), vp, vp,
                                          vp, vp, vp]
    lib.dppo_ppo_loss_rows.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(LossHp), vp, vp, vp, vp, vp]
    lib.dppo_gae_f64.argtypes = [vp, vp, vp, vp, i32, i32, f64, f64, f64, vp, vp, vp]
    lib.dppo_split3_pack.argtypes = [vp, i64, i32, i64, i32, vp, vp, i32, vp]
    lib.dppo_reward_scale_f64.argtypes = [vp, vp, i32, i32, C.c_longlong, f64, f64, f64, vp, vp, vp, vp, vp, i32, vp]
    lib.dppo_adamw_flat.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp, vp]
    lib.dppo_update_create.argtypes = [C.POINTER(vp), vp, C.POINTER(ResMlpDesc), i32]
    lib.dppo_update_destroy.argtypes = [vp]
    lib.dppo_update_bind.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), i32, C.POINTER(vp), C.POINTER(vp), i32]
    lib.dppo_update_forward.argtypes = [vp, C.POINTER(UpdateBatch), vp, vp, vp]
    lib.dppo_update_backward.argtypes = [vp, vp, vp, vp, vp, f32, i32, i32, vp]
    lib.dppo_update_minibatch.argtypes = [vp, C.POINTER(UpdateBatch), C.POINTER(LossHp), f32, i32, vp, vp, vp]
    lib.dppo_update_values.argtypes = [vp, vp, C.c_int, vp, vp]
    lib.dppo_update_set_actor_event.argtypes = [vp, vp]
    lib.dppo_update_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.dppo_memset_zero.argtypes = [vp, C.c_size_t, vp]
    lib.dppo_adamw_flat_dev.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp, vp, f32, vp, vp]
    lib.dppo_kl_check.argtypes = [vp, f32, i32, vp, vp, i32, vp]
    # bring-up / unit-test hooks (not part of the public header)
    lib.dppo_debug_linear.argtypes = [vp, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, i32, vp, vp]
    lib.dppo_debug_wgrad.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.dppo_debug_set_mn_desc.argtypes = [C.c_uint, C.c_uint]
    for name in ("dppo_debug_linear", "dppo_debug_wgrad", "dppo_debug_set_mn_desc"):
        getattr(lib, name).restype = i32
    for name in EXPORTS:
        if name not in ("dppo_last_error",):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


_test_lib = None


def load_test():
    """The test-only library (tcgen05 descriptor self-test + micro-benchmarks, csrc/umma_selftest.cu, csrc/microbench.cu)."""
    global _test_lib
    if _test_lib is None:
        load()
        path = os.path.join(_HERE, "lib", "libdppo_b200_test.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built; run `python -m dppo_b200.build`")
        lib = C.CDLL(path)
        vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
        lib.dppo_selftest_umma.argtypes = [vp, vp, vp, vp, i32, i32, u64, C.c_uint32, vp]
        lib.dppo_selftest_umma.restype = i32
        _test_lib = lib
    return _test_lib


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


def ptr_dev_or_pinned(t):
    """Pointer of a contiguous CUDA tensor or of a PINNED host tensor (page-locked memory is mapped into the device's
    address space under unified addressing: a kernel can read observations from it / write results into it directly,
    which replaces a separate copy launch by PCIe transactions overlapped with the kernel)."""
    if t is None:
        return None
    if not (t.is_cuda or t.is_pinned()):
        raise RuntimeError("dppo_b200 kernels need CUDA tensors or pinned host tensors; there is no CPU path")
    if not t.is_contiguous():
        raise RuntimeError("dppo_b200 kernels need contiguous tensors")
    return C.c_void_p(t.data_ptr())


def raw_stream():
    """cudaStream_t of torch's current stream on the current device as a plain int (argtypes convert it)."""
    import torch

    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return raw(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (raw getter: ~0.3 us instead of ~5 us)."""
    import torch

    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
