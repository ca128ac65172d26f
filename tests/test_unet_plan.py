"""
CPU check of the Unet1D lowering (csrc/unet_plan.cu): the program of dense layers the sm_100a kernel executes is
interpreted here in numpy (float64), with the dense matrices and the side table produced by the library's own host-side
lowering code, and compared with the oracle's Unet1D forward.  No GPU, no CUDA call.
"""

import ctypes as C

import numpy as np
import pytest
import torch

from dppo_b200 import _lib
from dppo_b200.engine import _unet_param_list, unet_desc_of
from dppo_b200.model.diffusion.unet import Unet1D
from oracle import dppo_oracle as O

EPI_OPERAND, EPI_FILM, EPI_EPS = 0, 1, 2
RES_NONE, RES_ACC, RES_SLOT = 0, 1, 2


def _mish(x):
    return x * np.tanh(np.log1p(np.exp(x)))


class Plan:
    def __init__(self, net, horizon, K):
        self.lib = _lib.load()
        self.desc = unet_desc_of(net, horizon)
        self.h = C.c_void_p()
        _lib.check(self.lib.dppo_unet_plan_create(C.byref(self.desc), K, _lib.PRECISION_SPLIT3, C.byref(self.h)),
                   "dppo_unet_plan_create")
        info = (C.c_int64 * 16)()
        self.lib.dppo_unet_plan_info(self.h, info)
        (self.n_layers, self.n_jobs, self.n_side, self.total_chunks, self.chunk_x, self.KX, self.chunk_state, self.KS,
         self.chunk_state_act, self.film_dim, self.MTmax, self.n_params, self.n_tiles, sz, self.macs, self.nsplit) = list(info)
        assert sz == C.sizeof(_lib.ULayer), (sz, C.sizeof(_lib.ULayer))
        self.layers = (_lib.ULayer * self.n_layers)()
        self.lib.dppo_unet_plan_layers(self.h, C.cast(self.layers, C.c_void_p))
        ps = [p.detach().contiguous().float().numpy() for p in _unet_param_list(net)]
        assert len(ps) == self.n_params == self.lib.dppo_unet_param_count(C.byref(self.desc))
        self._keep = ps
        self.params = (C.c_void_p * len(ps))(*[p.ctypes.data for p in ps])
        self.side = np.zeros(self.n_side, dtype=np.float32)
        self.lib.dppo_unet_plan_side(self.h, self.params, self.side.ctypes.data)

    def dense(self, job, mt, kc):
        out = np.zeros((mt * 128, kc * 64), dtype=np.float32)
        _lib.check(self.lib.dppo_unet_plan_dense(self.h, job, self.params, out.ctypes.data), "dppo_unet_plan_dense")
        return out

    def close(self):
        self.lib.dppo_unet_plan_destroy(self.h)


def run_program(plan, x, t, state, act="Mish"):
    """x (B, Ta, Da), state (B, cond_dim), scalar t -> eps (B, Ta, Da) by interpreting the layer program."""
    B, Ta, Da = x.shape
    D = Ta * Da
    actf = _mish if act == "Mish" else (lambda v: np.maximum(v, 0.0))
    op = np.zeros((B, plan.total_chunks * 64), dtype=np.float64)
    # channel-major sample operand: feature d * Ta + t
    op[:, plan.chunk_x * 64: plan.chunk_x * 64 + D] = np.transpose(x, (0, 2, 1)).reshape(B, D)
    op[:, plan.chunk_state * 64: plan.chunk_state * 64 + state.shape[1]] = state
    if plan.chunk_state_act >= 0:
        op[:, plan.chunk_state_act * 64: plan.chunk_state_act * 64 + state.shape[1]] = actf(state)
    film = np.zeros((B, plan.film_dim), dtype=np.float64)
    acc = np.zeros((B, 4 * plan.MTmax * 128), dtype=np.float64)
    side = plan.side.astype(np.float64)
    job = 0
    eps = None
    for L in plan.layers:
        for gi in range(L.n_gemm):
            G = L.g[gi]
            W = plan.dense(job, G.mt, G.kc).astype(np.float64)
            job += 1
            segs = [op[:, G.src_chunk[0] * 64: (G.src_chunk[0] + G.src_n[0]) * 64]]
            if G.src_n[1]:
                segs.append(op[:, G.src_chunk[1] * 64: (G.src_chunk[1] + G.src_n[1]) * 64])
            xin = np.concatenate(segs, axis=1)
            assert xin.shape[1] == G.kc * 64
            acc[:, G.acc_tile * 128: (G.acc_tile + G.mt) * 128] = xin @ W.T
        n = L.mt * 128
        v = acc[:, L.acc_tile * 128: L.acc_tile * 128 + n].copy()
        v += side[L.bias_off + L.bias_tstride * t: L.bias_off + L.bias_tstride * t + n]
        f = np.arange(n)
        if L.kind == EPI_OPERAND:
            if L.gn_size:
                # groups are runs of gn_size consecutive VALID features (a size that is not a power of two does not divide
                # the 128-feature tiles: the padding features behind nf belong to no group)
                nv = L.nf
                assert nv % L.gn_size == 0
                g = v[:, :nv].reshape(B, nv // L.gn_size, L.gn_size)
                m = g.mean(-1, keepdims=True)
                var = ((g - m) ** 2).mean(-1, keepdims=True)
                v[:, :nv] = ((g - m) / np.sqrt(var + L.gn_eps)).reshape(B, nv)
                v = v * side[L.gamma_off: L.gamma_off + n] + side[L.beta_off: L.beta_off + n]
            if L.act:
                v = actf(v)
            if L.film:
                ch = np.minimum(f >> L.film_tshift, L.film_c - 1)
                v = film[:, ch] * v + film[:, L.film_c + ch] if L.film == 2 else v + film[:, ch]
            if L.res == RES_ACC:
                v = v + acc[:, L.res_acc_tile * 128: L.res_acc_tile * 128 + n] + side[L.res_bias_off: L.res_bias_off + n]
            v[:, f >= L.nf] = 0.0
            if L.res == RES_SLOT:
                v = v + op[:, L.res_chunk * 64: L.res_chunk * 64 + n]
            op[:, L.dst_chunk * 64: L.dst_chunk * 64 + n] = v
        elif L.kind == EPI_FILM:
            film[:, :L.nf] = v[:, :L.nf]
        else:
            eps = v[:, :D].reshape(B, Ta, Da)  # rows are time-major: flat index t * Da + d
    assert job == plan.n_jobs
    # main-path bookkeeping of the tile hand-off: consecutive main layers use different accumulator sets and every one of
    # them names the chunks its (cyclic) predecessor writes
    main = [L for L in plan.layers if L.track == 0]
    for prev, cur in zip([main[-1]] + main[:-1], main):
        assert (prev.acc_tile >= 2 * plan.MTmax) != (cur.acc_tile >= 2 * plan.MTmax)
        if prev.kind == EPI_EPS:
            assert (cur.wait_chunk, cur.wait_tiles) == (plan.chunk_x, 1)
        else:
            assert (cur.wait_chunk, cur.wait_tiles) == (prev.dst_chunk, prev.mt)
        srcs = set()
        for gi in range(cur.n_gemm):
            G = cur.g[gi]
            srcs |= set(range(G.src_chunk[0], G.src_chunk[0] + G.src_n[0])) | set(range(G.src_chunk[1], G.src_chunk[1] + G.src_n[1]))
        assert cur.wait_chunk in srcs  # a main-path layer always consumes its predecessor's output
    return eps


CASES = [
    # (Da, Ta, cond_dim, e, dim, mults, k, groups, predict_scale, smaller_encoder)
    dict(Da=7, Ta=4, cond=23, e=16, dim=64, mults=(1, 2), k=5, groups=8, cps=True, small=False),   # cfg5 square
    dict(Da=7, Ta=4, cond=23, e=16, dim=64, mults=(1, 2), k=5, groups=8, cps=False, small=True),
    dict(Da=3, Ta=8, cond=11, e=8, dim=32, mults=(1, 2, 4), k=3, groups=8, cps=True, small=False),
    dict(Da=10, Ta=4, cond=58, e=16, dim=64, mults=(1,), k=5, groups=8, cps=False, small=False),
    # robomimic can / lift (cfg/robomimic/finetune/{can,lift}/ft_ppo_diffusion_unet.yaml): dim 40 -> GroupNorm groups of
    # 5 channels x 4 positions (10 x 2 on the second level) = 20 lowered features, the segmented path of the kernel
    dict(Da=7, Ta=4, cond=23, e=16, dim=40, mults=(1, 2), k=5, groups=8, cps=True, small=False),
    dict(Da=7, Ta=4, cond=19, e=16, dim=24, mults=(1, 2), k=5, groups=8, cps=True, small=False),  # groups of 12 / 12
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"Da{c['Da']}Ta{c['Ta']}dim{c['dim']}x{len(c['mults'])}")
def test_lowered_program_matches_oracle(case):
    torch.manual_seed(3)
    net = Unet1D(action_dim=case["Da"], cond_dim=case["cond"], diffusion_step_embed_dim=case["e"], dim=case["dim"],
                 dim_mults=case["mults"], smaller_encoder=case["small"], kernel_size=case["k"], n_groups=case["groups"],
                 cond_predict_scale=case["cps"])
    with torch.no_grad():
        for p in net.parameters():  # non-trivial norms / biases
            p.add_(0.05 * torch.randn_like(p))
    K = 20
    plan = Plan(net, case["Ta"], K)
    try:
        B = 5
        x = torch.randn(B, case["Ta"], case["Da"])
        state = torch.rand(B, 1, case["cond"]) * 2 - 1
        nc = O.NetCfg(kind="unet", obs_dim=case["cond"], cond_steps=1, action_dim=case["Da"], horizon_steps=case["Ta"],
                      time_dim=case["e"], activation="Mish", unet_dim=case["dim"], unet_mults=case["mults"],
                      unet_kernel=case["k"], unet_groups=case["groups"], unet_cond_predict_scale=case["cps"],
                      unet_smaller_encoder=case["small"])
        p = {"a." + k: v.detach() for k, v in net.state_dict().items()}
        for t in (0, 7, 19):
            ref = O.unet1d(p, "a.", nc, x, torch.full((B,), t, dtype=torch.long), state).numpy()
            also = net(x, torch.tensor(t), {"state": state}).detach().numpy()
            np.testing.assert_allclose(also, ref, rtol=1e-5, atol=1e-6)
            got = run_program(plan, x.numpy().astype(np.float64), t, state.reshape(B, -1).numpy().astype(np.float64))
            err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
            assert err < 2e-5, (t, err)
        # algorithmic size of the lowered net: never more than the conv-as-written MAC count
        assert plan.macs > 0 and plan.n_tiles * 16384 // (2 * plan.nsplit) >= plan.macs
    finally:
        plan.close()


def test_plan_rejects_unsupported_shapes():
    lib = _lib.load()
    d = _lib.UnetDesc()
    d.cond_dim, d.action_dim, d.horizon_steps, d.time_dim, d.dim, d.n_levels = 23, 7, 6, 16, 64, 2  # Ta not a power of two
    d.dim_mults[0], d.dim_mults[1] = 1, 2
    d.kernel_size, d.n_groups, d.activation, d.cond_predict_scale, d.larger_encoder, d.groupnorm_eps = 5, 8, 1, 1, 1, 1e-5
    h = C.c_void_p()
    assert lib.dppo_unet_plan_create(C.byref(d), 20, 0, C.byref(h)) < 0
    assert b"power of two" in lib.dppo_last_error()
    d.horizon_steps, d.dim = 4, 128  # 16 channels x 4 positions per group = 64 features: more than a warp's 32 TMEM lanes
    assert lib.dppo_unet_plan_create(C.byref(d), 20, 0, C.byref(h)) < 0
    assert b"GroupNorm" in lib.dppo_last_error()
