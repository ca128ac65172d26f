"""
GPU tests of the tensor-core update path (csrc/update_gemm.cu, csrc/update_plan.cu behind dppo_update_*): every kernel
against float64 torch, the whole minibatch against the torch-autograd path and (through tests/test_gpu_parity.py, which
calls PPODiffusion.loss) against the golden loss / gradient vectors recorded from the reference.

reference: actor_ft forward / backward dppo/model/diffusion/diffusion_vpg.py:398-461, mlp_diffusion.py:218-250,
common/mlp.py:84-154; critic common/critic.py:40-54; loss.backward() train_ppo_diffusion_agent.py:360-364.
"""

import numpy as np
import pytest
import torch

from dppo_b200.workloads import get_workload
from tests.helpers import GOLDEN_CASES, build_model, make_inputs, our_classes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _lib():
    from dppo_b200 import _lib as L

    return L, L.load()


def _rnd(g, *s):
    return torch.randn(*s, generator=g).to(DEV)


def _mish(x):
    return x * torch.tanh(torch.nn.functional.softplus(x))


def _relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("c", [
    dict(R=300, K=64, N=512), dict(R=1000, K=512, N=512, act_out=1), dict(R=257, K=512, N=24), dict(R=130, K=256, N=1),
    dict(R=500, K=192, N=1024, act_out=2), dict(R=640, K=24, N=512, transposed=True, bias=False),
    dict(R=700, K=512, N=512, transposed=True, bias=False, pre=1, res=True),
    dict(R=333, K=512, N=512, transposed=True, bias=False, pre=2), dict(R=128, K=17, N=256, act_out=2),
    dict(R=999, K=1024, N=80), dict(R=1, K=64, N=64, act_out=1), dict(R=20000, K=512, N=512, act_out=1, res=True),
])
def test_row_gemm_matches_float64(c):
    """out = epilogue(x W^T): bias, act'(pre), residual, fp32 output and activation + hi/lo operand-image output."""
    L, lib = _lib()
    g = torch.Generator().manual_seed(1)
    R, K, N = c["R"], c["K"], c["N"]
    tr, pre, act_out = c.get("transposed", False), c.get("pre", 0), c.get("act_out", 0)
    x = _rnd(g, R, K)
    W = _rnd(g, K, N) if tr else _rnd(g, N, K)
    b = _rnd(g, N) if c.get("bias", True) else None
    p = _rnd(g, R, N) if pre else None
    r = _rnd(g, R, N) if c.get("res") else None
    out = torch.full((R, N), float("nan"), device=DEV)
    oact = torch.full((R, N), float("nan"), device=DEV)
    L.check(lib.dppo_debug_linear(L.ptr(x), R, K, L.ptr(W), N, int(tr), L.ptr(b), L.ptr(p), pre, L.ptr(r), L.ptr(out), act_out,
                                  L.ptr(oact), L.stream_ptr()), "dppo_debug_linear")
    torch.cuda.synchronize()
    ref = x.double() @ ((W.t() if tr else W).double().t())
    if b is not None:
        ref = ref + b.double()
    if pre:
        pd = p.double().requires_grad_(True)
        (gr,) = torch.autograd.grad((torch.relu(pd) if pre == 1 else _mish(pd)).sum(), pd)
        ref = ref * gr
    if r is not None:
        ref = ref + r.double()
    ra = ref if act_out == 0 else (torch.relu(ref) if act_out == 1 else _mish(ref))
    assert _relerr(out, ref) < 3e-5 and _relerr(oact, ra) < 3e-5, (_relerr(out, ref), _relerr(oact, ra))


@pytest.mark.parametrize("R,N,K", [(300, 512, 512), (1000, 24, 512), (777, 512, 64), (130, 1, 256), (600, 1024, 192), (50, 256, 17),
                                   (64, 128, 64), (8192, 256, 256), (1, 64, 64), (20000, 512, 512)])
def test_wgrad_matches_float64(R, N, K):
    """dW += g^T x over MN-major operand images (contraction over rows), bias gradient from the ones-tile MMA."""
    L, lib = _lib()
    g = torch.Generator().manual_seed(2)
    gm, x = _rnd(g, R, N), _rnd(g, R, K)
    dW = torch.ones(N, K, device=DEV)  # accumulation (+=) on top of existing contents
    db = torch.ones(N, device=DEV)
    L.check(lib.dppo_debug_wgrad(L.ptr(gm), L.ptr(x), R, N, K, L.ptr(dW), L.ptr(db), L.stream_ptr()), "dppo_debug_wgrad")
    torch.cuda.synchronize()
    assert _relerr(dW - 1, gm.double().t() @ x.double()) < 3e-5
    assert _relerr(db - 1, gm.double().sum(0)) < 3e-5


def _loss_grads(case, mode, monkeypatch, rows):
    monkeypatch.setenv("DPPO_B200_UPDATE", mode)
    spec = GOLDEN_CASES[case]
    w = get_workload(spec["workload"])
    E, ft = spec["n_envs"], w["ft_denoising_steps"]
    inp = make_inputs(w, E, rows)
    model = build_model(w, DEV, our_classes())
    assert (model.fused_update_reason() is None) == (mode == "fused")
    with torch.no_grad():
        out = model(cond={"state": inp["state"].to(DEV)}, noise=inp["noise"].to(DEV))
        lp = model.get_logprobs({"state": inp["state"].to(DEV)}, out.chains).view(E, ft, w["horizon_steps"], w["action_dim"])
    b, d = inp["mb_b"].to(DEV), inp["mb_d"].to(DEV)
    state = inp["state"].to(DEV)
    r = model.loss({"state": state[b]}, out.chains[b, d], out.chains[b, d + 1], d, inp["returns"].to(DEV)[b],
                   inp["oldvalues"].to(DEV)[b], inp["advantages"].to(DEV)[b], lp[b, d] + inp["lp_shift"].to(DEV),
                   reward_horizon=w["act_steps"])
    (r[0] + 0.5 * r[2]).backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in
             list(model.actor_ft.named_parameters()) + [("critic." + n, p) for n, p in model.critic.named_parameters()]}
    return [float(r[0].detach()), float(r[2].detach()), r[3], r[4], r[5]], grads


@pytest.mark.parametrize("case", ["hopper", "walker2d", "transport_k20", "furniture"])
def test_fused_update_matches_autograd_path(case, monkeypatch):
    """The whole minibatch (ReLU bit masks, Mish, LayerNorm + cond_mlp, time-MLP backward) vs the torch-autograd path."""
    sa, ga = _loss_grads(case, "autograd", monkeypatch, 300)
    sf, gf = _loss_grads(case, "fused", monkeypatch, 300)
    np.testing.assert_allclose(np.array(sf)[[0, 1, 3, 4]], np.array(sa)[[0, 1, 3, 4]], rtol=2e-4, atol=1e-7)
    for n, g_ref in ga.items():
        assert _relerr(gf[n], g_ref) < 5e-4, (n, _relerr(gf[n], g_ref))


def _buffers(model, w, E, n_steps, seed=3):
    ft, Ta, Da = w["ft_denoising_steps"], w["horizon_steps"], w["action_dim"]
    g = torch.Generator(device=DEV).manual_seed(seed)
    N = E * n_steps
    obs_k = torch.rand((N, w["cond_steps"], w["obs_dim"]), device=DEV, generator=g) * 2 - 1
    with torch.no_grad():
        chains_k = model(cond={"state": obs_k}).chains.contiguous()
        lp_k = model.get_logprobs({"state": obs_k}, chains_k).view(N, ft, Ta, Da).contiguous()
        val_k = model.critic({"state": obs_k}).view(-1).contiguous()
    adv_k = torch.randn(N, device=DEV, generator=g)
    return obs_k, chains_k, lp_k, (adv_k + val_k).contiguous(), val_k, adv_k


@pytest.mark.parametrize("case", ["hopper", "furniture"])
def test_update_minibatch_equals_loss_and_shards_sum(case):
    """dppo_update_minibatch (gathers fused, gradients accumulated into .grad) == PPODiffusion.loss + backward on the
    gathered rows; two half-minibatch 'ranks' accumulate to the whole minibatch."""
    w = get_workload(GOLDEN_CASES[case]["workload"])
    model = build_model(w, DEV, our_classes())
    ft = w["ft_denoising_steps"]
    obs_k, chains_k, lp_k, ret_k, val_k, adv_k = _buffers(model, w, 32, 3)
    N = obs_k.shape[0]
    g = torch.Generator().manual_seed(11)
    inds = torch.randperm(N * ft, generator=g)[:384].to(DEV)
    b, d = inds // ft, inds % ft
    params = list(model.actor_ft.parameters()) + list(model.critic.parameters())

    def grads_of(fn):
        for p in params:
            p.grad = torch.zeros_like(p)
        out = fn()
        torch.cuda.synchronize()
        return out, [p.grad.detach().clone() for p in params]

    def ref():
        r = model.loss({"state": obs_k[b]}, chains_k[b, d], chains_k[b, d + 1], d, ret_k[b], val_k[b], adv_k[b], lp_k[b, d],
                       reward_horizon=w["act_steps"])
        (r[0] + 0.5 * r[2]).backward()
        return [float(r[0].detach()), float(r[2].detach()), r[4], r[3], r[5]]

    s_ref, g_ref = grads_of(ref)
    s_all, g_all = grads_of(lambda: model.update_minibatch(obs_k, chains_k, lp_k, ret_k, val_k, adv_k, inds,
                                                           reward_horizon=w["act_steps"], vf_coef=0.5).tolist())
    np.testing.assert_allclose(s_all[:5], s_ref, rtol=1e-3, atol=1e-6)  # pg_loss and approx_kl sit at round-off level (ratio == 1)
    for a, r_ in zip(g_all, g_ref):
        assert _relerr(a, r_) < 1e-4

    def halves():
        model.update_minibatch(obs_k, chains_k, lp_k, ret_k, val_k, adv_k, inds, row_begin=0, row_count=200,
                               reward_horizon=w["act_steps"], vf_coef=0.5)
        model.update_minibatch(obs_k, chains_k, lp_k, ret_k, val_k, adv_k, inds, row_begin=200, row_count=184,
                               reward_horizon=w["act_steps"], vf_coef=0.5)

    _, g_half = grads_of(halves)
    for a, r_ in zip(g_half, g_all):
        assert _relerr(a, r_) < 1e-4
    # critic warm-up: with_actor = 0 leaves the actor gradients untouched
    n_actor = len(list(model.actor_ft.parameters()))
    _, g_c = grads_of(lambda: model.update_minibatch(obs_k, chains_k, lp_k, ret_k, val_k, adv_k, inds, reward_horizon=w["act_steps"],
                                                     vf_coef=0.5, with_actor=False))
    assert all(float(x.abs().max()) == 0.0 for x in g_c[:n_actor])
    for a, r_ in zip(g_c[n_actor:], g_all[n_actor:]):
        assert _relerr(a, r_) < 1e-4


def test_ratio_is_one_before_any_optimiser_step():
    """Old log-probs come from the chain kernel, new ones from the update kernels: at identical weights the PPO ratio must be
    1 to far better than clip_ploss_coef_base (1e-3), row by row."""
    w = get_workload("walker2d")
    model = build_model(w, DEV, our_classes())
    ft, D = w["ft_denoising_steps"], w["horizon_steps"] * w["action_dim"]
    obs_k, chains_k, lp_k, ret_k, val_k, adv_k = _buffers(model, w, 256, 2)
    N = obs_k.shape[0]
    inds = torch.randperm(N * ft, generator=torch.Generator().manual_seed(5))[:4096].to(DEV)
    b, d = inds // ft, inds % ft
    from dppo_b200.update_engine import UpdatePlan

    plan = model.update_plan(4096)
    plan.bind_model(model)
    eps = torch.empty((4096, D), device=DEV)
    vpred = torch.empty(4096, device=DEV)
    batch = UpdatePlan._batch(4096, 4096, 0, obs=obs_k, chains=chains_k, old_logprobs=lp_k, returns=ret_k, old_values=val_k,
                              advantages=adv_k, inds_all=inds)
    plan.forward(batch, eps, vpred)
    new_lp, _ = model.engine(sync=False).logprob_rows(eps, chains_k[b, d].reshape(4096, D), chains_k[b, d + 1].reshape(4096, D), d)
    old = lp_k[b, d].reshape(4096, D).clamp(-5, 2).mean(-1)
    new = new_lp.clamp(-5, 2).mean(-1)
    ratio = torch.exp(new - old)
    assert float((ratio - 1).abs().max()) < 1e-4, float((ratio - 1).abs().max())
    assert float((vpred - val_k[b]).abs().max()) < 1e-4 * max(1.0, float(val_k.abs().max()))


@pytest.mark.parametrize("case,rows", [("hopper", 1), ("hopper", 1000), ("walker2d", 40000), ("furniture", 777), ("transport", 130)])
def test_values_match_the_critic_module_in_float64(case, rows):
    """PPODiffusion.values (dppo_update_values: the critic through the tcgen05 row GEMMs, chunked over the plan's
    workspace) against the critic module evaluated in float64 - the value pass of train_ppo_diffusion_agent.py:197-206."""
    import copy

    w = get_workload(GOLDEN_CASES[case]["workload"])
    model = build_model(w, DEV, our_classes())
    assert model.fused_update_reason() is None
    g = torch.Generator().manual_seed(11)
    obs = (torch.rand(rows, w["cond_steps"], w["obs_dim"], generator=g) * 2 - 1).to(DEV)
    got = model.values({"state": obs})
    ref = copy.deepcopy(model.critic).double()
    with torch.no_grad():
        want = ref({"state": obs.double()}).view(-1)
    assert got.shape == (rows,)
    assert _relerr(got, want) < 1e-4, _relerr(got, want)
    again = model.values({"state": obs[: min(rows, 7)]})  # a second, shorter call reuses the bound plan
    assert _relerr(again, want[: min(rows, 7)]) < 1e-4
