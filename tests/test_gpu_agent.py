"""The B200 agent loop end to end on cuda:0 (tiny Hopper-shaped run) and its pieces against the oracle."""

import numpy as np
import pytest
import torch

from dppo_b200.workloads import get_workload, make_agent_cfg

pytestmark = pytest.mark.gpu


def _agent(tmp_path, workload="hopper", **kw):
    from dppo_b200.agent.finetune.train_ppo_diffusion_agent import TrainPPODiffusionAgent

    w = get_workload(workload)
    cfg = make_agent_cfg(w, "cuda:0", str(tmp_path), **kw)
    return w, TrainPPODiffusionAgent(cfg)


def test_prologue_matches_oracle_gae_and_model_calls(tmp_path):
    from oracle import dppo_oracle as O

    w, ag = _agent(tmp_path, n_envs=8, n_steps=6, batch_size=64, update_epochs=1)
    firsts = np.zeros((ag.n_steps + 1, ag.n_envs))
    firsts[0] = 1
    obs0 = ag.reset_env_all()
    ag.model.train()
    obs_buf, chains_buf, rew, term, last_obs, done, _ = ag.rollout(obs0, False, firsts)
    rew0 = rew.copy()
    values, logprobs, adv, ret = ag.prologue(obs_buf, chains_buf, rew, term, firsts, last_obs)
    # values / log-probs are the model's own calls on the flattened (step, env) rows
    obs_k = obs_buf.view(-1, 1, w["obs_dim"])
    with torch.no_grad():
        v_ref = ag.model.values({"state": obs_k}).view(ag.n_steps, ag.n_envs)
        v_mod = ag.model.critic({"state": obs_k}).view(ag.n_steps, ag.n_envs)  # the torch module (library GEMM route)
        lp_ref = ag.model.get_logprobs({"state": obs_k}, chains_buf.view(-1, *chains_buf.shape[2:]))
    assert torch.equal(values, v_ref) and torch.equal(logprobs.view_as(lp_ref), lp_ref)
    np.testing.assert_allclose(values.cpu().numpy(), v_mod.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # GAE (float64 kernel) against the oracle's restatement of the reference's numpy loop, with the same scaled rewards
    scaler = O.RunningRewardScaler(ag.n_envs)
    rew_s = scaler(rew0.T, firsts[:-1].T).T
    with torch.no_grad():
        nxt = ag.model.values({"state": torch.from_numpy(last_obs["state"]).cuda()}).double().cpu().numpy()
    adv_o, ret_o = O.gae(rew_s, term, values.double().cpu().numpy(), nxt, ag.gamma, ag.gae_lambda, 1.0)
    np.testing.assert_allclose(adv.cpu().numpy(), adv_o, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ret.cpu().numpy(), ret_o, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("workload,n_envs", [("hopper", 8), ("furniture", 70), ("square_unet", 24)])
def test_host_buffer_rollout_equals_device_call_rollout(tmp_path, workload, n_envs):
    """rollout() with one dppo_sample_chain_host call per decision (observations from one page-locked buffer, chains into
    the device-resident rollout buffer, actions into page-locked memory) == the device call with explicit copies, bit for
    bit: same Philox keys, same simulator trajectory."""
    outs = []
    for host in (True, False):
        w, ag = _agent(tmp_path / str(host), workload, n_envs=n_envs, n_steps=5, batch_size=64, update_epochs=1)
        ag.host_rollout = host
        torch.manual_seed(77)
        firsts = np.zeros((ag.n_steps + 1, ag.n_envs))
        firsts[0] = 1
        ag.model.train()
        obs_buf, chains_buf, rew, term, last_obs, done, steps = ag.rollout(ag.reset_env_all(), False, firsts)
        torch.cuda.synchronize()
        outs.append((obs_buf.cpu(), chains_buf.cpu(), rew.copy(), term.copy(), last_obs["state"].copy(), firsts.copy(), steps))
        assert not ag.model.engine(sync=False).nonfinite()
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[6] == b[6]
    for i in (2, 3, 4, 5):
        assert np.array_equal(a[i], b[i])
    assert float(a[1].abs().max()) > 0


def test_run_two_iterations_updates_and_checkpoints(tmp_path):
    w, ag = _agent(tmp_path, n_envs=8, n_steps=8, batch_size=128, update_epochs=2, n_train_itr=2)
    before = {k: v.clone() for k, v in ag.model.actor_ft.state_dict().items()}
    base = {k: v.clone() for k, v in ag.model.actor.state_dict().items()}
    res = ag.run()
    assert len(res) == 2 and all(np.isfinite(r["pg_loss"]) and np.isfinite(r["v_loss"]) for r in res)
    assert res[-1]["step"] == 2 * 8 * 8 * w["act_steps"] and res[-1]["minibatches"] >= 1
    assert any(not torch.equal(before[k], v) for k, v in ag.model.actor_ft.state_dict().items())
    assert all(torch.equal(base[k], v) for k, v in ag.model.actor.state_dict().items())  # frozen base policy
    ckpt = torch.load(tmp_path / "checkpoint" / "state_1.pt", weights_only=True)
    keys = set(k.split(".")[0] for k in ckpt["model"])
    assert ckpt["itr"] == 1 and {"network", "actor", "actor_ft", "critic"} <= keys


def test_update_equals_plain_autograd_minibatch(tmp_path):
    """One minibatch through the flat-gradient path == model.loss (reference signature) + autograd on the same rows."""
    w, ag = _agent(tmp_path, n_envs=8, n_steps=6, batch_size=96, update_epochs=1)
    firsts = np.zeros((ag.n_steps + 1, ag.n_envs))
    obs_buf, chains_buf, rew, term, last_obs, _, _ = ag.rollout(ag.reset_env_all(), False, firsts)
    values, logprobs, adv, ret = ag.prologue(obs_buf, chains_buf, rew, term, firsts, last_obs)
    m, ft = ag.model, ag.model.ft_denoising_steps
    N = ag.n_steps * ag.n_envs
    obs_k, chains_k = obs_buf.view(N, 1, -1), chains_buf.view(N, ft + 1, *chains_buf.shape[3:])
    lp_k = logprobs.view(N, ft, *logprobs.shape[3:])
    inds = torch.randperm(N * ft, device="cuda")[:96]
    ag.grads.zero()
    r1 = m.loss_gathered(obs_k, chains_k, lp_k, ret.reshape(-1), values.reshape(-1), adv.reshape(-1), inds,
                         reward_horizon=ag.reward_horizon, scalars_out=ag.grads.scalars)
    (r1[0] + 0.5 * r1[2]).backward()
    g_flat = torch.cat([p.grad.reshape(-1).clone() for p in ag.grads.params])  # groups are padded to 16 bytes
    for p in ag.grads.params:
        p.grad = None
    b, d = inds // ft, inds % ft
    r0 = m.loss({"state": obs_k[b]}, chains_k[b, d], chains_k[b, d + 1], d, ret.reshape(-1)[b], values.reshape(-1)[b],
                adv.reshape(-1)[b], lp_k[b, d], reward_horizon=ag.reward_horizon)
    (r0[0] + 0.5 * r0[2]).backward()
    g_ref = torch.cat([p.grad.reshape(-1) for p in ag.grads.params])
    assert float((g_flat - g_ref).abs().max()) <= 1e-5 * max(1.0, float(g_ref.abs().max()))
    assert abs(float(ag.grads.scalars[0]) - float(r0[0])) <= 1e-6 * max(1.0, abs(float(r0[0])))


def test_unet_agent_runs_two_iterations(tmp_path):
    """cfg5 (Unet1D denoiser, DDIM-10): rollout through the unet chain kernel, update through autograd + the loss kernel."""
    w, ag = _agent(tmp_path, workload="square_unet", n_envs=8, n_steps=4, batch_size=64, update_epochs=1, n_train_itr=3)
    ag.n_critic_warmup_itr = 1
    before = {k: v.clone() for k, v in ag.model.actor_ft.state_dict().items()}
    res = ag.run()
    assert len(res) == 3 and all(np.isfinite(r["pg_loss"]) and np.isfinite(r["v_loss"]) for r in res)
    assert any(not torch.equal(before[k], v) for k, v in ag.model.actor_ft.state_dict().items())


@pytest.mark.parametrize("clip", [None, 0.05])
def test_flat_adamw_matches_torch_adamw(clip):
    """dppo_adamw_flat == torch.optim.AdamW (+ clip_grad_norm_) over several steps, odd sizes, both flat-buffer groups."""
    from dppo_b200 import distributed as D
    from dppo_b200.optim import FlatAdamW

    torch.manual_seed(0)
    shapes = [(37, 19), (19,), (5, 7, 3), (1,), (130, 64)]
    mine = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    other = [torch.nn.Parameter(torch.randn(3, 11, device="cuda"))]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    opt = FlatAdamW(mine, lr=3e-3, weight_decay=0.01)
    opt2 = FlatAdamW(other, lr=1e-3, weight_decay=0.0)
    buf = D.FlatGradBuffer([other, mine])  # `mine` is the SECOND group: its segment starts on a padded boundary
    topt = torch.optim.AdamW(ref, lr=3e-3, weight_decay=0.01)
    for it in range(6):
        for p, r in zip(mine, ref):
            g = torch.randn_like(p) * (0.1 + it)
            p.grad.copy_(g)
            r.grad = g.clone()
        other[0].grad.fill_(0.5)
        v0 = [p._version for p in mine]
        opt.param_groups[0]["lr"] = topt.param_groups[0]["lr"] = 3e-3 * (1 + 0.1 * it)
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(ref, clip)
        opt.step(max_grad_norm=clip)
        opt2.step()
        topt.step()
        assert all(p._version > v for p, v in zip(mine, v0))
        for p, r in zip(mine, ref):
            assert torch.allclose(p, r, rtol=2e-6, atol=2e-7), (it, float((p - r).abs().max()))
    assert buf.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in other + mine) + 8  # every tensor padded to 16 bytes


def test_reward_scaler_kernel_matches_host_mirror():
    """dppo_reward_scale_f64 (three phases) == RunningRewardScaler over successive iterations, state included."""
    from dppo_b200.util.reward_scaling import RunningRewardScaler, RunningRewardScalerCUDA

    rng = np.random.default_rng(9)
    for n, E in [(500, 40), (7, 1), (88, 1000)]:
        host, dev = RunningRewardScaler(E), RunningRewardScalerCUDA(E, "cuda:0")
        for it in range(3):
            r = rng.standard_normal((n, E)) * (1 + 5 * it)
            first = (rng.random((n, E)) < 0.03).astype(np.float64)
            first[0] = it == 0
            want = host(reward=r.T, first=first.T).T
            got = dev(reward=torch.from_numpy(r).cuda(), first=torch.from_numpy(first).cuda())
            np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
            hs, ds = host.state_dict(), dev.state_dict()
            np.testing.assert_allclose(ds["ret"], hs["ret"], rtol=1e-12, atol=1e-12)
            for k in ("mean", "var", "count"):
                assert abs(ds[k] - hs[k]) <= 1e-11 * max(1.0, abs(hs[k])), (k, ds[k], hs[k])
        fresh = RunningRewardScalerCUDA(E, "cuda:0")
        fresh.load_state_dict(host.state_dict())
        assert abs(fresh.state_dict()["var"] - host.var) <= 1e-15 * max(1.0, host.var)


@pytest.mark.parametrize("workload", ["hopper", "furniture", "square_unet"])
def test_diffusion_eval_loads_checkpoint_and_matches_rollout_model(tmp_path, workload):
    """DiffusionEval (checkpoint key remapping, base + fine-tuned nets) samples what PPODiffusion samples deterministically."""
    from dppo_b200.model.diffusion.diffusion_eval import DiffusionEval
    from tests.helpers import build_model, make_inputs, our_classes

    w = get_workload(workload)
    model = build_model(w, "cuda:0", our_classes())
    path = tmp_path / "state.pt"
    torch.save({"itr": 3, "model": model.state_dict()}, path)
    cls = our_classes()
    a = dict(w["actor"])
    kind = a.pop("kind")
    cond_dim = w["obs_dim"] * w["cond_steps"]
    net = (cls["mlp"](action_dim=w["action_dim"], horizon_steps=w["horizon_steps"], cond_dim=cond_dim, **a) if kind == "mlp"
           else cls["unet"](action_dim=w["action_dim"], cond_dim=cond_dim, **a))
    ev = DiffusionEval(network_path=str(path), ft_denoising_steps=w["ft_denoising_steps"], network=net,
                       horizon_steps=w["horizon_steps"], obs_dim=w["obs_dim"], action_dim=w["action_dim"], device="cuda:0",
                       denoising_steps=w["denoising_steps"], use_ddim=w["use_ddim"], ddim_steps=w["ddim_steps"],
                       randn_clip_value=w["ppo"]["randn_clip_value"])
    inp = make_inputs(w, 24, 8)
    state, noise = inp["state"].cuda(), inp["noise"].cuda()
    got = ev(cond={"state": state}, deterministic=True, noise=noise)
    want = model(cond={"state": state}, deterministic=True, noise=noise)
    assert got.chains is None and torch.equal(got.trajectories, want.trajectories)
    # a pre-training checkpoint (network.* only) needs ft_denoising_steps = 0
    torch.save({"model": {k: v for k, v in model.state_dict().items() if k.startswith("network.")}}, path)
    with pytest.raises(ValueError):
        DiffusionEval(network_path=str(path), ft_denoising_steps=2, network=net, horizon_steps=w["horizon_steps"],
                      obs_dim=w["obs_dim"], action_dim=w["action_dim"], device="cuda:0", denoising_steps=w["denoising_steps"],
                      use_ddim=w["use_ddim"], ddim_steps=w["ddim_steps"])
    pre = DiffusionEval(network_path=str(path), ft_denoising_steps=0, network=net, horizon_steps=w["horizon_steps"],
                        obs_dim=w["obs_dim"], action_dim=w["action_dim"], device="cuda:0", denoising_steps=w["denoising_steps"],
                        use_ddim=w["use_ddim"], ddim_steps=w["ddim_steps"], randn_clip_value=w["ppo"]["randn_clip_value"])
    base = model(cond={"state": state}, deterministic=True, noise=noise, use_base_policy=True)
    assert torch.equal(pre(cond={"state": state}, noise=noise).trajectories, base.trajectories)


def test_eval_agent_runs_on_the_synthetic_env(tmp_path):
    from dppo_b200.agent.eval.eval_diffusion_agent import EvalDiffusionAgent
    from dppo_b200.util.config import Cfg
    from tests.helpers import build_model, our_classes

    w = get_workload("hopper")
    model = build_model(w, "cuda:0", our_classes())
    path = tmp_path / "state.pt"
    torch.save({"itr": 0, "model": model.state_dict()}, path)
    cond_dim = w["obs_dim"] * w["cond_steps"]
    a = dict(w["actor"])
    a.pop("kind")
    cfg = Cfg(device="cuda:0", seed=7, logdir=str(tmp_path), obs_dim=w["obs_dim"], action_dim=w["action_dim"],
              cond_steps=w["cond_steps"], act_steps=w["act_steps"], horizon_steps=w["horizon_steps"], n_steps=30,
              env=dict(n_envs=6, name="synthetic", max_episode_steps=40, best_reward_threshold_for_success=0.1),
              model={"_target_": "dppo_b200.model.diffusion.diffusion_eval.DiffusionEval", "network_path": str(path),
                     "ft_denoising_steps": w["ft_denoising_steps"], "horizon_steps": w["horizon_steps"], "obs_dim": w["obs_dim"],
                     "action_dim": w["action_dim"], "device": "cuda:0", "denoising_steps": w["denoising_steps"],
                     "randn_clip_value": w["ppo"]["randn_clip_value"],
                     "network": dict({"_target_": "dppo_b200.model.diffusion.mlp_diffusion.DiffusionMLP",
                                      "action_dim": w["action_dim"], "horizon_steps": w["horizon_steps"], "cond_dim": cond_dim}, **a)})
    res = EvalDiffusionAgent(cfg).run()
    assert res["num_episode"] >= 6 and np.isfinite(res["eval_episode_reward"]) and 0.0 <= res["eval_success_rate"] <= 1.0
    saved = np.load(tmp_path / "eval.npz")
    assert int(saved["num_episode"]) == res["num_episode"]


@pytest.mark.parametrize("M,K,N", [(50000, 512, 512), (1000, 57, 512), (777, 512, 24), (64, 11, 256), (5, 8, 8)])
def test_split_linear_matches_fp32_linear(M, K, N):
    """SplitLinear (3-product bf16 split on the tensor cores) vs F.linear in float64: forward, dx, dW, db."""
    from dppo_b200.model.common import split_linear as SL

    torch.manual_seed(1)
    lin = SL.SplitLinear(K, N).cuda()
    x = torch.randn(M, K, device="cuda", requires_grad=True)
    gy = torch.randn(M, N, device="cuda")
    y = lin(x)
    y.backward(gy)
    got = [y.detach(), x.grad, lin.weight.grad, lin.bias.grad]
    xd = x.detach().double().requires_grad_(True)
    wd, bd = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(gy.double())
    want = [yd.detach(), xd.grad, wd.grad, bd.grad]
    for name, a, b in zip(("y", "dx", "dW", "db"), got, want):
        err = float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
        assert err < 3e-5, (name, err)
    # and it is not slower than the truth it replaces: plain fp32 path for reference
    SL.ENABLED = False
    try:
        y32 = lin(x.detach())
    finally:
        SL.ENABLED = True
    assert float((y32 - got[0]).abs().max() / got[0].abs().max()) < 3e-5


def test_graphed_minibatch_equals_eager(tmp_path):
    """One PPO minibatch replayed as a CUDA graph leaves the same gradients and diagnostics as the eager launches."""
    from dppo_b200.agent.finetune.graphed import GraphedMinibatch

    w, ag = _agent(tmp_path, n_envs=16, n_steps=8, batch_size=256, update_epochs=1)
    firsts = np.zeros((ag.n_steps + 1, ag.n_envs))
    obs_buf, chains_buf, rew, term, last_obs, _, _ = ag.rollout(ag.reset_env_all(), False, firsts)
    values, logprobs, adv, ret = ag.prologue(obs_buf, chains_buf, rew, term, firsts, last_obs)
    m, ft = ag.model, ag.model.ft_denoising_steps
    N = ag.n_steps * ag.n_envs
    obs_k, chains_k = obs_buf.view(N, 1, -1), chains_buf.view(N, ft + 1, *chains_buf.shape[3:])
    lp_k = logprobs.view(N, ft, *logprobs.shape[3:])

    def fwd_bwd(inds):
        ag.grads.zero()
        r = m.loss_gathered(obs_k, chains_k, lp_k, ret.reshape(-1), values.reshape(-1), adv.reshape(-1), inds,
                            reward_horizon=ag.reward_horizon, scalars_out=ag.grads.scalars)
        (r[0] + 0.5 * r[2]).backward()

    g = GraphedMinibatch(fwd_bwd, 256, "cuda:0")
    for seed in (1, 2):
        inds = torch.randperm(N * ft, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed))[:256]
        fwd_bwd(inds)
        eager = ag.grads.flat.clone()
        ag.grads.flat.fill_(float("nan"))
        g(inds)
        torch.cuda.synchronize()
        assert torch.isfinite(ag.grads.flat).all()
        assert float((ag.grads.flat - eager).abs().max()) <= 1e-6 * float(eager.abs().max())


def test_agent_run_with_graphed_update(tmp_path):
    """Enough minibatches per iteration for the agent to capture its update (>= 16): runs, learns, stays finite."""
    w, ag = _agent(tmp_path, n_envs=16, n_steps=8, batch_size=64, update_epochs=2, n_train_itr=2)
    assert ag.cuda_graph_update
    res = ag.run()
    assert res[-1]["minibatches"] == 2 * (16 * 8 * w["ft_denoising_steps"] // 64)
    assert all(np.isfinite(r["pg_loss"]) and np.isfinite(r["v_loss"]) and np.isfinite(r["approx_kl"]) for r in res)
    assert ag.cuda_graph_update  # capture did not fall back


@pytest.mark.parametrize("kind,cin,cout,ks,stride,pad,T", [
    ("conv", 7, 64, 5, 1, 2, 4), ("conv", 64, 64, 5, 1, 2, 4), ("conv", 256, 64, 5, 1, 2, 2), ("conv", 64, 64, 3, 2, 1, 4),
    ("conv", 64, 7, 1, 1, 0, 4), ("convT", 64, 64, 4, 2, 1, 2), ("conv", 32, 32, 3, 1, 1, 16),
])
def test_dense_lowered_convs_match_float64_convs(kind, cin, cout, ks, stride, pad, T):
    """DenseConv1d / DenseConvTranspose1d (gathered Toeplitz matrix + split-3 tensor-core GEMM) vs float64 convolutions."""
    import torch.nn.functional as F

    from dppo_b200.model.diffusion.dense_conv import DenseConv1d, DenseConvTranspose1d

    torch.manual_seed(2)
    mod = (DenseConv1d(cin, cout, ks, stride, pad) if kind == "conv" else DenseConvTranspose1d(cin, cout, ks, stride, pad)).cuda()
    x = torch.randn(300, cin, T, device="cuda", requires_grad=True)
    y = mod(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xd = x.detach().double().requires_grad_(True)
    wd, bd = mod.weight.detach().double().requires_grad_(True), mod.bias.detach().double().requires_grad_(True)
    yd = F.conv1d(xd, wd, bd, stride, pad) if kind == "conv" else F.conv_transpose1d(xd, wd, bd, stride, pad)
    assert yd.shape == y.shape
    yd.backward(gy.double())
    for name, a, b in (("y", y.detach(), yd.detach()), ("dx", x.grad, xd.grad), ("dW", mod.weight.grad, wd.grad),
                       ("db", mod.bias.grad, bd.grad)):
        err = float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
        assert err < 5e-5, (name, err)


def test_graft_entry_smoke_runs():
    """The driver's smoke() entry point (one Hopper chain + log-probs against the CPU oracle)."""
    import __graft_entry__ as g

    g.smoke()


@pytest.mark.parametrize("graph", [False, True])
def test_device_side_kl_early_stop_equals_host_loop(tmp_path, graph):
    """The device stop flag (dppo_kl_check + no-op AdamW launches, polled with a lag) applies exactly the minibatches the
    reference's host-side loop applies (train_ppo_diffusion_agent.py:313-382: step, then test approx_kl, then break)."""
    from dppo_b200 import distributed as D

    w, ag = _agent(tmp_path, n_envs=16, n_steps=8, batch_size=64, update_epochs=3)
    ag.cuda_graph_update = graph
    ag.cfg.train.actor_lr = 3e-3
    for grp in ag.actor_optimizer.param_groups:
        grp["lr"] = 3e-3  # large steps: approx_kl crosses the target after a few minibatches
    ag.target_kl = 2e-4
    firsts = np.zeros((ag.n_steps + 1, ag.n_envs))
    obs_buf, chains_buf, rew, term, last_obs, _, _ = ag.rollout(ag.reset_env_all(), False, firsts)
    values, logprobs, adv, ret = ag.prologue(obs_buf, chains_buf, rew, term, firsts, last_obs)
    opts = (ag.actor_optimizer, ag.critic_optimizer)
    snap = [(o.flat.clone(), o.exp_avg.clone(), o.exp_avg_sq.clone(), o._state.clone()) for o in opts]

    def restore():
        for o, (f, m, v, s) in zip(opts, snap):
            o.flat.copy_(f), o.exp_avg.copy_(m), o.exp_avg_sq.copy_(v), o._state.copy_(s)
            o.bump_versions()

    torch.cuda.manual_seed(77)
    stats = ag.update(obs_buf, chains_buf, logprobs, values, adv, ret)
    got = [o.flat.clone() for o in opts]
    steps_got = [o.step_count for o in opts]
    n_total = 3 * (16 * 8 * w["ft_denoising_steps"] // 64)
    assert 1 <= stats["minibatches"] < n_total, stats["minibatches"]  # the test tripped mid-way
    assert stats["approx_kl"] > ag.target_kl

    # the reference's loop on the host: same permutations, optimiser step, THEN the KL test
    restore()
    torch.cuda.manual_seed(77)
    m, ft = ag.model, ag.model.ft_denoising_steps
    N = ag.n_steps * ag.n_envs
    obs_k, chains_k = obs_buf.view(N, 1, -1), chains_buf.view(N, ft + 1, *chains_buf.shape[3:])
    lp_k = logprobs.view(N, ft, *logprobs.shape[3:])
    done, stop = 0, False
    for _ in range(ag.update_epochs):
        perm = D.broadcast_permutation(N * ft, "cuda:0")
        for b in range(N * ft // ag.batch_size):
            ag.grads.zero()
            m.update_minibatch(obs_k, chains_k, lp_k, ret.reshape(-1), values.reshape(-1), adv.reshape(-1),
                               perm[b * ag.batch_size:(b + 1) * ag.batch_size], reward_horizon=ag.reward_horizon, vf_coef=ag.vf_coef,
                               scalars_out=ag.grads.scalars)
            ag.actor_optimizer.step()
            ag.critic_optimizer.step()
            done += 1
            if float(ag.grads.scalars[2]) > ag.target_kl:
                stop = True
                break
        if stop:
            break
    assert stop and done == stats["minibatches"]
    assert [o.step_count for o in opts] == steps_got == [snap[0][3][0].item() + done, snap[1][3][0].item() + done]
    for a, o in zip(got, opts):
        assert float((a - o.flat).abs().max()) <= 2e-6 * max(1.0, float(o.flat.abs().max()))


def test_agent_run_matches_reference_run_golden(tmp_path):
    """Three iterations of TrainPPODiffusionAgent.run against the UNMODIFIED reference loop (tests/golden/agent_run.npz,
    recorded by make_golden_agent.py under the SURVEY §8c shim): same synthetic env, same injected draws, same recorded
    permutations.  Checks, per minibatch, the loss diagnostics and the sums of the loss inputs (reward scaling -> GAE with
    the bootstrap on the post-rollout observation -> tail-row drop -> (b, d) indexing), per iteration the number of
    minibatches (KL early stop in the last one) and the learning rates (critic warm-up, LR warm-up), and at the end the
    parameters of actor_ft / critic."""
    from dppo_b200.agent.finetune.train_ppo_diffusion_agent import TrainPPODiffusionAgent
    from tests.helpers import GOLDEN_DIR, PERTURB_SCALE, PERTURB_SEED, agent_golden_cfg

    gold = dict(np.load(f"{GOLDEN_DIR}/agent_run.npz", allow_pickle=False))
    cfg = agent_golden_cfg("cuda:0", str(tmp_path))
    ag = TrainPPODiffusionAgent(cfg)
    g = torch.Generator().manual_seed(PERTURB_SEED)
    with torch.no_grad():
        for p in ag.model.actor_ft.parameters():
            p.add_((PERTURB_SCALE * torch.randn(p.shape, generator=g)).to(p.device))
    for o in (ag.actor_optimizer, ag.critic_optimizer):
        o.bump_versions()
    gen = torch.Generator().manual_seed(int(gold["noise_seed"]))
    S = ag.model.denoising_steps
    shape = (cfg.horizon_steps, cfg.action_dim)
    perms = [torch.from_numpy(p.astype(np.int64)) for p in gold["perms"]]
    ag.test_hooks["noise"] = lambda E: torch.stack([torch.randn((E,) + shape, generator=gen) for _ in range(S + 1)])
    ag.test_hooks["perm"] = lambda n: perms.pop(0)
    records, update0 = [], ag.update

    def update(obs_buf, chains_buf, logprobs, values, adv, ret):
        n_before = len(gold["perms"]) - len(perms)
        lrs = (ag.actor_optimizer.param_groups[0]["lr"], ag.critic_optimizer.param_groups[0]["lr"])
        stats = update0(obs_buf, chains_buf, logprobs, values, adv, ret)
        used = [torch.from_numpy(p.astype(np.int64)) for p in gold["perms"][n_before:len(gold["perms"]) - len(perms)]]
        ft, bs = ag.model.ft_denoising_steps, ag.batch_size
        k = 0
        for perm in used:
            for b0 in range(0, (perm.numel() // bs) * bs, bs):
                if k >= len(ag.last_history):
                    break
                idx = perm[b0:b0 + bs].cuda()
                bb, dd = idx // ft, idx % ft
                h = ag.last_history[k]
                records.append([h[0], h[1], h[3], h[2], h[4], float(ret.reshape(-1)[bb].double().sum()),
                                float(values.reshape(-1)[bb].double().sum()), float(adv.reshape(-1)[bb].double().sum()),
                                float(logprobs.reshape(-1, ft, *logprobs.shape[3:])[bb, dd].double().sum()), float(dd.double().sum()),
                                ag.itr, lrs[0], lrs[1]])
                k += 1
        return stats

    ag.update = update
    res = ag.run()
    got, ref = np.array(records, dtype=np.float64), gold["minibatch"]
    assert got.shape == ref.shape, (got.shape, ref.shape)  # same number of applied minibatches in every iteration
    assert [r["minibatches"] for r in res] == [int((ref[:, 10] == i).sum()) for i in range(3)]
    np.testing.assert_array_equal(got[:, 9:11], ref[:, 9:11])                         # (b, d) indexing and iteration
    np.testing.assert_allclose(got[:, 11:13], ref[:, 11:13], rtol=1e-12)               # learning rates
    np.testing.assert_allclose(got[:, 5:9], ref[:, 5:9], rtol=2e-3, atol=2e-3)         # returns / values / advantages / old log-probs
    np.testing.assert_allclose(got[:, 0:2], ref[:, 0:2], rtol=5e-3, atol=2e-5)         # pg_loss, v_loss
    np.testing.assert_allclose(got[:, 3], ref[:, 3], rtol=2e-2, atol=2e-7)             # approx_kl
    np.testing.assert_allclose(got[:, 4], ref[:, 4], rtol=1e-4)                        # ratio
    assert np.abs(got[:, 2] - ref[:, 2]).max() <= 0.05                                 # clipfrac (rows on the clip boundary may flip)
    sd = ag.model.state_dict()
    for name, (s1, s2) in zip(gold["param_names"], gold["param_sums"]):
        v = sd[str(name)].double()
        assert abs(float(v.pow(2).sum()) - s2) <= 2e-3 * max(s2, 1e-12), name
    np.testing.assert_allclose(sd["critic.Q1.layers.2.bias"].cpu().numpy(), gold["critic_out_bias"], atol=5e-4)
    np.testing.assert_allclose(sd["actor_ft.mlp_mean.layers.2.bias"].cpu().numpy(), gold["actor_out_bias"], atol=2e-3)
