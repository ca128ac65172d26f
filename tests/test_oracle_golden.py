"""
The oracle (oracle/dppo_oracle.py) against the golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.  Forward paths must agree bit-for-bit; gradients to round-off.
"""

import numpy as np
import pytest
import torch

from dppo_b200.workloads import get_workload
from oracle import dppo_oracle as O
from tests.helpers import GOLDEN_CASES, build_model, load_golden, make_inputs, oracle_cfgs, oracle_params, our_classes, param_checksums

FAST_CASES = ["hopper", "walker2d", "transport_k20", "furniture", "square_unet", "kitchen", "avoid", "square_mlp", "can_unet"]


def _setup(case):
    spec = GOLDEN_CASES[case]
    w = get_workload(spec["workload"])
    model = build_model(w, "cpu", our_classes())
    gold = load_golden(case)
    inp = make_inputs(w, spec["n_envs"], spec["mb_rows"])
    nc, dc = oracle_cfgs(w)
    return w, model, gold, inp, nc, dc


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_seeded_weights_match_reference(case):
    w, model, gold, *_ = _setup(case)
    names, sums = param_checksums(model)
    assert names == list(gold["param_names"])
    np.testing.assert_array_equal(np.array(sums), gold["param_sums"])


@pytest.mark.parametrize("case", FAST_CASES)
def test_chain_and_logprobs(case):
    w, model, gold, inp, nc, dc = _setup(case)
    p = oracle_params(model)
    torch.set_num_threads(8)
    traj, chains = O.sample_chain(p, nc, dc, inp["state"], inp["noise"], deterministic=False)
    np.testing.assert_array_equal(traj.numpy(), gold["traj"])
    np.testing.assert_array_equal(chains.numpy(), gold["chains"])
    traj_d, chains_d = O.sample_chain(p, nc, dc, inp["state"], inp["noise"], deterministic=True)
    np.testing.assert_array_equal(traj_d.numpy(), gold["traj_det"])
    np.testing.assert_array_equal(chains_d.numpy(), gold["chains_det"])
    with torch.no_grad():
        lp = O.get_logprobs(p, nc, dc, inp["state"], chains)
        v = O.critic_obs(p, nc, inp["state"])
    np.testing.assert_array_equal(lp.numpy(), gold["logprobs"])
    np.testing.assert_array_equal(v.numpy(), gold["values"])
    # one-network-per-row evaluation (what the kernels do) gives the same values as the reference's two-network pass
    traj2, chains2 = O.sample_chain(p, nc, dc, inp["state"], inp["noise"], faithful_cost=False)
    np.testing.assert_allclose(chains2.numpy(), gold["chains"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("case", ["hopper", "furniture", "square_unet", "kitchen", "can_unet"])
def test_loss_and_gradients(case):
    w, model, gold, inp, nc, dc = _setup(case)
    p = oracle_params(model, requires_grad=True)
    E, ft = GOLDEN_CASES[case]["n_envs"], w["ft_denoising_steps"]
    chains = torch.from_numpy(gold["chains"])
    lp_k = torch.from_numpy(gold["logprobs"]).reshape(E, ft, w["horizon_steps"], w["action_dim"])
    b, d = inp["mb_b"], inp["mb_d"]
    res = O.ppo_loss(p, nc, dc, inp["state"][b], chains[b, d], chains[b, d + 1], d, inp["returns"][b], inp["oldvalues"][b],
                     inp["advantages"][b], lp_k[b, d] + inp["lp_shift"], reward_horizon=w["act_steps"])
    (res[0] + 0.5 * res[2]).backward()
    got = np.array([float(res[0]), float(res[1]), float(res[2]), res[3], res[4], res[5], float(res[6]), res[7]])
    np.testing.assert_allclose(got, gold["loss_scalars"], rtol=1e-6, atol=1e-9)
    last = "actor_ft." + str(gold["grad_last_name"])
    np.testing.assert_allclose(p[last].grad.numpy(), gold["grad_last"], rtol=1e-4, atol=1e-9)
    for name, (norm, total) in zip(gold["grad_names"], gold["grad_stats"]):
        key = str(name) if str(name).startswith("critic.") else "actor_ft." + str(name)
        g = p[key].grad
        assert abs(float(g.double().norm()) - norm) <= 1e-4 * max(norm, 1e-12), key


def test_gae_matches_reference_loop():
    """oracle.gae against a literal transcription of the reference's per-step numpy loop on random data."""
    rng = np.random.default_rng(3)
    n, E = 37, 11
    r, v = rng.standard_normal((n, E)), rng.standard_normal((n, E))
    term = (rng.random((n, E)) < 0.1).astype(np.float64)
    nxt = rng.standard_normal(E)
    adv, ret = O.gae(r, term, v, nxt, 0.99, 0.95, 0.7)
    ref = np.zeros_like(r)
    last = 0
    for t in reversed(range(n)):
        nv = nxt if t == n - 1 else v[t + 1]
        nonterminal = 1.0 - term[t]
        delta = r[t] * 0.7 + 0.99 * nv * nonterminal - v[t]
        ref[t] = last = delta + 0.99 * 0.95 * nonterminal * last
    np.testing.assert_array_equal(adv, ref)
    np.testing.assert_array_equal(ret, ref + v)


def test_minibatch_indices_bit_exact():
    g = torch.Generator().manual_seed(7)
    N, ft, bs = 200, 10, 300
    perm = torch.randperm(N * ft, generator=g)
    for k in range(N * ft // bs):
        b, d = O.minibatch_indices(perm, k, bs, ft)
        bb, dd = torch.unravel_index(perm[k * bs:(k + 1) * bs], (N, ft))
        assert torch.equal(b, bb) and torch.equal(d, dd)


# ------------------------------------------------------------------------------------------------ optional branches
def _variant(name):
    from tests.helpers import VARIANTS, perturb_again, variant_workload

    spec = VARIANTS[name]
    w = variant_workload(name)
    model = build_model(w, "cpu", our_classes())
    gold = dict(np.load(f"{__import__('tests.helpers').helpers.GOLDEN_DIR}/variant_{name}.npz", allow_pickle=False))
    inp = make_inputs(w, spec["n_envs"], spec["mb_rows"], seed=spec.get("seed", 0))
    if spec.get("anneal"):
        model.step()
        perturb_again(model)
        inp = make_inputs(w, spec["n_envs"], spec["mb_rows"], seed=spec.get("seed", 0), ft=model.ft_denoising_steps)
        w["ft_denoising_steps"] = int(model.ft_denoising_steps)
    nc, dc = oracle_cfgs(w)
    return spec, w, model, gold, inp, nc, dc


@pytest.mark.parametrize("name", ["bc", "vclip_quant", "epsclip", "finalclip", "anneal"])
def test_variant_chain_and_logprobs(name):
    """eps_clip_value, final_action_clip_value and the annealed fine-tuning window (one VPGDiffusion.step()) bit for bit."""
    spec, w, model, gold, inp, nc, dc = _variant(name)
    if spec.get("anneal"):
        assert int(gold["ft_after"]) == model.ft_denoising_steps == 7
    p = oracle_params(model)
    torch.set_num_threads(8)
    traj, chains = O.sample_chain(p, nc, dc, inp["state"], inp["noise"])
    np.testing.assert_array_equal(chains.numpy(), gold["chains"])
    np.testing.assert_array_equal(traj.numpy(), gold["traj"])
    with torch.no_grad():
        lp = O.get_logprobs(p, nc, dc, inp["state"], chains)
    np.testing.assert_array_equal(lp.numpy(), gold["logprobs"])


@pytest.mark.parametrize("name", ["bc", "vclip_quant", "epsclip"])
def test_variant_loss_and_gradients(name):
    """use_bc_loss, clip_vloss_coef + advantage quantiles, eps_clip_value through the loss (scalars and gradient norms)."""
    spec, w, model, gold, inp, nc, dc = _variant(name)
    p = oracle_params(model, requires_grad=True)
    E, ft = spec["n_envs"], w["ft_denoising_steps"]
    chains = torch.from_numpy(gold["chains"])
    lp_k = torch.from_numpy(gold["logprobs"]).reshape(E, ft, w["horizon_steps"], w["action_dim"])
    b, d = inp["mb_b"], inp["mb_d"]
    bc_noise = None
    if spec.get("use_bc_loss"):
        S = w["ddim_steps"] if w["use_ddim"] else w["denoising_steps"]
        bc_noise = torch.randn((S + 1, spec["mb_rows"], w["horizon_steps"], w["action_dim"]),
                               generator=torch.Generator().manual_seed(int(gold["bc_noise_seed"])))
    res = O.ppo_loss(p, nc, dc, inp["state"][b], chains[b, d], chains[b, d + 1], d, inp["returns"][b], inp["oldvalues"][b],
                     inp["advantages"][b], lp_k[b, d] + inp["lp_shift"], reward_horizon=w["act_steps"],
                     use_bc_loss=bool(spec.get("use_bc_loss")), bc_noise=bc_noise)
    (res[0] + 0.5 * res[2] + spec.get("bc_coeff", 0.0) * res[6]).backward()
    got = np.array([float(res[0]), float(res[1]), float(res[2]), res[3], res[4], res[5], float(res[6]), res[7]])
    np.testing.assert_allclose(got, gold["loss_scalars"], rtol=1e-6, atol=1e-9)
    for gname, (norm, total) in zip(gold["grad_names"], gold["grad_stats"]):
        key = str(gname) if str(gname).startswith("critic.") else "actor_ft." + str(gname)
        g = p[key].grad
        assert abs(float(g.double().norm()) - norm) <= 1e-4 * max(norm, 1e-12), key


def test_reward_scaler_matches_reference_vectors():
    """oracle and host mirror of RunningRewardScaler against three consecutive calls of the reference class."""
    from dppo_b200.util.reward_scaling import RunningRewardScaler
    from tests.helpers import GOLDEN_DIR

    gold = dict(np.load(f"{GOLDEN_DIR}/reward_scaler.npz"))
    E = gold["reward0"].shape[0]
    for cls in (O.RunningRewardScaler, RunningRewardScaler):
        sc = cls(E)
        for k in range(3):
            out = sc(reward=gold[f"reward{k}"], first=gold[f"first{k}"])
            np.testing.assert_array_equal(out, gold[f"scaled{k}"])
            state = np.concatenate([sc.ret, [sc.mean, sc.var, sc.count]])
            np.testing.assert_array_equal(state, gold[f"state{k}"])


def test_lr_schedule_matches_reference_scheduler():
    """cosine_warmup_lr(n) == the optimiser's lr after n CosineAnnealingWarmupRestarts.step() calls (n = 0: after
    construction), the mapping the agent uses (one scheduler step per iteration, actor only after the critic warm-up)."""
    from dppo_b200.agent.finetune.train_ppo_diffusion_agent import cosine_warmup_lr
    from tests.helpers import GOLDEN_DIR

    gold = dict(np.load(f"{GOLDEN_DIR}/lr_schedule.npz"))
    for tag in "abcd":
        first, mx, mn, warm = gold[f"cfg_{tag}"]
        got = [cosine_warmup_lr(n, int(first), float(mx), float(mn), int(warm)) for n in range(len(gold[f"lr_{tag}"]))]
        np.testing.assert_allclose(got, gold[f"lr_{tag}"], rtol=1e-12, atol=0)
