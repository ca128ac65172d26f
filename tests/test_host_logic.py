"""Host-side pieces of the agent loop (CPU): reward scaler vs the oracle, env stub contract, LR schedule, config."""

import math

import numpy as np

from dppo_b200.agent.finetune.train_ppo_diffusion_agent import cosine_warmup_lr
from dppo_b200.env.synthetic import SyntheticVecEnv
from dppo_b200.util.config import Cfg, instantiate
from dppo_b200.util.reward_scaling import RunningRewardScaler
from oracle import dppo_oracle as O


def test_running_reward_scaler_matches_oracle_over_iterations():
    rng = np.random.default_rng(3)
    ours, ref = RunningRewardScaler(6), O.RunningRewardScaler(6)
    for _ in range(4):
        reward = rng.standard_normal((6, 11)) * 3.0
        first = (rng.random((6, 11)) < 0.2).astype(np.float64)
        np.testing.assert_array_equal(ours(reward.copy(), first), ref(reward.copy(), first))
    s = ours.state_dict()
    again = RunningRewardScaler(6)
    again.load_state_dict(s)
    r, f = rng.standard_normal((6, 5)), np.zeros((6, 5))
    np.testing.assert_array_equal(again(r.copy(), f), ours(r.copy(), f))


def test_synthetic_env_contract_and_sharding():
    env = SyntheticVecEnv(5, obs_dim=11, action_dim=3, cond_steps=1, act_steps=4, max_episode_steps=12, seed=42)
    obs = env.reset_arg([{} for _ in range(5)])
    assert obs["state"].shape == (5, 1, 11) and obs["state"].dtype == np.float32 and np.abs(obs["state"]).max() <= 1
    trunc_seen = False
    for _ in range(3):
        o, r, term, trunc, info = env.step(np.zeros((5, 4, 3), dtype=np.float32))
        assert o["state"].shape == (5, 1, 11) and r.shape == (5,) and term.dtype == bool and trunc.dtype == bool and len(info) == 5
        trunc_seen |= bool(trunc.any())
    assert trunc_seen  # 12 // 4 = 3 decisions per episode
    # env-sharded ranks draw what the single process draws for the same envs
    whole = SyntheticVecEnv(5, 11, 3, seed=42).reset_arg()["state"]
    part = SyntheticVecEnv(2, 11, 3, seed=42, env_offset=3).reset_arg()["state"]
    np.testing.assert_array_equal(whole[3:], part)


def test_cosine_warmup_schedule_values():
    kw = dict(first_cycle_steps=10, max_lr=1e-3, min_lr=1e-4, warmup_steps=2)
    assert cosine_warmup_lr(-1, **kw) == 1e-4
    assert cosine_warmup_lr(0, **kw) == 1e-4
    assert math.isclose(cosine_warmup_lr(1, **kw), 1e-4 + 9e-4 / 2)
    assert math.isclose(cosine_warmup_lr(2, **kw), 1e-3)
    assert math.isclose(cosine_warmup_lr(6, **kw), 1e-4 + 9e-4 * (1 + math.cos(math.pi * 4 / 8)) / 2)
    assert math.isclose(cosine_warmup_lr(10, **kw), cosine_warmup_lr(0, **kw))  # restart


def test_cfg_and_instantiate():
    cfg = Cfg(a=1, b=dict(c=2, d=dict(e=3)))
    assert cfg.b.d.e == 3 and cfg.get("zz", 7) == 7 and "a" in cfg
    obj = instantiate({"_target_": "dppo_b200.util.reward_scaling.RunningRewardScaler", "num_envs": 3})
    assert obj.ret.shape == (3,)


def test_dense_conv_lowering_equals_convolutions_on_cpu():
    """The index / mask / inverse maps of dense_conv._lowering reproduce Conv1d / ConvTranspose1d (forward and dW)."""
    import numpy as np
    import torch
    import torch.nn.functional as F

    from dppo_b200.model.diffusion.dense_conv import _GatherDense, _lowering

    torch.manual_seed(0)
    for kind, cin, cout, ks, stride, pad, T in [("conv", 7, 16, 5, 1, 2, 4), ("conv", 16, 16, 3, 2, 1, 4), ("conv", 8, 4, 1, 1, 0, 2),
                                                ("convT", 16, 16, 4, 2, 1, 2), ("conv", 6, 5, 5, 1, 2, 8)]:
        w = torch.randn((cout, cin, ks) if kind == "conv" else (cin, cout, ks), dtype=torch.float64, requires_grad=True)
        x = torch.randn(9, cin, T, dtype=torch.float64)
        t_out, idx, mask, inv = _lowering(kind, cin, cout, ks, stride, pad, T)
        dense = _GatherDense.apply(w, torch.from_numpy(idx), torch.from_numpy(mask).double(), torch.from_numpy(inv))
        y = (x.reshape(9, -1) @ dense.t()).view(9, cout, t_out)
        ref = F.conv1d(x, w, None, stride, pad) if kind == "conv" else F.conv_transpose1d(x, w, None, stride, pad)
        assert ref.shape == y.shape
        np.testing.assert_allclose(y.detach().numpy(), ref.detach().numpy(), rtol=1e-12, atol=1e-12)
        g = torch.randn_like(ref)
        (gw,) = torch.autograd.grad(y, w, g)
        (gw_ref,) = torch.autograd.grad(ref, w, g)
        np.testing.assert_allclose(gw.numpy(), gw_ref.numpy(), rtol=1e-12, atol=1e-12)


def test_fast_group_norm_formula_and_split_linear_cpu_path():
    import numpy as np
    import torch
    import torch.nn.functional as F

    from dppo_b200.model.common.split_linear import SplitLinear
    from dppo_b200.model.diffusion.modules import FastGroupNorm

    torch.manual_seed(1)
    gn = FastGroupNorm(8, 64, eps=1e-5).double()
    with torch.no_grad():
        gn.weight.normal_()
        gn.bias.normal_()
    for shape in [(5, 64, 1, 4), (3, 64, 2)]:
        x = torch.randn(shape, dtype=torch.float64)
        np.testing.assert_allclose(gn.via_layer_norm(x).detach().numpy(),
                                   F.group_norm(x, 8, gn.weight, gn.bias, 1e-5).detach().numpy(), rtol=1e-10, atol=1e-12)
    lin = SplitLinear(11, 7)
    x = torch.randn(4, 11)
    assert torch.equal(lin(x), F.linear(x, lin.weight, lin.bias))  # CPU tensors: the stock fp32 path
    assert set(lin.state_dict()) == {"weight", "bias"}


def test_flat_grad_buffer_pads_every_tensor_to_16_bytes():
    import torch

    from dppo_b200 import distributed as D

    a = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7))]
    b = [torch.nn.Parameter(torch.zeros(2, 2))]
    buf = D.FlatGradBuffer([a, b], n_scalars=8)
    assert buf.flat.numel() == 16 + 8 + 4 + 8
    offs = [(p.grad.data_ptr() - buf.flat.data_ptr()) // 4 for p in a + b]
    assert offs == [0, 16, 24] and all(o % 4 == 0 for o in offs)
    a[1].grad.fill_(2.0)
    assert float(buf.flat[16:23].sum()) == 14.0 and float(buf.flat[23]) == 0.0


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (written by bench.py on the GPU box) have every key the contract names."""
    import glob
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = sorted(glob.glob(os.path.join(root, "profiles", "r1*_bench*.json")) +
                   glob.glob(os.path.join(root, "profiles", "r2", "r2y_bench*.json")))  # + the final tree of round 2
    assert lines, "no bench lines under profiles/"
    for path in lines:
        d = json.load(open(path))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"):
            assert k in d, (path, k)
        assert d["config"].get("workload") and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        if d.get("impl") == "reference":
            assert d["gpu_launches"] == 0 and d["e2e"]["h2d_bytes_per_step"] == 0
            continue
        assert d["gpu_launches"] == d["steps"] and d["steps"] > 0 and d["warmup"] >= 3
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and d["clocks"]["samples"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["e2e"]["value"] < d["value"]  # the end-to-end number is never the device-timed one


def test_update_route_per_reference_yaml_geometry():
    """Which route the PPO update takes for every DiffusionMLP / Unet1D geometry among the reference's state-based
    fine-tuning YAMLs: the tcgen05 program where dppo_update_create's limits hold, else the autograd route WITH a reason
    (kitchen's 32-wide cond_mlp output once reached dppo_update_create and failed the first minibatch)."""
    from dppo_b200.update_engine import unsupported_reason
    from dppo_b200.workloads import get_workload
    from tests.helpers import build_model, our_classes

    want = {"hopper": None, "walker2d": None, "avoid": None, "transport": None, "furniture": None, "square_mlp": None,
            "kitchen": "cond_mlp output width 32 is not a multiple of 64", "square_unet": "actor is not a DiffusionMLP"}
    for name, reason in want.items():
        model = build_model(get_workload(name), "cpu", our_classes(), perturb=False)
        assert unsupported_reason(model) == reason, name
        assert model.fused_update_reason() == reason, name
