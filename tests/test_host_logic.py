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
