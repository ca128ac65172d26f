"""
Agent-level golden vectors (run in the build container only; needs /root/reference):

    python tests/golden/make_golden_agent.py        # rewrites tests/golden/agent_run.npz

Runs the UNMODIFIED reference loop, dppo.agent.finetune.train_ppo_diffusion_agent.TrainPPODiffusionAgent.run
(reference :47-483), for a few iterations on CPU under a sys.modules shim (omegaconf / hydra stand-ins,
`env.gym_utils.make_async` returning the synthetic vector env) with

  * every torch.randn / randn_like inside diffusion_vpg replaced by draws from ONE seeded CPU generator (the B200 agent
    replays the same sequence through its `test_hooks["noise"]`),
  * every torch.randperm of the minibatch loop recorded (the B200 agent replays them through `test_hooks["perm"]`: a CPU
    permutation stream cannot be reproduced by the CUDA generator).

Recorded per minibatch: what PPODiffusion.loss returned (pg_loss, v_loss, clipfrac, approx_kl, ratio) and float64 sums of
its inputs (returns, values, advantages, old log-probs = the GAE bootstrap on the post-rollout observation (:259-278), the
tail-row drop (:312), reward scaling (:243-247) and the (b, d) indexing (:316-327) all feed these); per iteration the
learning rates (critic warm-up :365,407-411) and the number of minibatches (KL early stop :379); at the end float64
checksums of every actor_ft / critic parameter.
"""

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from dppo_b200.env.synthetic import SyntheticVecEnv  # noqa: E402
from dppo_b200.util.config import Cfg, instantiate  # noqa: E402
from tests.helpers import agent_golden_cfg  # noqa: E402

NOISE_SEED = 2024


def install_shims(cfg):
    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = type("OmegaConf", (), {"to_container": staticmethod(lambda c, resolve=True: dict(c))})
    sys.modules["omegaconf"] = oc
    hy = types.ModuleType("hydra")
    hy.utils = types.ModuleType("hydra.utils")
    hy.utils.instantiate = instantiate
    sys.modules["hydra"], sys.modules["hydra.utils"] = hy, hy.utils
    env = types.ModuleType("env")
    gu = types.ModuleType("env.gym_utils")

    def make_async(name, num_envs=None, max_episode_steps=None, obs_dim=None, action_dim=None, **kw):
        return SyntheticVecEnv(num_envs, obs_dim, action_dim, cfg.cond_steps, cfg.act_steps, max_episode_steps, seed=cfg.seed)

    gu.make_async = make_async
    env.gym_utils = gu
    sys.modules["env"], sys.modules["env.gym_utils"] = env, gu


class _Proxy(types.ModuleType):
    def __init__(self, **over):
        super().__init__("torch")
        self._over = over

    def __getattr__(self, name):
        over = object.__getattribute__(self, "_over")
        return over[name] if name in over else getattr(torch, name)


def main():
    cfg = agent_golden_cfg("cpu", tempfile.mkdtemp(), reference=True)
    install_shims(cfg)
    from dppo.agent.finetune import train_ppo_diffusion_agent as ref_agent
    from dppo.model.diffusion import diffusion_vpg as ref_vpg

    gen = torch.Generator().manual_seed(NOISE_SEED)
    ref_vpg.torch = _Proxy(randn=lambda *s, **k: torch.randn(*[d for d in (s[0] if isinstance(s[0], (tuple, list, torch.Size)) else s)], generator=gen),
                           randn_like=lambda x, **k: torch.randn(x.shape, generator=gen))
    perms = []

    def randperm(n, **k):
        p = torch.randperm(n)
        perms.append(p.clone())
        return p

    ref_agent.torch = _Proxy(randperm=randperm)
    torch.set_num_threads(8)
    ag = ref_agent.TrainPPODiffusionAgent(cfg)
    # same weight recipe as tests/helpers.build_model: actor_ft perturbed so that ft != base
    from tests.helpers import PERTURB_SCALE, PERTURB_SEED

    g = torch.Generator().manual_seed(PERTURB_SEED)
    with torch.no_grad():
        for p in ag.model.actor_ft.parameters():
            p.add_(PERTURB_SCALE * torch.randn(p.shape, generator=g))
    mb, lrs = [], []
    loss0 = ag.model.loss

    def loss(obs, chains_prev, chains_next, dinds, returns, values, adv, lp, **kw):
        out = loss0(obs, chains_prev, chains_next, dinds, returns, values, adv, lp, **kw)
        mb.append([float(out[0]), float(out[2]), out[3], out[4], out[5], float(returns.double().sum()), float(values.double().sum()),
                   float(adv.double().sum()), float(lp.double().sum()), float(dinds.double().sum()), ag.itr,
                   ag.actor_optimizer.param_groups[0]["lr"], ag.critic_optimizer.param_groups[0]["lr"]])
        return out

    ag.model.loss = loss
    res = ag.run()
    ref_vpg.torch = torch
    ref_agent.torch = torch
    mb = np.array(mb, dtype=np.float64)
    names, sums = [], []
    for k, v in ag.model.state_dict().items():
        if k.startswith("actor_ft.") or k.startswith("critic."):
            names.append(k)
            sums.append([float(v.double().sum()), float(v.double().pow(2).sum())])
    out = dict(
        meta=np.array(f"torch {torch.__version__} cpu fp32; reference TrainPPODiffusionAgent.run, {cfg.train.n_train_itr} iterations"),
        minibatch=mb, perms=torch.stack(perms).numpy().astype(np.int32), noise_seed=np.array(NOISE_SEED),
        param_names=np.array(names), param_sums=np.array(sums),
        critic_out_bias=ag.model.critic.state_dict()["Q1.layers.2.bias"].numpy(),
        actor_out_bias=ag.model.actor_ft.state_dict()["mlp_mean.layers.2.bias"].numpy(),
    )
    path = os.path.join(HERE, "agent_run.npz")
    np.savez_compressed(path, **out)
    print("agent_run ->", os.path.getsize(path) // 1024, "KiB;", len(mb), "minibatches; per iteration:",
          [int((mb[:, 10] == i).sum()) for i in range(cfg.train.n_train_itr)])
    print("approx_kl per minibatch:", np.array2string(mb[:, 3], precision=3))
    print("lrs:", sorted(set(map(tuple, mb[:, 11:13]))))


if __name__ == "__main__":
    main()
