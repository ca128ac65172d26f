"""
Golden vectors for the OPTIONAL branches of the hot path and for the host-side pieces around it (run in the build
container only; needs /root/reference):

    python tests/golden/make_golden_variants.py      # rewrites tests/golden/variant_*.npz, reward_scaler.npz, lr_schedule.npz

  variant_bc          use_bc_loss=True                      reference diffusion_ppo.py:105-126
  variant_vclip_quant clip_vloss_coef=0.2, advantage quantiles 0.05 / 0.95   diffusion_ppo.py:133-135,178-187
  variant_epsclip     eps_clip_value=0.5 (DDIM)             diffusion_vpg.py:194-195
  variant_finalclip   final_action_clip_value=0.5           diffusion_vpg.py:300-303
  variant_anneal      ft_denoising_steps_d=3, _t=1, one model.step()   diffusion_vpg.py:102-127
  reward_scaler       RunningRewardScaler over three calls  dppo/util/reward_scaling.py:42-87
  lr_schedule         CosineAnnealingWarmupRestarts         dppo/util/scheduler.py (agent usage train_ppo_agent.py:39-65,
                                                            train_ppo_diffusion_agent.py:406-411)
Every vector comes from the UNMODIFIED reference classes on CPU fp32 with the seeded recipe of tests/helpers.py and
injected noise (the `torch` proxy of make_golden.py).
"""

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from dppo.model.common.critic import CriticObs  # noqa: E402
from dppo.model.diffusion import diffusion_vpg as ref_vpg  # noqa: E402
from dppo.model.diffusion.diffusion_ppo import PPODiffusion  # noqa: E402
from dppo.model.diffusion.eta import EtaFixed  # noqa: E402
from dppo.model.diffusion.mlp_diffusion import DiffusionMLP  # noqa: E402
from dppo.model.diffusion.unet import Unet1D  # noqa: E402
from dppo.util.reward_scaling import RunningRewardScaler  # noqa: E402
from dppo.util.scheduler import CosineAnnealingWarmupRestarts  # noqa: E402

from dppo_b200.workloads import get_workload  # noqa: E402
from tests.golden.make_golden import _NoiseProxy, run_forward  # noqa: E402
from tests.helpers import VARIANTS, build_model, make_inputs, perturb_again, variant_workload  # noqa: E402

REF = dict(ppo=PPODiffusion, mlp=DiffusionMLP, unet=Unet1D, critic=CriticObs, eta=EtaFixed)


def grad_stats(model):
    names, stats = [], []
    for name, p in list(model.actor_ft.named_parameters()) + [("critic." + n, q) for n, q in model.critic.named_parameters()]:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        names.append(name)
        stats.append([float(g.double().norm()), float(g.double().sum())])
    return np.array(names), np.array(stats)


def loss_record(model, w, inp, E, chains, lp, out, use_bc_loss=False, bc_noise=None, bc_coeff=0.0):
    ft = w["ft_denoising_steps"]
    b, d = inp["mb_b"], inp["mb_d"]
    lp_k = lp.reshape(E, ft, w["horizon_steps"], w["action_dim"])
    for p in model.parameters():
        p.grad = None
    if use_bc_loss:
        ref_vpg.torch = _NoiseProxy(bc_noise)
    try:
        res = model.loss({"state": inp["state"][b]}, chains[b, d], chains[b, d + 1], d, inp["returns"][b], inp["oldvalues"][b],
                         inp["advantages"][b], lp_k[b, d] + inp["lp_shift"], use_bc_loss=use_bc_loss,
                         reward_horizon=w["act_steps"])
    finally:
        ref_vpg.torch = torch
    (res[0] + 0.5 * res[2] + bc_coeff * res[6]).backward()
    out["loss_scalars"] = np.array([float(res[0]), float(res[1]), float(res[2]), res[3], res[4], res[5], float(res[6]), res[7]],
                                   dtype=np.float64)
    out["grad_names"], out["grad_stats"] = grad_stats(model)
    last_w = [n for n, _ in model.actor_ft.named_parameters() if n.endswith("weight")][-1]
    out["grad_last_name"] = np.array(last_w)
    out["grad_last"] = dict(model.actor_ft.named_parameters())[last_w].grad.numpy().copy()


def main():
    torch.set_num_threads(8)
    for name, spec in VARIANTS.items():
        w = variant_workload(name)
        E = spec["n_envs"]
        model = build_model(w, "cpu", REF)
        inp = make_inputs(w, E, spec["mb_rows"], seed=spec.get("seed", 0))
        out = {"meta": np.array(f"torch {torch.__version__} cpu fp32; variant {name}")}
        if spec.get("anneal"):
            model.step()  # ft 10 -> 7, actor <- actor_ft, fresh actor_ft (reference diffusion_vpg.py:102-127)
            perturb_again(model)
            inp = make_inputs(w, E, spec["mb_rows"], seed=spec.get("seed", 0), ft=model.ft_denoising_steps)
            out["ft_after"] = np.array(model.ft_denoising_steps)
        model.train()
        traj, chains = run_forward(model, inp["state"], inp["noise"], deterministic=False)
        out["traj"], out["chains"] = traj.numpy(), chains.numpy()
        with torch.no_grad():
            lp = model.get_logprobs({"state": inp["state"]}, chains)
        out["logprobs"] = lp.numpy()
        if spec.get("loss", True) and not spec.get("anneal"):
            bc_noise = None
            if spec.get("use_bc_loss"):
                g = torch.Generator().manual_seed(99)
                S = w["ddim_steps"] if w["use_ddim"] else w["denoising_steps"]
                bc_noise = torch.randn((S + 1, spec["mb_rows"], w["horizon_steps"], w["action_dim"]), generator=g)
                out["bc_noise_seed"] = np.array(99)
            loss_record(model, w, inp, E, chains, lp, out, use_bc_loss=bool(spec.get("use_bc_loss")), bc_noise=bc_noise,
                        bc_coeff=spec.get("bc_coeff", 0.0))
        path = os.path.join(HERE, f"variant_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", os.path.getsize(path) // 1024, "KiB", out.get("loss_scalars"))

    # ---- RunningRewardScaler: three consecutive iterations of (E, n_steps) rewards
    rng = np.random.default_rng(21)
    E, n = 7, 13
    sc = RunningRewardScaler(E)
    rec = {}
    for k in range(3):
        reward = rng.standard_normal((E, n)) * (1.0 + k)
        first = (rng.random((E, n)) < 0.15).astype(np.float64)
        rec[f"reward{k}"], rec[f"first{k}"] = reward, first
        rec[f"scaled{k}"] = sc(reward=reward, first=first)
        rec[f"state{k}"] = np.concatenate([sc.ret, [sc.ret_rms.mean, sc.ret_rms.var, sc.ret_rms.count]])
    np.savez_compressed(os.path.join(HERE, "reward_scaler.npz"), **rec)
    print("reward_scaler ->", rec["state2"][-3:])

    # ---- LR schedules as the agent drives them: the optimiser's lr after construction and after each scheduler.step()
    rec = {}
    for tag, kw in {"a": dict(first_cycle_steps=10, max_lr=1e-3, min_lr=1e-4, warmup_steps=2),
                    "b": dict(first_cycle_steps=100, max_lr=1e-5, min_lr=1e-6, warmup_steps=10),
                    "c": dict(first_cycle_steps=7, max_lr=3e-4, min_lr=3e-4, warmup_steps=0),
                    "d": dict(first_cycle_steps=6, max_lr=1e-3, min_lr=1e-5, warmup_steps=0)}.items():
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=kw["max_lr"])
        sched = CosineAnnealingWarmupRestarts(opt, cycle_mult=1.0, gamma=1.0, **kw)
        lrs = [opt.param_groups[0]["lr"]]
        for _ in range(25):
            sched.step()
            lrs.append(opt.param_groups[0]["lr"])
        rec[f"lr_{tag}"] = np.array(lrs)
        rec[f"cfg_{tag}"] = np.array([kw["first_cycle_steps"], kw["max_lr"], kw["min_lr"], kw["warmup_steps"]])
    np.savez_compressed(os.path.join(HERE, "lr_schedule.npz"), **rec)
    print("lr_schedule ->", rec["lr_a"][:5])


if __name__ == "__main__":
    main()
