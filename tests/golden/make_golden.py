"""
Golden-vector generator (run in the build container only; needs /root/reference, which does not travel to the GPU box).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
    python tests/golden/make_golden.py kitchen    # only the named cases (the others keep their recorded bytes)

For each workload it instantiates the UNMODIFIED reference classes (dppo.model.diffusion.diffusion_ppo.PPODiffusion
with DiffusionMLP / Unet1D / CriticObs / EtaFixed) on CPU in fp32, with the seeded weight recipe of
tests/helpers.py, injects a pre-drawn noise tensor by swapping the name `torch` inside
dppo.model.diffusion.diffusion_vpg for a proxy whose randn / randn_like pop slices of it, and records

  forward(return_chain=True)  (train and deterministic), get_logprobs, loss scalars + actor_ft / critic gradients.

The .npz files also hold float64 checksums of every parameter so the tests can prove that the weights they rebuild
from the seed are the ones the reference used.  Recorded with torch.__version__ (see `meta`).
"""

import os
import sys
import types

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

from dppo_b200.workloads import chain_evals, get_workload  # noqa: E402
from tests.helpers import GOLDEN_CASES, build_model, make_inputs, param_checksums  # noqa: E402


class _NoiseProxy(types.ModuleType):
    """Stands in for `torch` inside diffusion_vpg: randn / randn_like replay a pre-drawn tensor, everything else passes through."""

    def __init__(self, noise):
        super().__init__("torch")
        self._noise = noise
        self._k = 0

    def __getattr__(self, name):
        return getattr(torch, name)

    def randn(self, *a, **k):
        out = self._noise[self._k].clone()
        self._k += 1
        return out

    def randn_like(self, x, **k):
        out = self._noise[self._k].clone()
        self._k += 1
        return out


def run_forward(model, state, noise, deterministic):
    ref_vpg.torch = _NoiseProxy(noise)
    try:
        out = model(cond={"state": state}, deterministic=deterministic, return_chain=True)
    finally:
        ref_vpg.torch = torch
    return out.trajectories, out.chains


def main():
    torch.set_num_threads(8)
    only = set(sys.argv[1:])
    for case, spec in GOLDEN_CASES.items():
        if only and case not in only:
            continue
        w = get_workload(spec["workload"])
        E = spec["n_envs"]
        model = build_model(
            w, "cpu",
            classes=dict(ppo=PPODiffusion, mlp=DiffusionMLP, unet=Unet1D, critic=CriticObs, eta=EtaFixed),
        )
        S = chain_evals(w)
        ft = w["ft_denoising_steps"]
        inp = make_inputs(w, E, spec["mb_rows"])
        state, noise = inp["state"], inp["noise"]
        out = {"meta": np.array(f"torch {torch.__version__} cpu fp32; workload {spec['workload']}; E={E}; S={S}; ft={ft}")}
        model.train()
        traj, chains = run_forward(model, state, noise, deterministic=False)
        out["traj"], out["chains"] = traj.numpy(), chains.numpy()
        model.eval()
        traj_d, chains_d = run_forward(model, state, noise, deterministic=True)
        out["traj_det"], out["chains_det"] = traj_d.numpy(), chains_d.numpy()
        model.train()
        with torch.no_grad():
            lp = model.get_logprobs({"state": state}, chains)
            out["logprobs"] = lp.numpy()
            out["values"] = model.critic({"state": state}).numpy()
        # one PPO minibatch drawn from the (E, ft) rows with the seeded (b, d) indices of make_inputs
        b, d = inp["mb_b"], inp["mb_d"]
        lp_k = lp.reshape(E, ft, w["horizon_steps"], w["action_dim"])
        for p in model.parameters():
            p.grad = None
        res = model.loss(
            {"state": state[b]}, chains[b, d], chains[b, d + 1], d, inp["returns"][b], inp["oldvalues"][b],
            inp["advantages"][b], lp_k[b, d] + inp["lp_shift"], use_bc_loss=False, reward_horizon=w["act_steps"],
        )
        pg, ent, vl = res[0], res[1], res[2]
        (pg + 0.5 * vl).backward()
        out["loss_scalars"] = np.array([float(pg), float(ent), float(vl), res[3], res[4], res[5], float(res[6]), res[7]], dtype=np.float64)
        gnames, gnorm = [], []
        for name, p in list(model.actor_ft.named_parameters()) + [("critic." + n, q) for n, q in model.critic.named_parameters()]:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            gnames.append(name)
            gnorm.append([float(g.double().norm()), float(g.double().sum())])
        out["grad_names"] = np.array(gnames)
        out["grad_stats"] = np.array(gnorm)
        # full gradient of the output layer (small) as an element-wise pin
        last_w = [n for n, _ in model.actor_ft.named_parameters() if n.endswith("weight")][-1]
        out["grad_last_name"] = np.array(last_w)
        out["grad_last"] = dict(model.actor_ft.named_parameters())[last_w].grad.numpy()
        names, sums = param_checksums(model)
        out["param_names"], out["param_sums"] = np.array(names), np.array(sums)
        path = os.path.join(HERE, f"{case}.npz")
        np.savez_compressed(path, **out)
        print(case, "->", path, os.path.getsize(path) // 1024, "KiB", "| loss", out["loss_scalars"][:3])


if __name__ == "__main__":
    main()
