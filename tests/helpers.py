"""
Shared test utilities: the seeded weight / input recipe used by the golden generator (with the reference classes), by
the oracle tests (parameters rebuilt from the seed through dppo_b200's host-side containers) and by the GPU tests.
"""

import os

import numpy as np
import torch

from dppo_b200.workloads import chain_evals, get_workload

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# case name -> workload + reduced sizes so the fixtures stay small
GOLDEN_CASES = {
    "hopper": dict(workload="hopper", n_envs=40, mb_rows=160),
    "walker2d": dict(workload="walker2d", n_envs=96, mb_rows=256),
    "transport_k20": dict(workload="transport_k20", n_envs=50, mb_rows=128),
    "transport": dict(workload="transport", n_envs=24, mb_rows=96),
    "furniture": dict(workload="furniture", n_envs=40, mb_rows=96),
    "furniture_ddpm100": dict(workload="furniture_ddpm100", n_envs=16, mb_rows=48),
    "square_unet": dict(workload="square_unet", n_envs=40, mb_rows=128),
    # the other DiffusionMLP geometries of the reference's fine-tuning YAMLs (H = 256 with a 128 / 32 cond_mlp; a 2-D action
    # with a 4-D observation; H = 1024 with cond_mlp and no LayerNorm)
    "kitchen": dict(workload="kitchen", n_envs=40, mb_rows=96),
    "avoid": dict(workload="avoid", n_envs=50, mb_rows=96),
    "square_mlp": dict(workload="square_mlp", n_envs=24, mb_rows=64),
    # Unet1D with dim 40 (robomimic can / lift): GroupNorm groups of 20 lowered features (segmented path of chain_unet.cu)
    "can_unet": dict(workload="can_unet", n_envs=50, mb_rows=96),
}

# optional branches of the hot path (tests/golden/make_golden_variants.py): base workload + constructor overrides
VARIANTS = {
    "bc": dict(base="hopper", ppo={}, n_envs=24, mb_rows=64, use_bc_loss=True, bc_coeff=0.1),
    "vclip_quant": dict(base="hopper", n_envs=24, mb_rows=96,
                        ppo=dict(clip_vloss_coef=0.2, clip_advantage_lower_quantile=0.05, clip_advantage_upper_quantile=0.95)),
    "epsclip": dict(base="furniture", ppo=dict(eps_clip_value=0.3), n_envs=16, mb_rows=48),
    "finalclip": dict(base="hopper", ppo=dict(final_action_clip_value=0.5), n_envs=24, mb_rows=64, loss=False),
    "anneal": dict(base="hopper", ppo=dict(ft_denoising_steps_d=3, ft_denoising_steps_t=1), n_envs=24, mb_rows=32, anneal=True),
}

WEIGHT_SEED = 42
PERTURB_SEED = 1234
PERTURB_SCALE = 1e-2


def build_model(w, device, classes, perturb=True):
    """
    Seeded construction shared by reference and dppo_b200 classes: actor, critic, eta are created in this order under
    torch.manual_seed(WEIGHT_SEED) (the YAML seed), then actor_ft is perturbed so that ft != base.
    `classes`: dict(ppo=, mlp=, unet=, critic=, eta=).
    """
    torch.manual_seed(WEIGHT_SEED)
    a = dict(w["actor"])
    kind = a.pop("kind")
    cond_dim = w["obs_dim"] * w["cond_steps"]
    if kind == "mlp":
        actor = classes["mlp"](action_dim=w["action_dim"], horizon_steps=w["horizon_steps"], cond_dim=cond_dim, **a)
    else:
        actor = classes["unet"](action_dim=w["action_dim"], cond_dim=cond_dim, **a)
    critic = classes["critic"](cond_dim=cond_dim, **w["critic"])
    eta = classes["eta"](**w["eta"]) if w.get("eta") else None
    model = classes["ppo"](
        actor=actor, critic=critic, eta=eta, learn_eta=False,
        ft_denoising_steps=w["ft_denoising_steps"], horizon_steps=w["horizon_steps"], obs_dim=w["obs_dim"],
        action_dim=w["action_dim"], denoising_steps=w["denoising_steps"], device=device,
        use_ddim=w["use_ddim"], ddim_steps=w["ddim_steps"], network_path=None, **w["ppo"],
    )
    if perturb:
        g = torch.Generator().manual_seed(PERTURB_SEED)
        with torch.no_grad():
            for p in model.actor_ft.parameters():
                p.add_((PERTURB_SCALE * torch.randn(p.shape, generator=g)).to(p.device))
    return model


def variant_workload(name):
    spec = VARIANTS[name]
    w = get_workload(spec["base"])
    w["ppo"] = dict(w["ppo"], **spec["ppo"])
    return w


def perturb_again(model, seed=4321):
    """Second perturbation of actor_ft (after VPGDiffusion.step() made it a copy of the new base policy)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.actor_ft.parameters():
            p.add_((PERTURB_SCALE * torch.randn(p.shape, generator=g)).to(p.device))


def make_inputs(w, n_envs, mb_rows, seed=0, ft=None):
    """Seeded synthetic observations U(-1,1), injected noise (S+1,E,Ta,Da), and one PPO minibatch worth of scalars."""
    rng = np.random.default_rng(seed)
    S, ft = chain_evals(w), (w["ft_denoising_steps"] if ft is None else ft)
    state = torch.from_numpy(rng.uniform(-1, 1, (n_envs, w["cond_steps"], w["obs_dim"])).astype(np.float32))
    noise = torch.from_numpy(rng.standard_normal((S + 1, n_envs, w["horizon_steps"], w["action_dim"])).astype(np.float32))
    noise[1:] *= 1.5  # make the +-randn_clip clamp bite on some elements
    mb_b = torch.from_numpy(rng.integers(0, n_envs, mb_rows))
    mb_d = torch.from_numpy(rng.integers(0, ft, mb_rows))
    return dict(
        state=state, noise=noise, mb_b=mb_b, mb_d=mb_d,
        returns=torch.from_numpy(rng.standard_normal(n_envs).astype(np.float32)),
        oldvalues=torch.from_numpy(rng.standard_normal(n_envs).astype(np.float32)),
        advantages=torch.from_numpy(rng.standard_normal(n_envs).astype(np.float32)),
        lp_shift=torch.from_numpy((0.02 * rng.standard_normal((mb_rows, w["horizon_steps"], w["action_dim"]))).astype(np.float32)),
    )


def param_checksums(model):
    names, sums = [], []
    for k, v in model.state_dict().items():
        names.append(k)
        sums.append([float(v.double().sum()), float(v.double().pow(2).sum())])
    return names, sums


def load_golden(case):
    path = os.path.join(GOLDEN_DIR, f"{case}.npz")
    return dict(np.load(path, allow_pickle=False))


def our_classes():
    from dppo_b200.model.common.critic import CriticObs
    from dppo_b200.model.diffusion.diffusion_ppo import PPODiffusion
    from dppo_b200.model.diffusion.eta import EtaFixed
    from dppo_b200.model.diffusion.mlp_diffusion import DiffusionMLP
    from dppo_b200.model.diffusion.unet import Unet1D

    return dict(ppo=PPODiffusion, mlp=DiffusionMLP, unet=Unet1D, critic=CriticObs, eta=EtaFixed)


def oracle_cfgs(w):
    """NetCfg / DiffCfg of oracle/dppo_oracle.py for a workload dict."""
    from oracle.dppo_oracle import DiffCfg, NetCfg

    a = w["actor"]
    if a["kind"] == "mlp":
        nc = NetCfg(
            kind="mlp", obs_dim=w["obs_dim"], cond_steps=w["cond_steps"], action_dim=w["action_dim"],
            horizon_steps=w["horizon_steps"], time_dim=a["time_dim"], mlp_dims=list(a["mlp_dims"]),
            cond_mlp_dims=a.get("cond_mlp_dims"), activation=a.get("activation_type", "Mish"),
            use_layernorm=a.get("use_layernorm", False), critic_dims=list(w["critic"]["mlp_dims"]),
            critic_activation=w["critic"]["activation_type"],
        )
    else:
        nc = NetCfg(
            kind="unet", obs_dim=w["obs_dim"], cond_steps=w["cond_steps"], action_dim=w["action_dim"],
            horizon_steps=w["horizon_steps"], time_dim=a["diffusion_step_embed_dim"], activation="Mish",
            critic_dims=list(w["critic"]["mlp_dims"]), critic_activation=w["critic"]["activation_type"],
            unet_dim=a["dim"], unet_mults=tuple(a["dim_mults"]), unet_kernel=a["kernel_size"],
            unet_groups=a["n_groups"], unet_cond_predict_scale=a["cond_predict_scale"],
            unet_smaller_encoder=a["smaller_encoder"],
        )
    p = w["ppo"]
    dc = DiffCfg(
        denoising_steps=w["denoising_steps"], ft_denoising_steps=w["ft_denoising_steps"], use_ddim=w["use_ddim"],
        ddim_steps=w["ddim_steps"], eta=(w["eta"]["base_eta"] if w.get("eta") else 1.0),
        randn_clip_value=p["randn_clip_value"], min_sampling_denoising_std=p["min_sampling_denoising_std"],
        min_logprob_denoising_std=p.get("min_logprob_denoising_std", 0.1), gamma_denoising=p["gamma_denoising"],
        clip_ploss_coef=p["clip_ploss_coef"], clip_ploss_coef_base=p["clip_ploss_coef_base"],
        clip_ploss_coef_rate=p["clip_ploss_coef_rate"],
        **{k: p[k] for k in ("clip_vloss_coef", "clip_advantage_lower_quantile", "clip_advantage_upper_quantile", "eps_clip_value",
                             "final_action_clip_value", "denoised_clip_value", "norm_adv") if k in p},
    )
    return nc, dc


def oracle_params(model, requires_grad=False):
    """state_dict of a PPODiffusion-like model as the oracle's dict-of-tensors on CPU (keys actor.*, actor_ft.*, critic.*)."""
    out = {}
    for k, v in model.state_dict().items():
        if k.startswith("network."):
            continue
        t = v.detach().cpu().clone()
        if requires_grad and (k.startswith("actor_ft.") or k.startswith("critic.")):
            t.requires_grad_(True)
        out[k] = t
    return out


def agent_golden_cfg(device, logdir, reference=False, target_kl=None):
    """Config of the agent-level golden run (tests/golden/make_golden_agent.py): Hopper-shaped, small networks, 3 iterations
    with one critic warm-up iteration, LR warm-up, several minibatches per epoch.  `reference=True` points the `_target_`
    nodes at the reference classes and adds the keys only the reference constructor reads."""
    from dppo_b200.workloads import get_workload, make_agent_cfg

    w = get_workload("hopper")
    w["actor"] = dict(w["actor"], mlp_dims=[128, 128, 128])
    w["critic"] = dict(w["critic"], mlp_dims=[128, 128, 128])
    w["train"] = dict(w["train"], actor_lr=2e-3, critic_lr=1e-3, n_critic_warmup_itr=1)
    cfg = make_agent_cfg(w, device, logdir, n_envs=8, n_steps=6, n_train_itr=3, batch_size=96, update_epochs=2)
    cfg.env.reset_at_iteration = True
    cfg.train.actor_lr_scheduler = type(cfg)(first_cycle_steps=10, warmup_steps=2, min_lr=2e-4)
    cfg.train.critic_lr_scheduler = type(cfg)(first_cycle_steps=10, warmup_steps=2, min_lr=1e-4)
    cfg.train.target_kl = AGENT_GOLDEN_TARGET_KL if target_kl is None else target_kl
    cfg.train.logprob_batch_size = 16  # several chunks in the old-log-prob prologue (a multiple of n_envs, train_ppo_agent.py:24)
    if reference:
        def retarget(node):
            for k, v in list(node.items()):
                if isinstance(v, dict):
                    retarget(v)
                elif k == "_target_":
                    node[k] = v.replace("dppo_b200.", "dppo.")
        retarget(cfg.model)
        cfg.model.pop("engine_precision", None)
        cfg.wandb = None
        cfg.train.render = type(cfg)(freq=10 ** 9, num=0)
        cfg.train.save_trajs = False
        cfg.env.save_video = False
    return cfg


# chosen between two well separated approx_kl values of the recorded reference run (printed by make_golden_agent.py)
AGENT_GOLDEN_TARGET_KL = 7.5e-5
