"""
The five BASELINE.json workloads as plain dicts (plus YAML-faithful / literal variants).

Every value is taken from the reference YAML cited in `yaml`; `n_envs` is the BASELINE.json figure (synthetic envs).
The dicts carry constructor keywords only — they are consumed by dppo_b200 (build_model), by the golden-vector
generator (which feeds the same keywords to the reference classes) and by the tests (oracle NetCfg / DiffCfg).
"""

from copy import deepcopy

_GYM_PPO = dict(
    gamma_denoising=0.99, clip_ploss_coef=0.01, clip_ploss_coef_base=0.01, clip_ploss_coef_rate=3,
    randn_clip_value=3, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1,
)
_ROBOMIMIC_PPO = dict(
    gamma_denoising=0.99, clip_ploss_coef=0.01, clip_ploss_coef_base=0.001, clip_ploss_coef_rate=3,
    randn_clip_value=3, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1,
)
_CRITIC_256 = dict(mlp_dims=[256, 256, 256], activation_type="Mish", residual_style=True)

WORKLOADS = {
    # cfg1: dppo/cfg/gym/finetune/hopper-v2/ft_ppo_diffusion_mlp.yaml
    "hopper": dict(
        yaml="gym/finetune/hopper-v2/ft_ppo_diffusion_mlp.yaml",
        n_envs=40, obs_dim=11, action_dim=3, horizon_steps=4, act_steps=4, cond_steps=1,
        denoising_steps=20, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
        actor=dict(kind="mlp", time_dim=16, mlp_dims=[512, 512, 512], activation_type="ReLU", residual_style=True),
        critic=_CRITIC_256, ppo=_GYM_PPO,
        train=dict(n_steps=500, gamma=0.99, gae_lambda=0.95, batch_size=50000, update_epochs=5, vf_coef=0.5,
                   target_kl=1, actor_lr=1e-4, critic_lr=1e-3, n_critic_warmup_itr=0, max_episode_steps=1000),
    ),
    # cfg2: dppo/cfg/gym/finetune/walker2d-v2/ft_ppo_diffusion_mlp.yaml scaled to 4096 synthetic envs
    "walker2d": dict(
        yaml="gym/finetune/walker2d-v2/ft_ppo_diffusion_mlp.yaml",
        n_envs=4096, obs_dim=17, action_dim=6, horizon_steps=4, act_steps=4, cond_steps=1,
        denoising_steps=20, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
        actor=dict(kind="mlp", time_dim=16, mlp_dims=[512, 512, 512], activation_type="ReLU", residual_style=True),
        critic=_CRITIC_256, ppo=_GYM_PPO,
        train=dict(n_steps=50, gamma=0.99, gae_lambda=0.95, batch_size=50000, update_epochs=5, vf_coef=0.5,
                   target_kl=1, actor_lr=1e-4, critic_lr=1e-3, n_critic_warmup_itr=0, max_episode_steps=1000),
    ),
    # cfg3: dppo/cfg/robomimic/finetune/transport/ft_ppo_diffusion_mlp.yaml with BASELINE.json's literal K=100
    "transport": dict(
        yaml="robomimic/finetune/transport/ft_ppo_diffusion_mlp.yaml",
        n_envs=50, obs_dim=59, action_dim=14, horizon_steps=8, act_steps=8, cond_steps=1,
        denoising_steps=100, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
        actor=dict(kind="mlp", time_dim=32, mlp_dims=[1024, 1024, 1024], activation_type="Mish", residual_style=True),
        critic=_CRITIC_256, ppo=_ROBOMIMIC_PPO,
        train=dict(n_steps=400, gamma=0.999, gae_lambda=0.95, batch_size=10000, update_epochs=5, vf_coef=0.5,
                   target_kl=1, actor_lr=1e-4, critic_lr=1e-3, n_critic_warmup_itr=2, max_episode_steps=800),
    ),
    # cfg4: dppo/cfg/furniture/finetune/one_leg_low/ft_ppo_diffusion_mlp.yaml (YAML-faithful: DDIM-5 of K=100, eta=1)
    "furniture": dict(
        yaml="furniture/finetune/one_leg_low/ft_ppo_diffusion_mlp.yaml",
        n_envs=1000, obs_dim=58, action_dim=10, horizon_steps=8, act_steps=8, cond_steps=1,
        denoising_steps=100, ft_denoising_steps=5, use_ddim=True, ddim_steps=5,
        eta=dict(base_eta=1, min_eta=0.1, max_eta=1.0),
        actor=dict(kind="mlp", time_dim=32, mlp_dims=[1024] * 7, cond_mlp_dims=[512, 64], use_layernorm=True,
                   activation_type="Mish", residual_style=True),
        critic=dict(mlp_dims=[512, 512, 512], activation_type="Mish", residual_style=True),
        ppo=dict(gamma_denoising=0.9, clip_ploss_coef=0.001, clip_ploss_coef_base=0.001, clip_ploss_coef_rate=3,
                 randn_clip_value=3, min_sampling_denoising_std=0.04),
        train=dict(n_steps=88, gamma=0.999, gae_lambda=0.95, batch_size=17600, update_epochs=5, vf_coef=0.5,
                   target_kl=1, actor_lr=1e-5, critic_lr=1e-3, n_critic_warmup_itr=1, max_episode_steps=700),
    ),
    # cfg5: dppo/cfg/robomimic/finetune/square/ft_ppo_diffusion_unet.yaml + BASELINE.json's DDIM-10, 1024 envs
    "square_unet": dict(
        yaml="robomimic/finetune/square/ft_ppo_diffusion_unet.yaml",
        n_envs=1024, obs_dim=23, action_dim=7, horizon_steps=4, act_steps=4, cond_steps=1,
        denoising_steps=20, ft_denoising_steps=10, use_ddim=True, ddim_steps=10,
        eta=dict(base_eta=1, min_eta=0.1, max_eta=1.0),
        actor=dict(kind="unet", diffusion_step_embed_dim=16, dim=64, dim_mults=[1, 2], kernel_size=5, n_groups=8,
                   smaller_encoder=False, cond_predict_scale=True),
        critic=_CRITIC_256, ppo=_ROBOMIMIC_PPO,
        train=dict(n_steps=40, gamma=0.999, gae_lambda=0.95, batch_size=10000, update_epochs=10, vf_coef=0.5,
                   target_kl=1, actor_lr=2e-5, critic_lr=1e-3, n_critic_warmup_itr=2, max_episode_steps=400),
    ),
}

# The remaining DiffusionMLP geometries among the reference's state-based fine-tuning YAMLs (parity cases, not bench lines)
WORKLOADS["kitchen"] = dict(  # dppo/cfg/gym/finetune/kitchen-{complete,mixed,partial}-v0/ft_ppo_diffusion_mlp.yaml
    yaml="gym/finetune/kitchen-complete-v0/ft_ppo_diffusion_mlp.yaml",
    n_envs=40, obs_dim=60, action_dim=9, horizon_steps=4, act_steps=4, cond_steps=1,
    denoising_steps=20, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
    actor=dict(kind="mlp", time_dim=16, mlp_dims=[256, 256, 256], cond_mlp_dims=[128, 32], residual_style=True),
    critic=_CRITIC_256, ppo=_GYM_PPO,
    train=dict(n_steps=70, gamma=0.99, gae_lambda=0.95, batch_size=5600, update_epochs=10, vf_coef=0.5,
               target_kl=1, actor_lr=1e-4, critic_lr=1e-3, n_critic_warmup_itr=0, max_episode_steps=280),
)
WORKLOADS["avoid"] = dict(  # dppo/cfg/d3il/finetune/avoid_m{1,2,3}/ft_ppo_diffusion_mlp.yaml
    yaml="d3il/finetune/avoid_m1/ft_ppo_diffusion_mlp.yaml",
    n_envs=50, obs_dim=4, action_dim=2, horizon_steps=4, act_steps=4, cond_steps=1,
    denoising_steps=20, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
    actor=dict(kind="mlp", time_dim=16, mlp_dims=[512, 512, 512], activation_type="ReLU", residual_style=True),
    critic=_CRITIC_256,
    ppo=dict(gamma_denoising=0.95, clip_ploss_coef=0.1, clip_ploss_coef_base=0.1, clip_ploss_coef_rate=1,
             randn_clip_value=3, min_sampling_denoising_std=0.1, min_logprob_denoising_std=0.1),
    train=dict(n_steps=25, gamma=0.99, gae_lambda=0.95, batch_size=6250, update_epochs=10, vf_coef=0.5,
               target_kl=1, actor_lr=1e-5, critic_lr=1e-3, n_critic_warmup_itr=1, max_episode_steps=100),
)
WORKLOADS["square_mlp"] = dict(  # dppo/cfg/robomimic/finetune/square/ft_ppo_diffusion_mlp.yaml
    yaml="robomimic/finetune/square/ft_ppo_diffusion_mlp.yaml",
    n_envs=50, obs_dim=23, action_dim=7, horizon_steps=4, act_steps=4, cond_steps=1,
    denoising_steps=20, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
    actor=dict(kind="mlp", time_dim=32, mlp_dims=[1024, 1024, 1024], cond_mlp_dims=[512, 64], residual_style=True),
    critic=_CRITIC_256, ppo=_ROBOMIMIC_PPO,
    train=dict(n_steps=400, gamma=0.999, gae_lambda=0.95, batch_size=10000, update_epochs=10, vf_coef=0.5,
               target_kl=1, actor_lr=1e-4, critic_lr=1e-3, n_critic_warmup_itr=2, max_episode_steps=400),
)

WORKLOADS["can_unet"] = dict(  # dppo/cfg/robomimic/finetune/{can,lift}/ft_ppo_diffusion_unet.yaml (DDPM-20; dim 40: GroupNorm groups of 20)
    yaml="robomimic/finetune/can/ft_ppo_diffusion_unet.yaml",
    n_envs=50, obs_dim=23, action_dim=7, horizon_steps=4, act_steps=4, cond_steps=1,
    denoising_steps=20, ft_denoising_steps=10, use_ddim=False, ddim_steps=None, eta=None,
    actor=dict(kind="unet", diffusion_step_embed_dim=16, dim=40, dim_mults=[1, 2], kernel_size=5, n_groups=8,
               smaller_encoder=False, cond_predict_scale=True),
    critic=_CRITIC_256, ppo=_ROBOMIMIC_PPO,
    train=dict(n_steps=300, gamma=0.999, gae_lambda=0.95, batch_size=7500, update_epochs=10, vf_coef=0.5,
               target_kl=1, actor_lr=1e-4, critic_lr=1e-3, n_critic_warmup_itr=2, max_episode_steps=300),
)

# variants: same networks, different schedules
WORKLOADS["transport_k20"] = dict(deepcopy(WORKLOADS["transport"]), denoising_steps=20)
WORKLOADS["furniture_ddpm100"] = dict(deepcopy(WORKLOADS["furniture"]), use_ddim=False, ddim_steps=None, eta=None)


def get_workload(name: str) -> dict:
    if name not in WORKLOADS:
        raise KeyError(f"unknown workload {name!r}; have {sorted(WORKLOADS)}")
    return deepcopy(WORKLOADS[name])


def chain_evals(w: dict) -> int:
    """Number of network evaluations per sampled action chunk (S in SURVEY.md §8)."""
    return w["ddim_steps"] if w["use_ddim"] else w["denoising_steps"]


def make_agent_cfg(w: dict, device: str, logdir: str, n_envs: int = None, n_steps: int = None, n_train_itr: int = 2,
                   batch_size: int = None, update_epochs: int = None, seed: int = 42, precision: str = None):
    """
    Config tree for dppo_b200.agent.finetune.train_ppo_diffusion_agent.TrainPPODiffusionAgent with the keys of the
    reference YAML cited in `w["yaml"]` (model node = Hydra-style `_target_` dicts); DiffusionMLP or Unet1D actors.
    """
    from dppo_b200.util.config import Cfg

    a = dict(w["actor"])
    kind = a.pop("kind")
    t = w["train"]
    cond_dim = w["obs_dim"] * w["cond_steps"]
    E = n_envs or w["n_envs"]
    if kind == "mlp":
        actor = dict({"_target_": "dppo_b200.model.diffusion.mlp_diffusion.DiffusionMLP", "action_dim": w["action_dim"],
                      "horizon_steps": w["horizon_steps"], "cond_dim": cond_dim}, **a)
    else:
        actor = dict({"_target_": "dppo_b200.model.diffusion.unet.Unet1D", "action_dim": w["action_dim"],
                      "cond_dim": cond_dim}, **a)
    model = {
        "_target_": "dppo_b200.model.diffusion.diffusion_ppo.PPODiffusion",
        "actor": actor,
        "critic": dict({"_target_": "dppo_b200.model.common.critic.CriticObs", "cond_dim": cond_dim}, **w["critic"]),
        "ft_denoising_steps": w["ft_denoising_steps"], "horizon_steps": w["horizon_steps"], "obs_dim": w["obs_dim"],
        "action_dim": w["action_dim"], "denoising_steps": w["denoising_steps"], "device": device,
        "use_ddim": w["use_ddim"], "ddim_steps": w["ddim_steps"], "network_path": None, "learn_eta": False,
        "engine_precision": precision, **w["ppo"],
    }
    if w.get("eta"):
        model["eta"] = dict({"_target_": "dppo_b200.model.diffusion.eta.EtaFixed"}, **w["eta"])
    sched = dict(first_cycle_steps=1000, warmup_steps=10, min_lr=t["actor_lr"])
    return Cfg(
        device=device, seed=seed, logdir=logdir, obs_dim=w["obs_dim"], action_dim=w["action_dim"], cond_steps=w["cond_steps"],
        act_steps=w["act_steps"], horizon_steps=w["horizon_steps"],
        env=dict(n_envs=E, name="synthetic", max_episode_steps=t["max_episode_steps"], reset_at_iteration=False,
                 best_reward_threshold_for_success=3),
        train=dict(n_train_itr=n_train_itr, val_freq=10 ** 9, force_train=True, n_steps=n_steps or t["n_steps"],
                   gamma=t["gamma"], gae_lambda=t["gae_lambda"], n_critic_warmup_itr=t["n_critic_warmup_itr"],
                   actor_lr=t["actor_lr"], critic_lr=t["critic_lr"], actor_weight_decay=0, critic_weight_decay=0,
                   actor_lr_scheduler=sched, critic_lr_scheduler=dict(sched, min_lr=t["critic_lr"]),
                   batch_size=batch_size or t["batch_size"], update_epochs=update_epochs or t["update_epochs"],
                   vf_coef=t["vf_coef"], target_kl=t["target_kl"], reward_scale_running=True, reward_scale_const=1.0,
                   save_model_freq=10 ** 9, logprob_batch_size=max(E, (10000 // E) * E)),
        model=model,
    )
