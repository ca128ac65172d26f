/*
 * dppo_b200 — C ABI of the B200-native DPPO hot path (libdppo_b200.so).
 *
 * The reference (enyen/dppo) has no FFI: its seam is the Python class named by Hydra `_target_`
 * (dppo/agent/finetune/train_agent.py:84) and the duck-typed method calls of
 * TrainPPODiffusionAgent.run (dppo/agent/finetune/train_ppo_diffusion_agent.py:47-483).  This header is the
 * C boundary a replacement for those methods binds; every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every data pointer is a CUDA device pointer owned by the caller (contiguous, fp32 row-major unless stated);
 *     pointers inside the *_desc structs are HOST pointers, read during the call only;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no call synchronises the device;
 *   - return value 0 = success, < 0 = error code below; dppo_last_error() returns a thread-local message;
 *   - a context is not thread-safe (the reference agent is single-threaded per process); the library allocates
 *     device memory only inside dppo_ctx_create / dppo_pack_* (CUDA-graph friendly afterwards).
 */
#ifndef DPPO_B200_H_
#define DPPO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPPO_OK 0
#define DPPO_ERR_INVALID (-1)     /* bad argument / unsupported shape */
#define DPPO_ERR_CUDA (-2)        /* CUDA runtime error (message holds cudaGetErrorString) */
#define DPPO_ERR_UNSUPPORTED (-3) /* feature outside the hot path (e.g. learned eta) */
#define DPPO_ERR_STATE (-4)       /* call order violated (e.g. weights not packed) */

#define DPPO_ACT_RELU 0
#define DPPO_ACT_MISH 1

#define DPPO_NET_ACTOR 0    /* frozen base policy  (state_dict keys actor.*)    */
#define DPPO_NET_ACTOR_FT 1 /* fine-tuned policy   (state_dict keys actor_ft.*) */

#define DPPO_PRECISION_SPLIT3 0 /* 3 x bf16 split MMAs (hi/lo operands), meets the fp32 1e-3 tolerance */
#define DPPO_PRECISION_BF16 1   /* single-pass bf16 MMAs, fast mode */

typedef struct dppo_ctx dppo_ctx;

/* DiffusionMLP geometry.  reference dppo/model/diffusion/mlp_diffusion.py:174-216, dppo/model/common/mlp.py:84-154 */
typedef struct dppo_mlp_desc {
  int32_t cond_dim;      /* To * Do, flattened observation */
  int32_t action_dim;    /* Da */
  int32_t horizon_steps; /* Ta */
  int32_t time_dim;      /* sinusoidal embedding width d; time MLP is Lin(d,2d) Mish Lin(2d,d) */
  int32_t hidden_dim;    /* H of the residual trunk (multiple of 128) */
  int32_t n_blocks;      /* number of two-layer pre-activation residual blocks */
  int32_t activation;    /* DPPO_ACT_* of the trunk (time / cond MLPs are Mish resp. trunk activation) */
  int32_t use_layernorm; /* LayerNorm(eps=1e-6) before each activation inside the blocks */
  int32_t cond_hidden;   /* cond_mlp_dims[0], 0 = no cond_mlp */
  int32_t cond_out;      /* cond_mlp_dims[1] */
} dppo_mlp_desc;

/* Unet1D geometry (state conditioning, no cond_mlp).  reference dppo/model/diffusion/unet.py:121-265 */
#define DPPO_UNET_MAX_LEVELS 4
typedef struct dppo_unet_desc {
  int32_t cond_dim;      /* To * Do, flattened observation (the FiLM conditioning vector is [time_mlp(t) | obs]) */
  int32_t action_dim;    /* Da = input / output channels */
  int32_t horizon_steps; /* Ta = sequence length; must be divisible by 2^(n_levels-1) */
  int32_t time_dim;      /* diffusion_step_embed_dim e; time_mlp is Lin(e,4e) Mish Lin(4e,e) */
  int32_t dim;           /* base channel width */
  int32_t n_levels;      /* len(dim_mults) */
  int32_t dim_mults[DPPO_UNET_MAX_LEVELS];
  int32_t kernel_size;   /* odd; Conv1d(padding = k/2) */
  int32_t n_groups;      /* GroupNorm groups */
  int32_t activation;    /* DPPO_ACT_* */
  int32_t cond_predict_scale; /* FiLM: 1 = scale and bias, 0 = bias only */
  int32_t larger_encoder;     /* 1: Lin act Lin act Lin; 0: act Lin   (unet.py:66-84) */
  float groupnorm_eps;
} dppo_unet_desc;

/* Schedule tables exactly as the reference builds them (fp32 tensors from the float64 cosine schedule).
 * reference dppo/model/diffusion/diffusion.py:98-148 (DDPM), :155-196 (DDIM, already flipped: index 0 = noisiest) */
typedef struct dppo_sched_desc {
  int32_t denoising_steps;    /* K */
  int32_t ft_denoising_steps; /* ft */
  int32_t use_ddim;
  int32_t ddim_steps; /* S when use_ddim */
  float eta;          /* the value EtaFixed returns (dppo/model/diffusion/eta.py:33-40); learned eta unsupported */
  float denoised_clip_value;     /* < 0: none */
  float randn_clip_value;
  float final_action_clip_value; /* < 0: none */
  float eps_clip_value;          /* < 0: none (DDIM only) */
  float min_logprob_denoising_std;
  const float* sqrt_recip_alphas_cumprod;   /* [K] host */
  const float* sqrt_recipm1_alphas_cumprod; /* [K] host */
  const float* ddpm_mu_coef1;               /* [K] host */
  const float* ddpm_mu_coef2;               /* [K] host */
  const float* ddpm_logvar_clipped;         /* [K] host */
  const int32_t* ddim_t;                    /* [S] host or NULL */
  const float* ddim_alphas;                 /* [S] host or NULL */
  const float* ddim_alphas_prev;            /* [S] host or NULL */
  const float* ddim_sqrt_one_minus_alphas;  /* [S] host or NULL */
} dppo_sched_desc;

/* PPO hyper-parameters of one loss call.  reference dppo/model/diffusion/diffusion_ppo.py:25-36,57-199 */
typedef struct dppo_loss_hp {
  int32_t ft_denoising_steps;
  int32_t horizon_steps;  /* Ta */
  int32_t action_dim;     /* Da */
  int32_t reward_horizon; /* rows h < reward_horizon enter the mean */
  int32_t norm_adv;
  float gamma_denoising;
  float clip_ploss_coef, clip_ploss_coef_base, clip_ploss_coef_rate;
  float clip_vloss_coef; /* < 0: unclipped value loss */
  float adv_clip_lo;     /* bounds applied AFTER normalisation; the caller passes quantile values, or */
  float adv_clip_hi;     /* -inf / +inf to let the kernel use min / max (quantiles 0 / 1 = no-op)    */
} dppo_loss_hp;

const char* dppo_last_error(void);
int dppo_version(void);

/* ---- context ------------------------------------------------------------------------------------------------ */
/* Replaces the constructor-time table building of DiffusionModel.__init__ (diffusion.py:98-196). */
int dppo_ctx_create(dppo_ctx** out, const dppo_mlp_desc* actor, const dppo_sched_desc* sched, int precision, int device);
int dppo_ctx_destroy(dppo_ctx* ctx);

/* Repack one network's fp32 parameters into the kernel layout (swizzled bf16 hi/lo weight tiles in streaming order,
 * per-timestep layer-0 bias table that folds the time-embedding MLP).  Call whenever the parameters change
 * (after every optimiser step for actor_ft; after VPGDiffusion.step(), diffusion_vpg.py:123-127).
 * `params` is a HOST array of DEVICE pointers in state_dict order of DiffusionMLP:
 *   time_embedding.1.{weight,bias}, time_embedding.3.{weight,bias},
 *   [cond_mlp.moduleList.0.linear_1.{weight,bias}, cond_mlp.moduleList.1.linear_1.{weight,bias}],
 *   mlp_mean.layers.0.{weight,bias},
 *   per block b: l1.{weight,bias}, l2.{weight,bias}, [norm1.{weight,bias}, norm2.{weight,bias}],
 *   mlp_mean.layers.{n_blocks+1}.{weight,bias}                                                              */
int dppo_pack_mlp(dppo_ctx* ctx, int which, const float* const* params, int n_params, void* stream);

/* ---- Unet1D denoiser ------------------------------------------------------------------------------------------ */
/* Same context type and the same dppo_sample_chain / dppo_chain_logprobs calls, with Unet1D.forward
 * (unet.py:267-327) as the per-step network.  At pack time every Conv1d / ConvTranspose1d over the short action
 * horizon is lowered to the dense map it is ([C_in*T_in] -> [C_out*T_out], a banded block-Toeplitz matrix), so the
 * whole net becomes a program of tensor-core GEMMs with GroupNorm / Mish / FiLM / residual epilogues.              */
int dppo_ctx_create_unet(dppo_ctx** out, const dppo_unet_desc* actor, const dppo_sched_desc* sched, int precision,
                         int device);
/* `params`: HOST array of DEVICE pointers, execution order:
 *   time_mlp.1.{weight,bias}, time_mlp.3.{weight,bias};
 *   residual blocks in execution order (down level l: 0, 1; mid 0, 1; up level l: 0, 1), each
 *     blocks.0.block.0.{weight,bias}, blocks.0.block.2.{weight,bias}, blocks.1.block.0.{weight,bias},
 *     blocks.1.block.2.{weight,bias}, cond_encoder Linear {weight,bias} x (3 | 1), [residual_conv.{weight,bias}];
 *   with the Downsample1d conv {weight,bias} after a down level's two blocks (all but the last level) and the
 *   Upsample1d conv {weight,bias} after an up level's two blocks;
 *   final_conv.0.block.0.{weight,bias}, final_conv.0.block.2.{weight,bias}, final_conv.1.{weight,bias}.           */
int dppo_pack_unet(dppo_ctx* ctx, int which, const float* const* params, int n_params, void* stream);
/* number of parameter tensors dppo_pack_unet expects for this geometry (< 0: error) */
int dppo_unet_param_count(const dppo_unet_desc* actor);

/* ---- rollout: K-step denoising chain ------------------------------------------------------------------------- */
/* Replaces VPGDiffusion.forward + p_mean_var (diffusion_vpg.py:139-315) and DiffusionMLP.forward
 * (mlp_diffusion.py:218-250) for all envs and all S steps in ONE persistent kernel launch.
 *   state  [E, cond_dim]                       traj  [E, Ta*Da]
 *   noise  [(S+1), E, Ta*Da] or NULL           chain [E, ft+1, Ta*Da] or NULL
 * noise slot 0 is x_T, slot i+1 the draw of step i BEFORE the +-randn_clip clamp (what torch.randn / randn_like
 * return at diffusion_vpg.py:258,294).  With noise == NULL the kernel draws Philox4x32-10 normals from
 * (seed, offset).  env_offset shifts the Philox counter for env-sharded multi-GPU runs.                          */
int dppo_sample_chain(dppo_ctx* ctx, const float* state, int n_envs, const float* noise, uint64_t seed,
                      uint64_t offset, int64_t env_offset, int deterministic, int use_base_policy,
                      float min_sampling_denoising_std, float* traj, float* chain, void* stream);

/* Host-buffer form of dppo_sample_chain: what the reference's rollout loop does around VPGDiffusion.forward
 * (train_ppo_diffusion_agent.py:107-122: torch.from_numpy(obs).to(device) -> model(cond) -> .cpu().numpy()) as ONE call.
 * `state`, `traj`, `chain` are HOST pointers; the call returns when the results are in `traj` / `chain` (it synchronises
 * `stream`).  Page-locked buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are read and written by the
 * kernel IN PLACE over PCIe - no copy launch on either side of the chain; say so in `flags`.  Pageable buffers are staged
 * through a page-locked area owned by the context (one host memcpy per buffer).  Draws come from Philox (seed, offset). */
#define DPPO_HOST_STATE_PINNED 1 /* `state` is page-locked: the kernel prologue reads it directly */
#define DPPO_HOST_OUT_PINNED 2   /* `traj` and `chain` are page-locked: the kernel stores into them directly */
int dppo_sample_chain_host(dppo_ctx* ctx, const float* state, int n_envs, uint64_t seed, uint64_t offset,
                           int64_t env_offset, int deterministic, int use_base_policy,
                           float min_sampling_denoising_std, float* traj, float* chain, int flags, void* stream);

/* NaN / Inf guard (the reference documents NaN observations from IsaacGym, README.md:184): every dppo_sample_chain
 * launch ORs a device flag when an element of `traj` is not finite.  This call copies the flag to *flag (HOST int; it
 * synchronises `stream`) and, with reset != 0, clears it.  The sampling calls themselves never synchronise.        */
int dppo_sample_nonfinite(dppo_ctx* ctx, int* flag, int reset, void* stream);

/* Replaces VPGDiffusion.get_logprobs (diffusion_vpg.py:319-396): log-density of every fine-tuned transition of
 * stored chains, network evaluation included.
 *   state [Bc, cond_dim], chains [Bc, ft+1, Ta*Da]  ->  logp [Bc*ft, Ta*Da]  (row = env-major, denoise-minor)   */
int dppo_chain_logprobs(dppo_ctx* ctx, const float* state, const float* chains, int n_rows, int use_base_policy,
                        float* logp, void* stream);

/* ---- update: elementwise halves of get_logprobs_subsample + PPODiffusion.loss -------------------------------- */
/* Gaussian log-density given the network output (diffusion_vpg.py:165-224,453-458), per-row denoising index.
 *   eps, x_prev, x_next, logp: [B, Ta*Da]; denoising_inds [B] int64                                              */
int dppo_logprob_rows(dppo_ctx* ctx, const float* eps, const float* x_prev, const float* x_next,
                      const int64_t* denoising_inds, int n_rows, float* logp, float* dlogp_deps /* or NULL */,
                      void* stream);

/* Fused minibatch gather + log-prob + clipped PPO loss, forward AND closed-form backward in one pass
 * (train_ppo_diffusion_agent.py:316-327 gathers; diffusion_vpg.py:398-461; diffusion_ppo.py:57-199).
 *   rollout buffers (all N = n_steps*n_envs rows): chains [N, ft+1, D], old_logprobs [N, ft, D],
 *     returns [N], old_values [N], advantages [N]
 *   minibatch: inds_all [global_rows] int64 flat indices into (N, ft) of the WHOLE minibatch (b = idx / ft,
 *     d = idx % ft; advantage statistics are minibatch-global); this rank owns rows
 *     [row_begin, row_begin + n_rows) of it; eps [n_rows, D] = actor_ft output for (chains[b,d], t_d, obs[b]);
 *     vpred [n_rows] = critic(obs[b]); global_rows = divisor of every mean
 *   outputs: grad_eps [n_rows, D] = d pg_loss / d eps, grad_vpred [n_rows] = d v_loss / d vpred,
 *     scalars[8] (sums over this rank's rows, ALREADY divided by global_rows, so a sum over ranks is the mean):
 *       0 pg_loss, 1 v_loss, 2 approx_kl, 3 clipfrac, 4 ratio, 5 reserved, 6 adv mean, 7 adv std
 *   workspace: >= 128 bytes of device scratch, 8-byte aligned.                                                   */
int dppo_ppo_loss_fwd_bwd(dppo_ctx* ctx, const float* chains, const float* old_logprobs, const float* returns,
                          const float* old_values, const float* advantages, const int64_t* inds_all, int row_begin,
                          const float* eps, const float* vpred, int n_rows, int global_rows, const dppo_loss_hp* hp,
                          float* grad_eps, float* grad_vpred, float* scalars, void* workspace, void* stream);

/* Same loss with the inputs ALREADY gathered per minibatch row, i.e. the argument list of PPODiffusion.loss
 * (diffusion_ppo.py:57-69): x_prev / x_next / old_logprobs [B, D], returns / old_values / advantages [B],
 * denoising_inds [B] int64.  Single-process semantics: every mean is over these B rows.                          */
int dppo_ppo_loss_rows(dppo_ctx* ctx, const float* x_prev, const float* x_next, const float* old_logprobs,
                       const float* returns, const float* old_values, const float* advantages,
                       const int64_t* denoising_inds, const float* eps, const float* vpred, int n_rows,
                       const dppo_loss_hp* hp, float* grad_eps, float* grad_vpred, float* scalars, void* workspace,
                       void* stream);

/* ---- update: the PPO minibatch on hand-written tensor-core kernels ------------------------------------------------- */
/* Replaces torch autograd (cuBLAS GEMMs + elementwise kernels) for the update of a DiffusionMLP actor_ft and a
 * residual-MLP critic: PPODiffusion.loss -> get_logprobs_subsample -> actor_ft forward (diffusion_vpg.py:398-461,
 * mlp_diffusion.py:218-250, common/mlp.py:84-154), CriticObs.forward (common/critic.py:40-54) and loss.backward()
 * (train_ppo_diffusion_agent.py:360-364).  Every Linear is one tcgen05 GEMM launch (forward, dgrad, wgrad) with the
 * activation / residual / bf16 hi-lo split fused into its epilogue; gradients are ACCUMULATED (+=) straight into the
 * caller's gradient tensors (e.g. the views of one flat all-reduce buffer), which the caller zeroes.              */
typedef struct dppo_update dppo_update;

/* Residual MLP of the critic: Linear(in, H), n_blocks x [h + l2(act(l1(act(h))))], Linear(H, out).
 * reference dppo/model/common/critic.py:15-54, dppo/model/common/mlp.py:84-154 */
typedef struct dppo_resmlp_desc {
  int32_t in_dim;        /* To * Do */
  int32_t hidden_dim;
  int32_t n_blocks;
  int32_t out_dim;       /* 1 */
  int32_t activation;    /* DPPO_ACT_* */
  int32_t use_layernorm;
} dppo_resmlp_desc;

/* One minibatch (slice).  Gather mode (inds_all != NULL): obs / chains / old_logprobs / returns / old_values /
 * advantages are the whole rollout buffers ([N, cond_dim], [N, ft+1, D], [N, ft, D], [N] x 3) and inds_all the
 * minibatch's flat indices into (N, ft) (train_ppo_diffusion_agent.py:316-327); this rank evaluates rows
 * [row_begin, row_begin + n_rows) and every mean divides by global_rows.  Direct mode (inds_all == NULL): the
 * arguments of PPODiffusion.loss, already gathered per row (chains = chains_prev [n_rows, D], x_next, old_logprobs
 * [n_rows, D], per-row scalars, denoising_inds [n_rows]); global_rows = n_rows.                                    */
typedef struct dppo_update_batch {
  const float* obs;
  const float* chains;
  const float* x_next;
  const float* old_logprobs;
  const float* returns;
  const float* old_values;
  const float* advantages;
  const int64_t* inds_all;
  const int64_t* denoising_inds;
  int32_t row_begin, n_rows, global_rows;
} dppo_update_batch;

/* Workspace for up to max_rows minibatch rows per call (activations, operand images, packed weights).            */
int dppo_update_create(dppo_update** out, dppo_ctx* ctx, const dppo_resmlp_desc* critic, int max_rows);
int dppo_update_destroy(dppo_update* up);
/* HOST arrays of DEVICE pointers to the fp32 parameters and to their gradient tensors: actor_ft in dppo_pack_mlp
 * order; critic as layers.0.{weight,bias}, per block l1.{weight,bias}, l2.{weight,bias} [, norm1.{weight,bias},
 * norm2.{weight,bias}], layers.last.{weight,bias}.  The current parameter VALUES are re-read on every call.      */
int dppo_update_bind(dppo_update* up, const float* const* actor_params, float* const* actor_grads, int n_actor,
                     const float* const* critic_params, float* const* critic_grads, int n_critic);
/* actor_ft(x_prev, t_d, obs) -> eps_out [n_rows, D] and critic(obs) -> vpred_out [n_rows]; NULL = library-owned
 * buffers (dppo_update_buffers).                                                                                   */
int dppo_update_forward(dppo_update* up, const dppo_update_batch* batch, float* eps_out, float* vpred_out,
                        void* stream);
/* critic(obs) -> vpred_out [n_rows] for plain observation rows obs [n_rows, cond_dim] (n_rows <= max_rows): the value
 * pass over the rollout buffer and the bootstrap value (train_ppo_diffusion_agent.py:197-206, 259-263;
 * common/critic.py:40-54) on the same kernels as the minibatch forward.  Invalidates the activations a later
 * dppo_update_backward would need (call dppo_update_forward again first).                                          */
int dppo_update_values(dppo_update* up, const float* obs, int n_rows, float* vpred_out, void* stream);
/* Back-propagate grad_eps [n_rows, D] / grad_vpred [n_rows] of the last forward into the bound gradient tensors.
 * scale_pg / scale_v: optional DEVICE scalars multiplied into the two gradients (autograd's incoming factors),
 * vf_coef an immediate factor on grad_vpred.  with_actor = 0 skips actor_ft (critic warm-up iterations).          */
int dppo_update_backward(dppo_update* up, const float* grad_eps, const float* grad_vpred, const float* scale_pg,
                         const float* scale_v, float vf_coef, int with_actor, int with_critic, void* stream);
/* Optional cudaEvent_t (NULL = off) that dppo_update_backward / dppo_update_minibatch record on the stream BEHIND the
 * actor backward and in front of the critic backward: the actor's gradient segment is final there, so a multi-GPU
 * caller can start its all-reduce on another stream while the critic backward runs (the reference's DDP-style overlap
 * of the gradient reduction with backward).                                                   */
int dppo_update_set_actor_event(dppo_update* up, void* event);
/* forward + fused PPO loss (dppo_ppo_loss_fwd_bwd / dppo_ppo_loss_rows) + backward of pg_loss + vf_coef * v_loss.
 * scalars[8] as in dppo_ppo_loss_fwd_bwd; workspace >= 128 bytes.                                                 */
int dppo_update_minibatch(dppo_update* up, const dppo_update_batch* batch, const dppo_loss_hp* hp, float vf_coef,
                          int with_actor, float* scalars, void* workspace, void* stream);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream`: zeroes the flat gradient buffer before a minibatch (what
 * optimizer.zero_grad() does at train_ppo_diffusion_agent.py:336-338) without a framework fill kernel.            */
int dppo_memset_zero(void* ptr, size_t bytes, void* stream);
/* the library-owned buffers of the last forward: eps, vpred, and the gradient buffers dppo_update_minibatch fills   */
int dppo_update_buffers(dppo_update* up, float** eps, float** vpred, float** grad_eps, float** grad_vpred);

/* ---- GAE ------------------------------------------------------------------------------------------------------ */
/* Reverse scan of train_ppo_diffusion_agent.py:255-279, one thread per env, float64 like the reference's numpy.
 *   reward, terminated, values: [n_steps, n_envs] float64; next_value [n_envs] float64 (critic of the post-rollout
 *   observation); outputs advantages, returns [n_steps, n_envs] float64.                                        */
int dppo_gae_f64(const double* reward, const double* terminated, const double* values, const double* next_value,
                 int n_steps, int n_envs, double gamma, double gae_lambda, double reward_scale_const,
                 double* advantages, double* returns, void* stream);

/* ---- update-path GEMM operands --------------------------------------------------------------------------------- */
/* fp32 [rows, cols] (row stride ldx) -> bf16 [rows, 3 * cols_p], cols_p = cols rounded up to 8 (zero padded), each row
 * [hi | hi | lo] (pattern 0: activations, gradients) or [hi | lo | hi] (pattern 1: weights), hi = bf16(x),
 * lo = bf16(x - hi).  A bf16 GEMM with fp32 accumulation over these rows is the 3-product split
 * x w ~= x_hi w_hi + x_hi w_lo + x_lo w_hi the chain kernel uses, i.e. fp32-grade Linear layers (forward, dgrad, wgrad
 * of actor_ft / critic: diffusion_vpg.py:398-461, critic.py:40-54, train_ppo_diffusion_agent.py:360-364) on the tensor
 * cores instead of fp32 SIMT GEMMs.  extra_mode 1 appends a column of ones before the padding (activations), 2 appends
 * extra[row] (weights: the bias), so that y = x W^T + b is ONE GEMM and the bias gradient is the last column of the
 * wgrad GEMM; cols_p then rounds cols + 1 up.  `out` must be 8-byte aligned.                                        */
int dppo_split3_pack(const float* x, int64_t rows, int cols, int64_t ldx, int extra_mode, const float* extra, void* out,
                     int pattern, void* stream);

/* ---- running reward scaling ----------------------------------------------------------------------------------- */
/* RunningRewardScaler.__call__ (dppo/util/reward_scaling.py:42-87; call site train_ppo_diffusion_agent.py:243-247) on
 * the device, float64, (n_steps, n_envs) row-major like the GAE inputs it feeds:
 *   phase 0  per-env forward scan rets_t = r_t + (1 - first_t) gamma rets_{t-1} from ret_state[n_envs] (updated),
 *            rets -> rets_scratch, ws[0] = sum(rets)
 *   phase 1  ws[1] = sum((rets - ws[0] / n_global)^2)
 *   phase 2  stats[3] = (mean, var, count) <- parallel-variance update with the batch (ws[0..1], n_global);
 *            scaled = clip(reward / sqrt(var_new + epsilon), +-cliprew)
 * A single process calls the phases back to back with n_global = n_steps * n_envs; env-sharded ranks all-reduce
 * ws[0] after phase 0 and ws[1] after phase 1 and pass the global element count.  ws: >= 8 doubles.              */
int dppo_reward_scale_f64(const double* reward, const double* first, int n_steps, int n_envs, long long n_global,
                          double gamma, double epsilon, double cliprew, double* ret_state, double* stats,
                          double* rets_scratch, double* ws, double* scaled, int phase, void* stream);

/* ---- optimiser ------------------------------------------------------------------------------------------------ */
/* torch.optim.AdamW step (train_ppo_diffusion_agent.py:360-373, optimisers built at train_ppo_agent.py:34-53) over ONE
 * flat fp32 segment: params / grads / exp_avg / exp_avg_sq are 16-byte aligned device arrays of n floats (the gradient
 * segment is the flat all-reduced buffer the backward writes into).  `step` = 1-based step count (bias correction).
 * max_grad_norm >= 0 applies torch.nn.utils.clip_grad_norm_'s coefficient min(1, max / (||g||_2 + 1e-6)) computed on the
 * device (workspace: >= 8 bytes); < 0 = no clipping.                                                              */
int dppo_adamw_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, float max_grad_norm,
                    void* workspace, void* stream);

/* The same update with the STEP COUNT on the device (CUDA-graph friendly: a replayed launch advances its own bias
 * correction) and an optional stop flag.  step_state: 4 ints on the device, [0] = step count (the call increments it),
 * [2..3] = scratch for the bias corrections of the step; zero it once.  stop_flag (device int or NULL): once set, the call
 * is a no-op and does not advance the step count.                                                                   */
int dppo_adamw_flat_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int* step_state, const int* stop_flag,
                        float max_grad_norm, void* workspace, void* stream);

/* KL early stop on the device (train_ppo_diffusion_agent.py:376-382).  Called after the gradient all-reduce and the
 * optimiser steps of every minibatch: appends scalars[8] (dppo_ppo_loss_fwd_bwd layout) to history[k] with k = state[2]
 * (a device-side minibatch counter, incremented here) and, when use_target and scalars[2] (approx_kl) > target_kl, sets
 * state[0] = 1 and state[1] = k.  Later optimiser launches that were handed &state[0] as stop_flag do nothing, so the
 * host may poll the flag lazily instead of synchronising on every minibatch.  state: 4 device ints, zeroed per update. */
int dppo_kl_check(const float* scalars, float target_kl, int use_target, int* state, float* history, int max_history,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPPO_B200_H_ */
