// C-ABI entry points: context, schedule rows, weight packing, dispatch.  See include/dppo_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>

#include "internal.h"
#include "unet_plan.h"

namespace dppo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return DPPO_ERR_CUDA;
}

int pack_mlp_impl(dppo_ctx* ctx, int which, const float* const* p, int n_params, cudaStream_t st);
int sample_chain_impl(dppo_ctx* ctx, const float* state, int E, const float* noise, uint64_t seed, uint64_t offset,
                      int64_t env_offset, int deterministic, int use_base, float min_std, float* traj, float* chain,
                      const float* chains_in, float* logp, cudaStream_t st);

int pack_unet_impl(dppo_ctx* ctx, int which, const float* const* p, int n_params, cudaStream_t st);
int sample_chain_unet_impl(dppo_ctx* ctx, const float* state, int E, const float* noise, uint64_t seed, uint64_t offset,
                           int64_t env_offset, int deterministic, int use_base, float min_std, float* traj, float* chain,
                           const float* chains_in, float* logp, cudaStream_t st);

int small_chain_query_clusters(int H, int D, int Dc, int nb, int S);

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Per-step rows.  DDPM follows reference diffusion_vpg.py:214-223 (+ the chain rule :305-311); DDIM :167-213.
static int build_rows(dppo_ctx* c, const dppo_sched_desc* s) {
  c->rows.assign(c->S, StepRow{});
  for (int i = 0; i < c->S; ++i) {
    StepRow& r = c->rows[i];
    if (!c->use_ddim) {
      const int t = c->K - 1 - i;
      r.t = t;
      r.ft = t < c->ft;
      r.slot = t <= c->ft ? c->ft - t : -1;
      r.f0 = s->sqrt_recip_alphas_cumprod[t];
      r.f1 = s->sqrt_recipm1_alphas_cumprod[t];
      r.f2 = s->ddpm_mu_coef1[t];
      r.f3 = s->ddpm_mu_coef2[t];
      r.std_train = expf(0.5f * s->ddpm_logvar_clipped[t]);
      r.f2_det = r.f2, r.f3_det = r.f3, r.std_det = r.std_train;
    } else {
      r.t = s->ddim_t[i];
      r.ft = i >= c->S - c->ft;
      r.slot = i >= c->S - c->ft - 1 ? i - (c->S - c->ft - 1) : -1;
      const float a = s->ddim_alphas[i], ap = s->ddim_alphas_prev[i];
      r.f0 = powf(a, 0.5f);
      r.f1 = s->ddim_sqrt_one_minus_alphas[i];
      r.f2 = powf(ap, 0.5f);
      // sigma = clamp(eta * ((1-ap)/(1-a) * (1 - a/ap))^0.5, min=1e-10)
      const float base = powf((1.f - ap) / (1.f - a) * (1.f - a / ap), 0.5f);
      const float sig = fmaxf(c->eta * base, 1e-10f);
      r.f3 = sqrtf(fmaxf(1.f - ap - sig * sig, 0.f));
      r.std_train = expf(0.5f * logf(sig * sig));
      const float sig0 = fmaxf(0.f * base, 1e-10f);
      r.f2_det = r.f2;
      r.f3_det = sqrtf(fmaxf(1.f - ap - sig0 * sig0, 0.f));
      r.std_det = 0.f;
    }
  }
  return 0;
}

static int build_geometry(dppo_ctx* c) {
  const dppo_mlp_desc& n = c->net;
  MlpGeom& g = c->g;
  g.D = n.action_dim * n.horizon_steps;
  g.Dc_in = n.cond_dim;
  g.CH = n.cond_hidden;
  g.CO = n.cond_out;
  g.Dc = g.CH ? g.CO : n.cond_dim;
  g.td = n.time_dim;
  g.H = n.hidden_dim;
  g.nb = n.n_blocks;
  g.act = n.activation;
  g.ln = n.use_layernorm;
  if (g.D < 1 || g.D > 128) return set_error("Ta*Da = %d outside [1,128]", g.D), DPPO_ERR_INVALID;
  if (g.H % 128 || g.H < 128 || g.H > 1024) return set_error("hidden_dim %d must be a multiple of 128 in [128,1024]", g.H), DPPO_ERR_INVALID;
  if (g.td % 2 || g.td < 4 || g.td > 64) return set_error("time_dim %d unsupported", g.td), DPPO_ERR_INVALID;
  if (g.nb < 1 || g.nb > 8) return set_error("n_blocks %d unsupported", g.nb), DPPO_ERR_INVALID;
  if (g.CH && (g.CH % 128 || g.CH > g.H || g.CO > 128 || g.CO < 1 || g.Dc_in > g.H))
    return set_error("cond_mlp dims (%d,%d) unsupported", g.CH, g.CO), DPPO_ERR_INVALID;
  if (g.act != DPPO_ACT_RELU && g.act != DPPO_ACT_MISH) return set_error("activation %d unsupported", g.act), DPPO_ERR_INVALID;
  g.MT = g.H / 128;
  g.KCH = g.H / 64;
  if (2 * g.D > g.H) return set_error("Ta*Da = %d needs hidden_dim >= %d", g.D, 2 * g.D), DPPO_ERR_INVALID;
  g.KC0 = ceil_div(g.D + g.Dc, 64);
  if (g.KC0 > g.KCH) return set_error("layer-0 width %d exceeds hidden_dim", g.D + g.Dc), DPPO_ERR_INVALID;
  g.KCc = g.CH ? ceil_div(g.Dc_in, 64) : 0;
  g.MTc = g.CH / 128;
  g.nsplit = c->precision == DPPO_PRECISION_SPLIT3 ? 2 : 1;
  // fp32 side table
  size_t o = 0;
  g.off_tb = o, o += (size_t)c->K * g.H;
  g.off_blk = o, g.blk_stride = (size_t)(g.ln ? 6 : 2) * g.H, o += g.blk_stride * g.nb;
  g.off_bout = o, o += 128;
  g.off_bc0 = o, o += g.CH;
  g.off_bc1 = o, o += 128;
  g.n_side = o;
  // tile stream
  g.n_cond_tiles = g.CH ? (size_t)g.nsplit * (g.MTc * g.KCc + g.CH / 64) : 0;
  g.n_step_tiles = (size_t)g.nsplit * (g.MT * g.KC0 + (size_t)g.nb * 2 * g.MT * g.KCH + g.KCH);
  g.off_cond_tiles = 0;
  g.off_step_tiles = g.n_cond_tiles * 16384;
  g.blob_bytes = (g.n_cond_tiles + g.n_step_tiles) * 16384;
  return 0;
}

}  // namespace dppo

using namespace dppo;

extern "C" const char* dppo_last_error(void) { return g_err; }
extern "C" int dppo_version(void) { return 103; }  // 103: dppo_sample_chain_host

// schedule + geometry + device allocations shared by the two denoiser kinds
static int ctx_create_common(dppo_ctx** out, const dppo_mlp_desc* mlp, const dppo_unet_desc* unet, const dppo_sched_desc* s,
                             int precision, int device) {
  if (!out || (!mlp && !unet) || !s) return set_error("dppo_ctx_create: null argument"), DPPO_ERR_INVALID;
  if (precision != DPPO_PRECISION_SPLIT3 && precision != DPPO_PRECISION_BF16)
    return set_error("dppo_ctx_create: unknown precision %d", precision), DPPO_ERR_INVALID;
  DPPO_CUDA(cudaSetDevice(device));
  dppo_ctx* c = new dppo_ctx();
  c->device = device;
  c->precision = precision;
  c->kind = unet ? 1 : 0;
  if (mlp) c->net = *mlp;
  c->K = s->denoising_steps;
  c->ft = s->ft_denoising_steps;
  c->use_ddim = s->use_ddim;
  c->S = s->use_ddim ? s->ddim_steps : s->denoising_steps;
  c->eta = s->eta;
  c->x0_clip = s->denoised_clip_value;
  c->randn_clip = s->randn_clip_value;
  c->final_clip = s->final_action_clip_value;
  c->eps_clip = s->eps_clip_value;
  c->min_logprob_std = s->min_logprob_denoising_std;
  int rc = DPPO_OK;
  if (c->K < 1 || c->S < 1 || c->ft < 0 || c->ft > c->S) {
    set_error("dppo_ctx_create: bad schedule K=%d S=%d ft=%d", c->K, c->S, c->ft);
    rc = DPPO_ERR_INVALID;
  } else if (!s->sqrt_recip_alphas_cumprod || !s->sqrt_recipm1_alphas_cumprod || !s->ddpm_mu_coef1 ||
             !s->ddpm_mu_coef2 || !s->ddpm_logvar_clipped ||
             (s->use_ddim && (!s->ddim_t || !s->ddim_alphas || !s->ddim_alphas_prev || !s->ddim_sqrt_one_minus_alphas))) {
    set_error("dppo_ctx_create: missing schedule table");
    rc = DPPO_ERR_INVALID;
  } else if (unet) {
    c->unet = new UnetPlan();
    rc = unet_build_plan(*unet, c->K, precision, c->unet);
  } else {
    rc = build_geometry(c);
  }
  if (rc == DPPO_OK) {
    c->sample_dim = unet ? c->unet->D : c->g.D;
    if (!unet && !c->g.CH && !c->g.ln && c->g.H % 64 == 0 && c->S <= 128)
      c->small_clusters = small_chain_query_clusters(c->g.H, c->g.D, c->g.Dc, c->g.nb, c->S);
    build_rows(c, s);
    int dev_sms = 0;
    if (cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && dev_sms > 0)
      c->sm_count = dev_sms;
    cudaError_t e = cudaMalloc(&c->d_rows, sizeof(StepRow) * c->S);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_nonfinite, 16);
    if (e == cudaSuccess) e = cudaMemset(c->d_nonfinite, 0, 16);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_rows, c->rows.data(), sizeof(StepRow) * c->S, cudaMemcpyHostToDevice);
    const size_t blob = unet ? c->unet->n_tiles * 16384 : c->g.blob_bytes;
    const size_t n_side = unet ? c->unet->n_side : c->g.n_side;
    for (int w = 0; w < 2 && e == cudaSuccess; ++w) {
      e = cudaMalloc(&c->nets[w].tiles, blob);
      if (e == cudaSuccess) e = cudaMalloc(&c->nets[w].side, n_side * sizeof(float));
      if (e == cudaSuccess && unet) e = cudaMalloc(&c->d_unet_params[w], sizeof(float*) * c->unet->n_params);
    }
    if (e == cudaSuccess && unet) {
      const UnetPlan& P = *c->unet;
      auto upload = [&](auto** dst, const auto& vec) {
        using T = typename std::remove_reference<decltype(vec)>::type::value_type;
        cudaError_t r = cudaMalloc(reinterpret_cast<void**>(dst), sizeof(T) * vec.size());
        if (r == cudaSuccess) r = cudaMemcpy(*dst, vec.data(), sizeof(T) * vec.size(), cudaMemcpyHostToDevice);
        return r;
      };
      e = upload(&c->d_unet_jobs, P.jobs);
      if (e == cudaSuccess) e = upload(&c->d_unet_side_jobs, P.side_jobs);
      if (e == cudaSuccess) e = upload(&c->d_unet_layers, P.layers);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "dppo_ctx_create allocation");
  }
  if (rc != DPPO_OK) {
    dppo_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return DPPO_OK;
}

extern "C" int dppo_ctx_create(dppo_ctx** out, const dppo_mlp_desc* actor, const dppo_sched_desc* s, int precision,
                               int device) {
  if (!actor) return set_error("dppo_ctx_create: null argument"), DPPO_ERR_INVALID;
  return ctx_create_common(out, actor, nullptr, s, precision, device);
}

extern "C" int dppo_ctx_create_unet(dppo_ctx** out, const dppo_unet_desc* actor, const dppo_sched_desc* s, int precision,
                                    int device) {
  if (!actor) return set_error("dppo_ctx_create_unet: null argument"), DPPO_ERR_INVALID;
  return ctx_create_common(out, nullptr, actor, s, precision, device);
}

extern "C" int dppo_unet_param_count(const dppo_unet_desc* actor) {
  if (!actor) return set_error("dppo_unet_param_count: null argument"), DPPO_ERR_INVALID;
  UnetPlan P;
  const int rc = unet_build_plan(*actor, 1, DPPO_PRECISION_SPLIT3, &P);
  return rc == DPPO_OK ? P.n_params : rc;
}

extern "C" int dppo_ctx_destroy(dppo_ctx* c) {
  if (!c) return DPPO_OK;
  cudaFree(c->d_rows);
  cudaFree(c->d_nonfinite);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->h_done) cudaFreeHost(c->h_done);
  for (int w = 0; w < 2; ++w) {
    cudaFree(c->nets[w].tiles);
    cudaFree(c->nets[w].side);
    cudaFree(c->d_unet_params[w]);
  }
  cudaFree(c->d_unet_jobs);
  cudaFree(c->d_unet_side_jobs);
  cudaFree(c->d_unet_layers);
  delete c->unet;
  delete c;
  return DPPO_OK;
}

extern "C" int dppo_pack_mlp(dppo_ctx* ctx, int which, const float* const* params, int n_params, void* stream) {
  if (!ctx || !params) return set_error("dppo_pack_mlp: null argument"), DPPO_ERR_INVALID;
  if (which != DPPO_NET_ACTOR && which != DPPO_NET_ACTOR_FT) return set_error("dppo_pack_mlp: which=%d", which), DPPO_ERR_INVALID;
  for (int i = 0; i < n_params; ++i)
    if (!params[i]) return set_error("dppo_pack_mlp: parameter %d is null", i), DPPO_ERR_INVALID;
  if (ctx->kind != 0) return set_error("dppo_pack_mlp: the context was created for a Unet1D denoiser"), DPPO_ERR_STATE;
  return pack_mlp_impl(ctx, which, params, n_params, static_cast<cudaStream_t>(stream));
}

extern "C" int dppo_pack_unet(dppo_ctx* ctx, int which, const float* const* params, int n_params, void* stream) {
  if (!ctx || !params) return set_error("dppo_pack_unet: null argument"), DPPO_ERR_INVALID;
  if (which != DPPO_NET_ACTOR && which != DPPO_NET_ACTOR_FT) return set_error("dppo_pack_unet: which=%d", which), DPPO_ERR_INVALID;
  if (ctx->kind != 1) return set_error("dppo_pack_unet: the context was created for a DiffusionMLP denoiser"), DPPO_ERR_STATE;
  for (int i = 0; i < n_params; ++i)
    if (!params[i]) return set_error("dppo_pack_unet: parameter %d is null", i), DPPO_ERR_INVALID;
  return pack_unet_impl(ctx, which, params, n_params, static_cast<cudaStream_t>(stream));
}

extern "C" int dppo_sample_chain(dppo_ctx* ctx, const float* state, int n_envs, const float* noise, uint64_t seed,
                                 uint64_t offset, int64_t env_offset, int deterministic, int use_base_policy,
                                 float min_std, float* traj, float* chain, void* stream) {
  if (!ctx || !state || !traj) return set_error("dppo_sample_chain: null argument"), DPPO_ERR_INVALID;
  if (n_envs < 0) return set_error("dppo_sample_chain: n_envs=%d", n_envs), DPPO_ERR_INVALID;
  if (n_envs == 0) return DPPO_OK;
  if (!ctx->nets[0].packed || (!use_base_policy && !ctx->nets[1].packed))
    return set_error("dppo_sample_chain: weights not packed (call dppo_pack_mlp / dppo_pack_unet for actor and actor_ft)"), DPPO_ERR_STATE;
  return (ctx->kind == 1 ? sample_chain_unet_impl : sample_chain_impl)(
      ctx, state, n_envs, noise, seed, offset, env_offset, deterministic, use_base_policy, min_std, traj, chain, nullptr,
      nullptr, static_cast<cudaStream_t>(stream));
}

// Host-buffer form (include/dppo_b200.h): page-locked caller buffers are used in place, pageable ones are staged through
// the context's own page-locked area; returns with the results in the caller's buffers.
extern "C" int dppo_sample_chain_host(dppo_ctx* ctx, const float* state, int n_envs, uint64_t seed, uint64_t offset,
                                      int64_t env_offset, int deterministic, int use_base_policy, float min_std,
                                      float* traj, float* chain, int flags, void* stream) {
  if (!ctx || !state || !traj) return set_error("dppo_sample_chain_host: null argument"), DPPO_ERR_INVALID;
  if (n_envs < 0) return set_error("dppo_sample_chain_host: n_envs=%d", n_envs), DPPO_ERR_INVALID;
  if (n_envs == 0) return DPPO_OK;
  const size_t Dc = size_t(ctx->kind == 1 ? ctx->unet->d.cond_dim : ctx->g.Dc_in), D = size_t(ctx->sample_dim);
  const size_t sb = size_t(n_envs) * Dc * sizeof(float), tb = size_t(n_envs) * D * sizeof(float);
  const size_t cb = chain ? tb * size_t(ctx->ft + 1) : 0;
  auto up = [](size_t n) { return (n + 255) & ~size_t(255); };
  const bool stage_in = !(flags & DPPO_HOST_STATE_PINNED), stage_out = !(flags & DPPO_HOST_OUT_PINNED);
  const size_t need = (stage_in ? up(sb) : 0) + (stage_out ? up(tb) + up(cb) : 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (need > ctx->h_stage_bytes) {
    // every earlier host call has synchronised before returning: nothing reads the old area any more
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->h_stage = nullptr, ctx->h_stage_bytes = 0;
    const size_t cap = need < (size_t(1) << 16) ? (size_t(1) << 16) : need + need / 2;
    DPPO_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_stage), cap, cudaHostAllocPortable | cudaHostAllocMapped));
    ctx->h_stage_bytes = cap;
  }
  uint8_t* p = ctx->h_stage;
  const float* k_state = state;
  float *k_traj = traj, *k_chain = chain;
  if (stage_in) {
    memcpy(p, state, sb);
    k_state = reinterpret_cast<const float*>(p), p += up(sb);
  }
  if (stage_out) {
    k_traj = reinterpret_cast<float*>(p), p += up(tb);
    if (chain) k_chain = reinterpret_cast<float*>(p);
  }
  if (!ctx->h_done) {
    DPPO_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_done), 64, cudaHostAllocPortable | cudaHostAllocMapped));
    *ctx->h_done = 0;
  }
  static int env_ticket = -1;  // DPPO_B200_TICKET=0: always wait through the driver (A/B switch)
  if (env_ticket < 0) {
    const char* e = getenv("DPPO_B200_TICKET");
    env_ticket = e ? atoi(e) : 1;
  }
  ctx->done_want = env_ticket, ctx->done_armed = 0;
  const int rc = dppo_sample_chain(ctx, k_state, n_envs, nullptr, seed, offset, env_offset, deterministic, use_base_policy,
                                   min_std, k_traj, k_chain, stream);
  ctx->done_want = 0;
  if (rc != DPPO_OK) return rc;
  bool done = false;
  if (ctx->done_armed) {
    // the latency case (small-batch kernel): poll the ticket the kernel's last cluster stores into page-locked memory
    // behind its results; past a deadline (a trapped or very long launch) fall through to the driver's synchronise,
    // which also reports the error
    volatile unsigned* flag = ctx->h_done;
    const unsigned want = ctx->done_seq;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 1; !done; ++spins) {
      if (*flag == want) {
        done = true;
      } else if ((spins & 4095u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(5)) {
        break;
      } else {
        __builtin_ia32_pause();
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  if (!done) DPPO_CUDA(cudaStreamSynchronize(st));
  if (stage_out) {
    memcpy(traj, k_traj, tb);
    if (chain) memcpy(chain, k_chain, cb);
  }
  return DPPO_OK;
}

extern "C" int dppo_chain_logprobs(dppo_ctx* ctx, const float* state, const float* chains, int n_rows,
                                   int use_base_policy, float* logp, void* stream) {
  if (!ctx || !state || !chains || !logp) return set_error("dppo_chain_logprobs: null argument"), DPPO_ERR_INVALID;
  if (n_rows < 0) return set_error("dppo_chain_logprobs: n_rows=%d", n_rows), DPPO_ERR_INVALID;
  if (n_rows == 0 || ctx->ft == 0) return DPPO_OK;
  if (!ctx->nets[use_base_policy ? 0 : 1].packed)
    return set_error("dppo_chain_logprobs: weights not packed"), DPPO_ERR_STATE;
  return (ctx->kind == 1 ? sample_chain_unet_impl : sample_chain_impl)(
      ctx, state, n_rows, nullptr, 0, 0, 0, 0, use_base_policy, ctx->min_logprob_std, nullptr, nullptr, chains, logp,
      static_cast<cudaStream_t>(stream));
}

extern "C" int dppo_sample_nonfinite(dppo_ctx* ctx, int* flag, int reset, void* stream) {
  if (!ctx || !flag) return set_error("dppo_sample_nonfinite: null argument"), DPPO_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DPPO_CUDA(cudaMemcpyAsync(flag, ctx->d_nonfinite, sizeof(int), cudaMemcpyDeviceToHost, st));
  DPPO_CUDA(cudaStreamSynchronize(st));
  if (reset && *flag) DPPO_CUDA(cudaMemsetAsync(ctx->d_nonfinite, 0, sizeof(int), st));
  return DPPO_OK;
}

// bring-up hook (not in the public header): the chain kernel writes 8 cycle counters per CTA into `buf`
extern "C" int dppo_debug_set_prof(dppo_ctx* ctx, unsigned long long* buf) {
  if (!ctx) return DPPO_ERR_INVALID;
  ctx->d_prof = buf;
  return DPPO_OK;
}

// bring-up / tuning hook (not in the public header): force the chain kernel's tile size (16/32/64 envs) and cluster
// size (1/2/4/8 CTAs sharing a tile); 0 = let the cost model choose
extern "C" int dppo_debug_set_shape(dppo_ctx* ctx, int tile_envs, int cluster) {
  if (!ctx) return DPPO_ERR_INVALID;
  ctx->force_ne = tile_envs, ctx->force_c = cluster;
  return DPPO_OK;
}
