// Update-side kernels (HBM-bound, no tensor cores):
//   gae_kernel            per-env reverse scan, float64           reference train_ppo_diffusion_agent.py:255-279
//   logprob_rows_kernel   Gaussian log-density of (x_prev -> x_next) given the network output eps
//                                                                  reference diffusion_vpg.py:165-224,453-458
//   adv_partial/finalize  minibatch advantage mean / unbiased std / min / max          diffusion_ppo.py:129-136
//   ppo_loss_kernel       fused gather + log-prob + clipped-ratio loss + value loss, forward and closed-form backward
//                         (one group of G lanes per minibatch row, float4 loads, shuffle reductions)
//                                                                  reference diffusion_ppo.py:57-199, SURVEY.md §3.3
#include <math.h>

#include "common.cuh"
#include "internal.h"

namespace dppo {

// ------------------------------------------------------------------------------------------------ GAE
__global__ void gae_kernel(const double* __restrict__ reward, const double* __restrict__ terminated,
                           const double* __restrict__ values, const double* __restrict__ next_value, int n_steps,
                           int E, double gamma, double lam, double scale, double* __restrict__ adv,
                           double* __restrict__ ret) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double last = 0.0;
  double nextv = next_value[e];
  for (int t = n_steps - 1; t >= 0; --t) {
    const size_t i = size_t(t) * E + e;
    const double live = 1.0 - terminated[i];
    const double v = values[i];
    const double delta = reward[i] * scale + gamma * nextv * live - v;
    last = delta + gamma * lam * live * last;
    adv[i] = last;
    ret[i] = last + v;
    nextv = v;
  }
}

// ------------------------------------------------------------------------------------------------ posterior helpers
struct Posterior {
  float mu;
  float dmu_deps;  // d mu / d eps including the x0 clamp mask
};

__device__ __forceinline__ Posterior posterior(const StepRow& r, bool ddim, float x0_clip, float eps_clip, float x,
                                               float eps) {
  Posterior p;
  if (!ddim) {
    float x0 = r.f0 * x - r.f1 * eps;
    float m = 1.f;
    if (x0_clip >= 0.f) {
      m = (x0 >= -x0_clip && x0 <= x0_clip) ? 1.f : 0.f;
      x0 = fminf(fmaxf(x0, -x0_clip), x0_clip);
    }
    p.mu = r.f2 * x0 + r.f3 * x;
    p.dmu_deps = -r.f2 * r.f1 * m;
  } else {
    float x0 = (x - r.f1 * eps) / r.f0;
    const float dx0 = -r.f1 / r.f0;
    if (x0_clip >= 0.f) {
      const float m = (x0 >= -x0_clip && x0 <= x0_clip) ? 1.f : 0.f;
      x0 = fminf(fmaxf(x0, -x0_clip), x0_clip);
      float e2 = (x - r.f0 * x0) / r.f1;
      float me = 1.f;
      if (eps_clip >= 0.f) {
        me = (e2 >= -eps_clip && e2 <= eps_clip) ? 1.f : 0.f;
        e2 = fminf(fmaxf(e2, -eps_clip), eps_clip);
      }
      p.mu = r.f2 * x0 + r.f3 * e2;
      p.dmu_deps = (r.f2 - r.f3 * me * r.f0 / r.f1) * dx0 * m;
    } else {
      float e2 = eps, me = 1.f;
      if (eps_clip >= 0.f) {
        me = (e2 >= -eps_clip && e2 <= eps_clip) ? 1.f : 0.f;
        e2 = fminf(fmaxf(e2, -eps_clip), eps_clip);
      }
      p.mu = r.f2 * x0 + r.f3 * e2;
      p.dmu_deps = r.f2 * dx0 + r.f3 * me;
    }
  }
  return p;
}

constexpr float kHalfLog2Pi = 0.91893853320467274f;

__global__ void logprob_rows_kernel(const float* __restrict__ eps, const float* __restrict__ x_prev,
                                    const float* __restrict__ x_next, const int64_t* __restrict__ dinds,
                                    const StepRow* __restrict__ rows, int row0, int ddim, float x0_clip,
                                    float eps_clip, float min_std, int D, long long total, float* __restrict__ logp,
                                    float* __restrict__ dlogp_deps) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = int(i / D);
  const StepRow r = rows[row0 + int(dinds[b])];
  const Posterior p = posterior(r, ddim != 0, x0_clip, eps_clip, x_prev[i], eps[i]);
  const float sd = fmaxf(r.std_train, min_std);
  const float diff = x_next[i] - p.mu;
  logp[i] = -(diff * diff) / (2.f * (sd * sd)) - logf(sd) - kHalfLog2Pi;
  if (dlogp_deps) dlogp_deps[i] = diff / (sd * sd) * p.dmu_deps;
}

// ------------------------------------------------------------------------------------------------ advantage stats
// ws[0..3] (double) = mean, unbiased std, min, max of advantages[inds[i] / ft].  Grid-wide: every block adds its partial
// sum / sum of squares (double) and its min / max (order-preserving integer keys) into ws[4..7]; a one-thread kernel
// turns them into the four statistics.  (inds == nullptr: advantages are already per row.)
__device__ __forceinline__ int float_order_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;  // signed-integer order == float order
}
__device__ __forceinline__ float float_from_key(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7FFFFFFF); }

__global__ void adv_partial_kernel(const float* __restrict__ adv, const int64_t* __restrict__ inds, int n, int ft,
                                   double* __restrict__ ws) {
  double s1 = 0.0, s2 = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float a = adv[inds ? inds[i] / ft : i];
    s1 += a, s2 += double(a) * double(a);
    mn = fminf(mn, a), mx = fmaxf(mx, a);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&ws[4], s1);
    atomicAdd(&ws[5], s2);
    int* keys = reinterpret_cast<int*>(&ws[6]);
    atomicMin(&keys[0], float_order_key(mn));
    atomicMax(&keys[1], float_order_key(mx));
  }
}

__global__ void adv_finalize_kernel(int n, double* __restrict__ ws) {
  const double mean = ws[4] / n;
  const double ss = fmax(ws[5] - double(n) * mean * mean, 0.0);
  const int* keys = reinterpret_cast<const int*>(&ws[6]);
  ws[0] = mean;
  ws[1] = n > 1 ? sqrt(ss / (n - 1)) : NAN;
  ws[2] = float_from_key(keys[0]);
  ws[3] = float_from_key(keys[1]);
}

// ------------------------------------------------------------------------------------------------ fused loss
struct LossArgs {
  const float *chains, *old_lp, *returns, *old_values, *adv, *eps, *vpred;
  const float* x_next;    // direct mode only (chains = x_prev rows)
  const int64_t* dinds;   // direct mode only: denoising index per row
  int direct;             // 1: inputs are already gathered per row (PPODiffusion.loss signature)
  const int64_t* inds;    // gather mode: this rank's slice of the minibatch
  const StepRow* rows;
  int row0;             // S - ft
  int n_rows, global_rows, D, ft, ddim;
  int horizon_elems;    // reward_horizon * Da
  int norm_adv;
  float x0_clip, eps_clip, min_std;
  float gamma_denoising, clip_coef, clip_base, clip_rate, clip_v;
  float adv_lo, adv_hi;
  float *grad_eps, *grad_v;
  double* ws;  // [0..3] adv stats (in), [8..12] sums (out)
};

template <int G, int VEC>
__global__ void __launch_bounds__(256) ppo_loss_kernel(const LossArgs a) {
  constexpr int kMaxFt = 128;
  __shared__ float s_disc[kMaxFt], s_clip[kMaxFt];
  __shared__ double s_sum[5];
  const int tid = threadIdx.x;
  if (tid < 5) s_sum[tid] = 0.0;
  for (int d = tid; d < a.ft; d += blockDim.x) {
    s_disc[d] = float(pow(double(a.gamma_denoising), double(a.ft - d - 1)));
    if (a.ft > 1) {
      const float t = float(d) / float(a.ft - 1);
      s_clip[d] = a.clip_base + (a.clip_coef - a.clip_base) * (expf(a.clip_rate * t) - 1.f) /
                                    float(exp(double(a.clip_rate)) - 1.0);
    } else {
      s_clip[d] = a.clip_coef;  // the reference divides 0/0 here (NaN); documented deviation
    }
  }
  __syncthreads();
  const float adv_mean = float(a.ws[0]), adv_std = float(a.ws[1]);
  const float lo = a.adv_lo, hi = a.adv_hi;  // -inf / +inf for quantiles 0 / 1 (clamp to [min, max] = no-op)
  const int rows_per_block = 256 / G;
  const int lg = tid % G;
  const int row = blockIdx.x * rows_per_block + tid / G;
  const bool row_ok = row < a.n_rows;
  constexpr int NJ = VEC == 4 ? 1 : 4;  // scalar variant: G = 32, up to 128 elements per row
  float dl[NJ * VEC];
  float new_sum = 0.f, old_sum = 0.f;
  int b = 0, d = 0;
  if (row_ok) {
    if (a.direct) {
      b = row, d = int(a.dinds[row]);
    } else {
      const int64_t idx = a.inds[row];
      b = int(idx / a.ft), d = int(idx % a.ft);
    }
  }
  const StepRow r = a.rows[a.row0 + d];
  const float sd = fmaxf(r.std_train, a.min_std);
  const float inv_var = 1.f / (sd * sd);
  const float log_sd = logf(sd);
  const float w_h = 1.f / float(a.horizon_elems);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e0 = (lg + j * G) * VEC;
    __align__(16) float xe[VEC];
    __align__(16) float xp[VEC];
    __align__(16) float xn[VEC];
    __align__(16) float ol[VEC];
    const bool ok = row_ok && e0 < a.D;
    if (ok) {
      const size_t pe = size_t(row) * a.D + e0;
      const size_t pc = a.direct ? pe : (size_t(b) * (a.ft + 1) + d) * a.D + e0;
      const size_t pl = a.direct ? pe : (size_t(b) * a.ft + d) * a.D + e0;
      const float* nxt = a.direct ? a.x_next + pe : a.chains + pc + a.D;
      if (VEC == 4) {
        *reinterpret_cast<float4*>(xe) = *reinterpret_cast<const float4*>(a.eps + pe);
        *reinterpret_cast<float4*>(xp) = *reinterpret_cast<const float4*>(a.chains + pc);
        *reinterpret_cast<float4*>(xn) = *reinterpret_cast<const float4*>(nxt);
        *reinterpret_cast<float4*>(ol) = *reinterpret_cast<const float4*>(a.old_lp + pl);
      } else {
        xe[0] = a.eps[pe], xp[0] = a.chains[pc], xn[0] = nxt[0], ol[0] = a.old_lp[pl];
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float g = 0.f;
      if (ok && e0 + v < a.horizon_elems) {
        const Posterior p = posterior(r, a.ddim != 0, a.x0_clip, a.eps_clip, xp[v], xe[v]);
        const float diff = xn[v] - p.mu;
        const float lp = -(diff * diff) * 0.5f * inv_var - log_sd - kHalfLog2Pi;
        const float m_lp = (lp >= -5.f && lp <= 2.f) ? 1.f : 0.f;
        new_sum += fminf(fmaxf(lp, -5.f), 2.f);
        old_sum += fminf(fmaxf(ol[v], -5.f), 2.f);
        g = w_h * m_lp * diff * inv_var * p.dmu_deps;
      }
      dl[j * VEC + v] = g;
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    new_sum += __shfl_xor_sync(0xffffffffu, new_sum, o);
    old_sum += __shfl_xor_sync(0xffffffffu, old_sum, o);
  }
  float g_new = 0.f;
  if (row_ok) {
    const float newlp = new_sum * w_h, oldlp = old_sum * w_h;
    float A = a.adv[b];
    if (a.norm_adv) A = (A - adv_mean) / (adv_std + 1e-8f);
    A = fminf(fmaxf(A, lo), hi);
    A *= s_disc[d];
    const float logratio = newlp - oldlp;
    const float ratio = expf(logratio);
    const float c = s_clip[d];
    const float u = -A * ratio;
    const float w = -A * fminf(fmaxf(ratio, 1.f - c), 1.f + c);
    const bool dead = (A > 0.f && ratio > 1.f + c) || (A < 0.f && ratio < 1.f - c);
    const float inv_b = 1.f / float(a.global_rows);
    g_new = dead ? 0.f : -A * ratio * inv_b;
    if (lg == 0) {
      // value loss
      const float v = a.vpred[row], R = a.returns[b];
      float vl, gv;
      if (a.clip_v >= 0.f) {
        const float ov = a.old_values[b];
        const float dv = v - ov;
        const float vc = ov + fminf(fmaxf(dv, -a.clip_v), a.clip_v);
        const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
        const float mc = (dv >= -a.clip_v && dv <= a.clip_v) ? 1.f : 0.f;
        vl = 0.5f * fmaxf(l1, l2);
        gv = l1 > l2 ? (v - R) : (l1 < l2 ? (vc - R) * mc : 0.5f * (v - R) + 0.5f * (vc - R) * mc);
      } else {
        vl = 0.5f * (v - R) * (v - R);
        gv = v - R;
      }
      a.grad_v[row] = gv * inv_b;
      atomicAdd(&s_sum[0], double(fmaxf(u, w)));
      atomicAdd(&s_sum[1], double(vl));
      atomicAdd(&s_sum[2], double((ratio - 1.f) - logratio));
      atomicAdd(&s_sum[3], fabsf(ratio - 1.f) > c ? 1.0 : 0.0);
      atomicAdd(&s_sum[4], double(ratio));
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e0 = (lg + j * G) * VEC;
    if (row_ok && e0 < a.D) {
      const size_t pe = size_t(row) * a.D + e0;
      if (VEC == 4) {
        float4 o4 = make_float4(g_new * dl[0], g_new * dl[1], g_new * dl[2], g_new * dl[3]);
        *reinterpret_cast<float4*>(a.grad_eps + pe) = o4;
      } else {
        a.grad_eps[pe] = g_new * dl[j];
      }
    }
  }
  __syncthreads();
  if (tid < 5) atomicAdd(&a.ws[8 + tid], s_sum[tid]);
}

__global__ void loss_finalize_kernel(const double* __restrict__ ws, int global_rows, float* __restrict__ scalars) {
  const int i = threadIdx.x;
  if (i < 5) scalars[i] = float(ws[8 + i] / double(global_rows));
  if (i == 5) scalars[5] = 0.f;
  if (i == 6) scalars[6] = float(ws[0]);
  if (i == 7) scalars[7] = float(ws[1]);
}


// ---------------------------------------------------------------------------------------------- fused AdamW
// torch.optim.AdamW (decoupled weight decay, bias-corrected, amsgrad off) over ONE flat fp32 segment, with the optional
// clip_grad_norm_ coefficient taken from a device-side squared norm (no host read).  reference
// train_ppo_diffusion_agent.py:360-373, train_ppo_agent.py:34-53.  HBM-bound: 16 B read + 12 B written per element.
__global__ void grad_sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = g[i];
    acc += double(v) * double(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s_part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    acc = lane < int(blockDim.x >> 5) ? s_part[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
  }
}

struct AdamArgs {
  float* p;
  const float* g;
  float *m, *v;
  long long n;
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, max_norm;
  const double* sqnorm;  // nullptr: no clipping
  const int* stop;       // optional device flag: a set flag turns the launch into a no-op (KL early stop, see kl_check_kernel)
  const float* coefs;    // optional device (bc1, bc2_sqrt) written by adam_tick_kernel (step count kept on the device)
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamArgs& a, float coef) {
  g *= coef;
  p *= 1.f - a.lr * a.weight_decay;
  m = m + (g - m) * (1.f - a.beta1);               // torch: exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + (1.f - a.beta2) * g * g;       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
}

// device-resident step count: ++step, bias corrections of that step (skipped once the stop flag is set)
__global__ void adam_tick_kernel(int* __restrict__ step, float* __restrict__ coefs, const int* __restrict__ stop, float beta1,
                                 float beta2) {
  if (stop && *stop) return;
  const int s = ++*step;
  coefs[0] = float(1.0 - pow(double(beta1), s));
  coefs[1] = float(sqrt(1.0 - pow(double(beta2), s)));
}

// after the gradient all-reduce of minibatch k (k = a device counter): keep the diagnostics, raise the stop flag when
// approx_kl exceeds target_kl (reference train_ppo_diffusion_agent.py:376-382: the minibatch that trips the test has
// already been applied; everything after it must not be)
__global__ void kl_check_kernel(const float* __restrict__ scalars, float target_kl, int use_target, int* __restrict__ state,
                                float* __restrict__ history, int max_history) {
  const int k = state[2];
  if (threadIdx.x < 8 && k < max_history) history[k * 8 + threadIdx.x] = scalars[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!state[0] && use_target && scalars[2] > target_kl) state[0] = 1, state[1] = k;
    state[2] = k + 1;
  }
}

__global__ void adamw_kernel(AdamArgs a) {
  if (a.stop && *a.stop) return;
  if (a.coefs) a.bc1 = a.coefs[0], a.bc2_sqrt = a.coefs[1];
  float coef = 1.f;
  if (a.sqnorm) {  // clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6))
    const float total = float(sqrt(*a.sqnorm));
    coef = fminf(a.max_norm / (total + 1e-6f), 1.f);
  }
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long i = i0; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    adamw_one(p.x, g.x, m.x, v.x, a, coef);
    adamw_one(p.y, g.y, m.y, v.y, a, coef);
    adamw_one(p.z, g.z, m.z, v.z, a, coef);
    adamw_one(p.w, g.w, m.w, v.w, a, coef);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
  for (long long i = (n4 << 2) + i0; i < a.n; i += stride) adamw_one(a.p[i], a.g[i], a.m[i], a.v[i], a, coef);
}


// ---------------------------------------------------------------------------------------------- running reward scaling
// RunningRewardScaler (reference dppo/util/reward_scaling.py:42-87, called at train_ppo_diffusion_agent.py:243-247):
// rets_t = r_t + (1 - first_t) gamma rets_{t-1} per env (forward scan, one thread per env, coalesced over envs), batch
// mean / variance of all rets folded into the running statistics (parallel-variance update), rewards divided by
// sqrt(var + eps) and clipped.  float64 like the reference's numpy.  Three phases so that an env-sharded run can
// all-reduce the two batch sums in between (ws[0] = sum, ws[1] = centred sum of squares).
__global__ void reward_scan_kernel(const double* __restrict__ reward, const double* __restrict__ first, int n_steps, int E,
                                   double gamma, double* __restrict__ ret_state, double* __restrict__ rets,
                                   double* __restrict__ ws) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double sum = 0.0;
  if (e < E) {
    double prev = ret_state[e];
    for (int t = 0; t < n_steps; ++t) {
      const size_t i = size_t(t) * E + e;
      prev = reward[i] + (1.0 - first[i]) * gamma * prev;
      rets[i] = prev;
      sum += prev;
    }
    ret_state[e] = prev;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&ws[0], sum);
}

__global__ void reward_centred_sq_kernel(const double* __restrict__ rets, long long n, long long n_global,
                                         double* __restrict__ ws) {
  const double mean = ws[0] / double(n_global);
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double d = rets[i] - mean;
    acc += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&ws[1], acc);
}

// stats = [mean, var, count]; thread 0 of block 0 folds the batch in, every thread then scales its rewards with the
// NEW variance (the reference updates the statistics before scaling)
__global__ void reward_apply_kernel(const double* __restrict__ reward, long long n, long long n_global, double eps,
                                    double cliprew, const double* __restrict__ ws, double* __restrict__ stats,
                                    double* __restrict__ out) {
  const double b_n = double(n_global);
  const double b_mean = ws[0] / b_n, b_var = ws[1] / b_n;
  const double mean = stats[0], var = stats[1], count = stats[2];
  const double delta = b_mean - mean, tot = count + b_n;
  const double m2 = var * count + b_var * b_n + delta * delta * count * b_n / tot;
  const double new_var = m2 / (tot - 1.0);
  const double inv = 1.0 / sqrt(new_var + eps);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = fmin(fmax(reward[i] * inv, -cliprew), cliprew);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // every block has read stats[] above only through registers of its own threads; the update is published through
    // ws[2..4] and copied back by the host-ordered follow-up memcpy, so no block can observe a half-updated state
    double* nxt = const_cast<double*>(ws) + 2;
    nxt[0] = mean + delta * b_n / tot, nxt[1] = new_var, nxt[2] = tot;
  }
}


// ---------------------------------------------------------------------------------------------- split-3 operand packing
// Update-path GEMMs (actor_ft / critic Linear layers, forward + dgrad + wgrad; reference get_logprobs_subsample
// diffusion_vpg.py:398-461, CriticObs.forward critic.py:40-54, loss.backward() train_ppo_diffusion_agent.py:360-364)
// run on the bf16 tensor cores at fp32-grade precision with the same 3-product split as the chain kernel:
// x w ~= x_hi w_hi + x_hi w_lo + x_lo w_hi.  One pass turns an fp32 matrix [M, K] into the bf16 matrix [M, 3 Kp]
// (Kp = K rounded up to 8, zero padded) whose row is [hi | hi | lo] (pattern 0, activations / gradients) or
// [hi | lo | hi] (pattern 1, weights): a single bf16 GEMM with fp32 accumulation over the 3 Kp-long rows is then the
// whole split product.  HBM-bound: 4 B read + 6 B written per element.
// extra_mode 1 appends a column of ones (activations: the bias rides along as one more K column of the weights, and the
// wgrad GEMM's last column is the bias gradient), extra_mode 2 appends extra[r] (weights: the bias vector).
__global__ void split3_pack_kernel(const float* __restrict__ x, long long M, int K, long long ldx, int Kp, int pattern,
                                   int extra_mode, const float* __restrict__ extra, __nv_bfloat16* __restrict__ out) {
  const int q = Kp >> 2;  // groups of 4 columns per row
  const long long total = M * q;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / q;
    const int c = int(i - r * q) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = c + j < K ? x[r * ldx + c + j] : 0.f;
      if (extra_mode && c + j == K) v[j] = extra_mode == 1 ? 1.f : extra[r];
    }
    __align__(8) __nv_bfloat16 hi[4];
    __align__(8) __nv_bfloat16 lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
    __nv_bfloat16* row = out + r * 3 * (long long)Kp + c;
    const uint2 h = *reinterpret_cast<const uint2*>(hi), l = *reinterpret_cast<const uint2*>(lo);
    *reinterpret_cast<uint2*>(row) = h;
    *reinterpret_cast<uint2*>(row + Kp) = pattern == 0 ? h : l;
    *reinterpret_cast<uint2*>(row + 2 * Kp) = pattern == 0 ? l : h;
  }
}

}  // namespace dppo

using namespace dppo;

// SM count of the current device (grid caps of the grid-stride kernels)
static int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return kDefaultSmCount;
  if (!cached[dev]) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : kDefaultSmCount;
  }
  return cached[dev];
}

extern "C" int dppo_split3_pack(const float* x, int64_t rows, int cols, int64_t ldx, int extra_mode, const float* extra,
                                void* out, int pattern, void* stream) {
  if (!x || !out) return set_error("dppo_split3_pack: null argument"), DPPO_ERR_INVALID;
  if (extra_mode < 0 || extra_mode > 2 || (extra_mode == 2 && !extra))
    return set_error("dppo_split3_pack: extra_mode %d", extra_mode), DPPO_ERR_INVALID;
  if (rows < 0 || cols < 1 || ldx < cols || (pattern != 0 && pattern != 1))
    return set_error("dppo_split3_pack: rows=%lld cols=%d ldx=%lld pattern=%d", (long long)rows, cols, (long long)ldx, pattern),
           DPPO_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(out) & 7) return set_error("dppo_split3_pack: output must be 8-byte aligned"), DPPO_ERR_INVALID;
  if (rows == 0) return DPPO_OK;
  const int Kp = (cols + (extra_mode ? 1 : 0) + 7) & ~7;
  const long long total = rows * (Kp / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  split3_pack_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, cols, ldx, Kp, pattern, extra_mode, extra, static_cast<__nv_bfloat16*>(out));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "split3_pack_kernel launch");
}

extern "C" int dppo_reward_scale_f64(const double* reward, const double* first, int n_steps, int n_envs, long long n_global,
                                     double gamma, double epsilon, double cliprew, double* ret_state, double* stats,
                                     double* rets_scratch, double* ws, double* scaled, int phase, void* stream) {
  if (!reward || !first || !ret_state || !stats || !rets_scratch || !ws || !scaled)
    return set_error("dppo_reward_scale_f64: null argument"), DPPO_ERR_INVALID;
  if (n_steps < 1 || n_envs < 1 || n_global < (long long)n_steps * n_envs)
    return set_error("dppo_reward_scale_f64: bad sizes %d x %d of %lld", n_steps, n_envs, n_global), DPPO_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)n_steps * n_envs;
  int blocks = int((n + 255) / 256);
  blocks = blocks > device_sm_count() * 8 ? device_sm_count() * 8 : blocks;
  if (phase == 0) {
    DPPO_CUDA(cudaMemsetAsync(ws, 0, 8 * sizeof(double), st));
    reward_scan_kernel<<<(n_envs + 127) / 128, 128, 0, st>>>(reward, first, n_steps, n_envs, gamma, ret_state, rets_scratch, ws);
  } else if (phase == 1) {
    reward_centred_sq_kernel<<<blocks, 256, 0, st>>>(rets_scratch, n, n_global, ws);
  } else if (phase == 2) {
    reward_apply_kernel<<<blocks, 256, 0, st>>>(reward, n, n_global, epsilon, cliprew, ws, stats, scaled);
    DPPO_CUDA(cudaMemcpyAsync(stats, ws + 2, 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  } else {
    return set_error("dppo_reward_scale_f64: phase %d", phase), DPPO_ERR_INVALID;
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "dppo_reward_scale_f64 launch");
}

extern "C" int dppo_adamw_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                               float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                               float max_grad_norm, void* workspace, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return set_error("dppo_adamw_flat: null argument"), DPPO_ERR_INVALID;
  if (n < 0 || step < 1) return set_error("dppo_adamw_flat: n=%lld step=%d", (long long)n, step), DPPO_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return set_error("dppo_adamw_flat: buffers must be 16-byte aligned"), DPPO_ERR_INVALID;
  if (n == 0) return DPPO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AdamArgs a{};
  a.p = params, a.g = grads, a.m = exp_avg, a.v = exp_avg_sq, a.n = n;
  a.lr = lr, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.weight_decay = weight_decay;
  a.bc1 = float(1.0 - pow(double(beta1), step));
  a.bc2_sqrt = float(sqrt(1.0 - pow(double(beta2), step)));
  a.max_norm = max_grad_norm;
  int blocks = int((n / 4 + 255) / 256);
  blocks = blocks < 1 ? 1 : (blocks > device_sm_count() * 8 ? device_sm_count() * 8 : blocks);
  if (max_grad_norm >= 0.f) {
    if (!workspace) return set_error("dppo_adamw_flat: clipping needs a workspace"), DPPO_ERR_INVALID;
    double* ws = static_cast<double*>(workspace);
    DPPO_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), st));
    grad_sqnorm_kernel<<<blocks, 256, 0, st>>>(grads, n, ws);
    a.sqnorm = ws;
  }
  adamw_kernel<<<blocks, 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "adamw_kernel launch");
}

extern "C" int dppo_adamw_flat_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int* step_state,
                                   const int* stop_flag, float max_grad_norm, void* workspace, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !step_state) return set_error("dppo_adamw_flat_dev: null argument"), DPPO_ERR_INVALID;
  if (n < 0) return set_error("dppo_adamw_flat_dev: n=%lld", (long long)n), DPPO_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return set_error("dppo_adamw_flat_dev: buffers must be 16-byte aligned"), DPPO_ERR_INVALID;
  if (n == 0) return DPPO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AdamArgs a{};
  a.p = params, a.g = grads, a.m = exp_avg, a.v = exp_avg_sq, a.n = n;
  a.lr = lr, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.weight_decay = weight_decay;
  a.max_norm = max_grad_norm, a.stop = stop_flag, a.coefs = reinterpret_cast<const float*>(step_state + 2);
  adam_tick_kernel<<<1, 1, 0, st>>>(step_state, reinterpret_cast<float*>(step_state + 2), stop_flag, beta1, beta2);
  int blocks = int((n / 4 + 255) / 256);
  blocks = blocks < 1 ? 1 : (blocks > device_sm_count() * 8 ? device_sm_count() * 8 : blocks);
  if (max_grad_norm >= 0.f) {
    if (!workspace) return set_error("dppo_adamw_flat_dev: clipping needs a workspace"), DPPO_ERR_INVALID;
    double* ws = static_cast<double*>(workspace);
    DPPO_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), st));
    grad_sqnorm_kernel<<<blocks, 256, 0, st>>>(grads, n, ws);
    a.sqnorm = ws;
  }
  adamw_kernel<<<blocks, 256, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "adamw_kernel launch");
}

extern "C" int dppo_kl_check(const float* scalars, float target_kl, int use_target, int* state, float* history, int max_history,
                             void* stream) {
  if (!scalars || !state || !history || max_history < 1) return set_error("dppo_kl_check: bad argument"), DPPO_ERR_INVALID;
  kl_check_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(scalars, target_kl, use_target, state, history, max_history);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "kl_check_kernel launch");
}

extern "C" int dppo_gae_f64(const double* reward, const double* terminated, const double* values,
                            const double* next_value, int n_steps, int n_envs, double gamma, double lam, double scale,
                            double* adv, double* ret, void* stream) {
  if (!reward || !terminated || !values || !next_value || !adv || !ret)
    return set_error("dppo_gae_f64: null argument"), DPPO_ERR_INVALID;
  if (n_steps < 0 || n_envs < 0) return set_error("dppo_gae_f64: negative size"), DPPO_ERR_INVALID;
  if (n_steps == 0 || n_envs == 0) return DPPO_OK;
  gae_kernel<<<(n_envs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(reward, terminated, values, next_value,
                                                                                   n_steps, n_envs, gamma, lam, scale, adv, ret);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "gae_kernel launch");
}

extern "C" int dppo_logprob_rows(dppo_ctx* ctx, const float* eps, const float* x_prev, const float* x_next,
                                 const int64_t* dinds, int n_rows, float* logp, float* dlogp_deps, void* stream) {
  if (!ctx || !eps || !x_prev || !x_next || !dinds || !logp) return set_error("dppo_logprob_rows: null argument"), DPPO_ERR_INVALID;
  if (n_rows <= 0) return n_rows == 0 ? DPPO_OK : (set_error("dppo_logprob_rows: n_rows < 0"), DPPO_ERR_INVALID);
  const long long total = (long long)n_rows * ctx->sample_dim;
  logprob_rows_kernel<<<unsigned((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      eps, x_prev, x_next, dinds, ctx->d_rows, ctx->S - ctx->ft, ctx->use_ddim, ctx->x0_clip, ctx->eps_clip,
      ctx->min_logprob_std, ctx->sample_dim, total, logp, dlogp_deps);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "logprob_rows_kernel launch");
}

static int loss_impl(dppo_ctx* ctx, LossArgs a, const int64_t* stat_inds, const dppo_loss_hp* hp, float* scalars,
                     void* workspace, void* stream, const char* who) {
  const int D = ctx->sample_dim;
  // the row kernels hold one sample (Ta * Da elements) in one group of <= 32 lanes x 4 elements
  if (D < 1 || D > 128) return set_error("%s: Ta * Da = %d outside [1, 128]", who, D), DPPO_ERR_UNSUPPORTED;
  if (a.n_rows < 0 || a.global_rows < 1 || a.n_rows > a.global_rows)
    return set_error("%s: bad row counts %d of %d", who, a.n_rows, a.global_rows), DPPO_ERR_INVALID;
  if (hp->ft_denoising_steps != ctx->ft || hp->horizon_steps * hp->action_dim != D)
    return set_error("%s: hyper-parameters disagree with the context", who), DPPO_ERR_INVALID;
  if (hp->ft_denoising_steps > 128 || hp->ft_denoising_steps < 1)
    return set_error("%s: ft_denoising_steps %d outside [1,128]", who, hp->ft_denoising_steps), DPPO_ERR_INVALID;
  if (hp->reward_horizon < 1 || hp->reward_horizon > hp->horizon_steps)
    return set_error("%s: reward_horizon %d", who, hp->reward_horizon), DPPO_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* ws = static_cast<double*>(workspace);
  DPPO_CUDA(cudaMemsetAsync(ws, 0, 16 * sizeof(double), st));
  DPPO_CUDA(cudaMemsetAsync(reinterpret_cast<int*>(ws + 6), 0x7F, sizeof(int), st));      // min key: +3.4e38
  DPPO_CUDA(cudaMemsetAsync(reinterpret_cast<int*>(ws + 6) + 1, 0x80, sizeof(int), st));  // max key: below every float
  {
    int blocks = (a.global_rows + 255) / 256;
    blocks = blocks > 2 * device_sm_count() ? 2 * device_sm_count() : blocks;
    adv_partial_kernel<<<blocks, 256, 0, st>>>(a.adv, stat_inds, a.global_rows, ctx->ft, ws);
    adv_finalize_kernel<<<1, 1, 0, st>>>(a.global_rows, ws);
  }
  a.rows = ctx->d_rows, a.row0 = ctx->S - ctx->ft;
  a.D = D, a.ft = ctx->ft, a.ddim = ctx->use_ddim;
  a.horizon_elems = hp->reward_horizon * hp->action_dim, a.norm_adv = hp->norm_adv;
  a.x0_clip = ctx->x0_clip, a.eps_clip = ctx->eps_clip, a.min_std = ctx->min_logprob_std;
  a.gamma_denoising = hp->gamma_denoising, a.clip_coef = hp->clip_ploss_coef, a.clip_base = hp->clip_ploss_coef_base;
  a.clip_rate = hp->clip_ploss_coef_rate, a.clip_v = hp->clip_vloss_coef, a.adv_lo = hp->adv_clip_lo, a.adv_hi = hp->adv_clip_hi;
  a.ws = ws;
  const int n_rows = a.n_rows;
  if (n_rows > 0) {
    if (D % 4 == 0) {
      const int v4 = D / 4;
      if (v4 <= 4) {
        ppo_loss_kernel<4, 4><<<(n_rows + 63) / 64, 256, 0, st>>>(a);
      } else if (v4 <= 8) {
        ppo_loss_kernel<8, 4><<<(n_rows + 31) / 32, 256, 0, st>>>(a);
      } else if (v4 <= 16) {
        ppo_loss_kernel<16, 4><<<(n_rows + 15) / 16, 256, 0, st>>>(a);
      } else {
        ppo_loss_kernel<32, 4><<<(n_rows + 7) / 8, 256, 0, st>>>(a);
      }
    } else {
      ppo_loss_kernel<32, 1><<<(n_rows + 7) / 8, 256, 0, st>>>(a);
    }
  }
  loss_finalize_kernel<<<1, 32, 0, st>>>(ws, a.global_rows, scalars);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, who);
}

extern "C" int dppo_ppo_loss_fwd_bwd(dppo_ctx* ctx, const float* chains, const float* old_logprobs,
                                     const float* returns, const float* old_values, const float* advantages,
                                     const int64_t* inds_all, int row_begin, const float* eps, const float* vpred,
                                     int n_rows, int global_rows, const dppo_loss_hp* hp, float* grad_eps,
                                     float* grad_vpred, float* scalars, void* workspace, void* stream) {
  if (!ctx || !chains || !old_logprobs || !returns || !old_values || !advantages || !inds_all || !eps || !vpred || !hp ||
      !grad_eps || !grad_vpred || !scalars || !workspace)
    return set_error("dppo_ppo_loss_fwd_bwd: null argument"), DPPO_ERR_INVALID;
  if (row_begin < 0 || row_begin + n_rows > global_rows)
    return set_error("dppo_ppo_loss_fwd_bwd: bad row range [%d, %d) of %d", row_begin, row_begin + n_rows, global_rows),
           DPPO_ERR_INVALID;
  LossArgs a{};
  a.chains = chains, a.old_lp = old_logprobs, a.returns = returns, a.old_values = old_values, a.adv = advantages;
  a.eps = eps, a.vpred = vpred, a.inds = inds_all + row_begin, a.direct = 0;
  a.n_rows = n_rows, a.global_rows = global_rows, a.grad_eps = grad_eps, a.grad_v = grad_vpred;
  return loss_impl(ctx, a, inds_all, hp, scalars, workspace, stream, "dppo_ppo_loss_fwd_bwd");
}

extern "C" int dppo_ppo_loss_rows(dppo_ctx* ctx, const float* x_prev, const float* x_next, const float* old_logprobs,
                                  const float* returns, const float* old_values, const float* advantages,
                                  const int64_t* denoising_inds, const float* eps, const float* vpred, int n_rows,
                                  const dppo_loss_hp* hp, float* grad_eps, float* grad_vpred, float* scalars,
                                  void* workspace, void* stream) {
  if (!ctx || !x_prev || !x_next || !old_logprobs || !returns || !old_values || !advantages || !denoising_inds || !eps ||
      !vpred || !hp || !grad_eps || !grad_vpred || !scalars || !workspace)
    return set_error("dppo_ppo_loss_rows: null argument"), DPPO_ERR_INVALID;
  LossArgs a{};
  a.chains = x_prev, a.x_next = x_next, a.old_lp = old_logprobs, a.returns = returns, a.old_values = old_values;
  a.adv = advantages, a.dinds = denoising_inds, a.eps = eps, a.vpred = vpred, a.direct = 1;
  a.n_rows = n_rows, a.global_rows = n_rows, a.grad_eps = grad_eps, a.grad_v = grad_vpred;
  return loss_impl(ctx, a, nullptr, hp, scalars, workspace, stream, "dppo_ppo_loss_rows");
}
