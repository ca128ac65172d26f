// The PPO minibatch as a program of hand-written sm_100a kernels (dppo_update_*, include/dppo_b200.h).
//
// Replaces, for DiffusionMLP actors and residual-MLP critics, what the reference does with torch autograd:
//   PPODiffusion.loss -> get_logprobs_subsample -> actor_ft(x, t, cond)   dppo/model/diffusion/diffusion_vpg.py:398-461
//   DiffusionMLP.forward / ResidualMLP                                     dppo/model/diffusion/mlp_diffusion.py:218-250,
//                                                                          dppo/model/common/mlp.py:84-154
//   CriticObs.forward                                                      dppo/model/common/critic.py:40-54
//   loss.backward()                                                        dppo/agent/finetune/train_ppo_diffusion_agent.py:360-364
//
// Forward = gather (pack_rows) -> [cond_mlp] -> layer 0 -> residual blocks -> output layer, every Linear one launch of
// ugemm_rows_kernel whose epilogue writes the next layer's operand images; backward = per layer one wgrad launch
// (ugemm_wgrad_kernel, straight into the caller's gradient tensors) and one dgrad launch (ugemm_rows_kernel on W^T tiles,
// epilogue multiplies by the activation derivative and adds the residual gradient).  LayerNorm runs as a row kernel
// before / after the GEMMs.  The time-embedding MLP depends on the denoising index only: it is folded, exactly like in the
// chain kernel, into a per-index bias row TB[d] = b0 + W0[:, time] temb(t_d) that enters layer 0 through a one-hot(d)
// block of input columns; the wgrad of those columns IS the per-index sum of the layer-0 output gradient, from which a
// one-block kernel back-propagates through the time MLP.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "internal.h"
#include "update_gemm.h"

namespace dppo {

struct LayerW {
  const float* W = nullptr;  // fp32 [N][ld]
  float* dW = nullptr;
  const float* b = nullptr;
  float* db = nullptr;
  int N = 0, K = 0;
  int64_t ld = 0;
  uint8_t* Bf = nullptr;  // forward tiles (rows = N output features, contraction K)
  uint8_t* Bd = nullptr;  // dgrad tiles (rows = K input features, contraction N), null when no dgrad is needed
  int bd_rows = 0, bd_col0 = 0;  // dgrad restricted to input columns [bd_col0, bd_col0 + bd_rows)
  int nt_f = 0, nt_d = 0;        // output-tile widths (NTILE) the forward / dgrad tiles are packed for, see call_ntile
};

struct BlockW {
  LayerW l1, l2;
  const float *g1 = nullptr, *be1 = nullptr, *g2 = nullptr, *be2 = nullptr;
  float *dg1 = nullptr, *dbe1 = nullptr, *dg2 = nullptr, *dbe2 = nullptr;
};

// one residual MLP (actor trunk or critic) with its activations
struct ResMlp {
  int K0 = 0, H = 0, nb = 0, Dout = 0, act = 0, ln = 0;
  LayerW L0, out;
  std::vector<BlockW> blk;
  // activations (sized for max_rows)
  uint8_t* X0 = nullptr;  // input images [R][K0]; may be shared (critic: the observation images)
  int FC0 = 0;
  bool own_x0 = true;
  std::vector<float*> Hf, Yf;       // fp32 pre-activations h_b (block inputs), y_b (l1 outputs)
  std::vector<uint8_t*> A, A1;      // operand images act(norm(h_b)), act(norm(y_b))
  std::vector<float*> st1, st2;     // LayerNorm statistics (mean, rstd) of h_b, y_b
  std::vector<float*> sf1, sf2, sb1, sb2;  // row sums accumulated by the GEMM epilogues (forward: of h_b / y_b, backward: of dz)
  float* stat_arena = nullptr;      // sf* / sb* live here, zeroed once per minibatch
  size_t stat_arena_bytes = 0;
  std::vector<uint32_t*> MH, MY;    // ReLU nets: (h_b > 0), (y_b > 0) as bit masks instead of the fp32 pre-activations
  int fmode = 1;                    // layout of the fp32 side tensors: 1 = tiled (update_gemm.h), 0 = row-major (LayerNorm nets)
  bool relu_bits = false;
  uint8_t* HL = nullptr;            // raw h_nb images (input of the output layer)
  float* OUT = nullptr;             // fp32 [R][Dout]
  // backward
  uint8_t* GOUT = nullptr;          // images of d loss / d OUT
  float* DHf = nullptr;
  uint8_t* DHop = nullptr;
  uint8_t* G1op = nullptr;
  float* DZf = nullptr;
  int FCH = 0, FCout = 0;
};

}  // namespace dppo

using namespace dppo;

struct dppo_update {
  dppo_ctx* ctx = nullptr;
  int max_rows = 0;
  cudaEvent_t actor_done = nullptr;  // optional: recorded on the stream behind the actor backward (dppo_update_set_actor_event)
  int sm_count = kDefaultSmCount;
  ResMlp actor, critic;
  // actor extras: time MLP, cond_mlp, assembled layer 0
  const float *tw1 = nullptr, *tb1 = nullptr, *tw2 = nullptr, *tb2 = nullptr, *W0 = nullptr, *b0 = nullptr;
  float *dtw1 = nullptr, *dtb1 = nullptr, *dtw2 = nullptr, *dtb2 = nullptr, *dW0 = nullptr, *db0 = nullptr;
  LayerW c0, c1;  // cond_mlp (Linear -> Mish -> Linear)
  float* YC0f = nullptr;
  uint8_t *AC0 = nullptr, *GC1 = nullptr, *GC0 = nullptr, *OBS = nullptr;
  int FCobs = 0;
  float *W0p = nullptr, *dW0p = nullptr;  // [H][K0p]: [cond | x | one-hot(d) -> TB] columns
  int K0p = 0;
  int* d_ts = nullptr;  // [ft] timestep of every denoising index
  float *t_emb = nullptr, *t_hpre = nullptr, *t_temb = nullptr, *TB = nullptr, *t_dtemb = nullptr, *t_dpre = nullptr;
  float *t_G = nullptr, *t_hid = nullptr;
  // weight-pack job table
  PackWJob* d_jobs = nullptr;
  int n_jobs_actor = 0, n_jobs = 0;
  long long max_job = 0;
  bool bound = false;
  // loss buffers
  float *geps = nullptr, *gv = nullptr;
  int n_rows = 0;  // rows of the last forward
  std::vector<void*> allocs;
};

namespace dppo {

static int dev_alloc(dppo_update* u, void** p, size_t bytes, bool zero = true) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) return cuda_fail(e, "dppo_update allocation");
  u->allocs.push_back(*p);
  if (zero) {
    e = cudaMemset(*p, 0, bytes ? bytes : 16);
    if (e != cudaSuccess) return cuda_fail(e, "dppo_update memset");
  }
  return DPPO_OK;
}
#define UALLOC(ptr, bytes)                                                              \
  do {                                                                                  \
    int rc__ = dev_alloc(u, reinterpret_cast<void**>(&(ptr)), (bytes));                 \
    if (rc__ != DPPO_OK) return rc__;                                                   \
  } while (0)

static int alloc_resmlp(dppo_update* u, ResMlp& m, int R) {
  m.FCH = opmat_chunks(m.H), m.FCout = opmat_chunks(m.Dout);
  m.FC0 = opmat_chunks(m.K0);
  if (m.own_x0) UALLOC(m.X0, opmat_bytes(R, m.K0));
  m.Hf.assign(m.nb, nullptr), m.Yf.assign(m.nb, nullptr), m.A.assign(m.nb, nullptr), m.A1.assign(m.nb, nullptr);
  m.st1.assign(m.nb, nullptr), m.st2.assign(m.nb, nullptr);
  m.MH.assign(m.nb, nullptr), m.MY.assign(m.nb, nullptr);
  m.fmode = 1;
  m.relu_bits = !m.ln && m.act == kUActRelu && m.H % 32 == 0;
  const size_t fbytes = f32_tiled_floats(R, m.H) * 4;
  for (int b = 0; b < m.nb; ++b) {
    UALLOC(m.Hf[b], fbytes);
    if (m.relu_bits) {
      UALLOC(m.MH[b], mask_words_total(R, m.H) * 4);
      UALLOC(m.MY[b], mask_words_total(R, m.H) * 4);
    } else {
      UALLOC(m.Yf[b], fbytes);
    }
    UALLOC(m.A[b], opmat_bytes(R, m.H));
    UALLOC(m.A1[b], opmat_bytes(R, m.H));
    if (m.ln) {
      UALLOC(m.st1[b], size_t(R) * 8);
      UALLOC(m.st2[b], size_t(R) * 8);
    }
  }
  m.sf1.assign(m.nb, nullptr), m.sf2.assign(m.nb, nullptr), m.sb1.assign(m.nb, nullptr), m.sb2.assign(m.nb, nullptr);
  if (m.ln) {
    m.stat_arena_bytes = size_t(4) * m.nb * R * 8;
    UALLOC(m.stat_arena, m.stat_arena_bytes);
    for (int b = 0; b < m.nb; ++b) {
      m.sf1[b] = m.stat_arena + (size_t(4) * b + 0) * R * 2, m.sf2[b] = m.stat_arena + (size_t(4) * b + 1) * R * 2;
      m.sb1[b] = m.stat_arena + (size_t(4) * b + 2) * R * 2, m.sb2[b] = m.stat_arena + (size_t(4) * b + 3) * R * 2;
    }
  }
  UALLOC(m.HL, opmat_bytes(R, m.H));
  UALLOC(m.OUT, size_t(R) * m.Dout * 4);
  UALLOC(m.GOUT, opmat_bytes(R, m.Dout));
  UALLOC(m.DHf, fbytes);
  UALLOC(m.DHop, opmat_bytes(R, m.H));
  UALLOC(m.G1op, opmat_bytes(R, m.H));
  if (m.ln) UALLOC(m.DZf, fbytes);
  return DPPO_OK;
}

// Output-tile width of one row-GEMM CALL on `rows` rows, given the width the weights are packed for.  A work unit is one
// 128-row x NTILE tile and costs its own ingest (128 x K of A plus NTILE x K of W, hi + lo); a shard of a minibatch
// (strong scaling: 2200 rows = 18 row tiles) with 256-wide tiles leaves half of the SMs without a unit, so the call runs
// narrower tiles of the same packed weights (RowGemmArgs.NTP) while twice the units still fit in one wave.
static int call_ntile(int packed, int N, int rows, int sm_count) {
  int nt = packed;
  if (sm_count <= 0) return nt;
  const int rt = (rows + 127) / 128;
  while (nt >= 128 && nt % 128 == 0 && rt * ((N + nt / 2 - 1) / (nt / 2)) <= sm_count) nt /= 2;
  return nt;
}

static int alloc_layer_tiles(dppo_update* u, LayerW& L, bool dgrad) {
  L.nt_f = row_gemm_ntile(L.N);
  UALLOC(L.Bf, packed_weight_bytes(L.N, L.K, L.nt_f));
  if (dgrad) {
    if (L.bd_rows == 0) L.bd_rows = L.K, L.bd_col0 = 0;
    L.nt_d = row_gemm_ntile(L.bd_rows);
    UALLOC(L.Bd, packed_weight_bytes(L.bd_rows, L.N, L.nt_d));
  }
  return DPPO_OK;
}

static void add_jobs(std::vector<PackWJob>& jobs, const LayerW& L) {
  PackWJob j{};
  j.W = L.W, j.s_row = L.ld, j.s_col = 1, j.rows = L.N, j.K = L.K, j.NTILE = L.nt_f;
  j.NT = (L.N + j.NTILE - 1) / j.NTILE, j.KC = (L.K + 63) / 64, j.out = L.Bf;
  jobs.push_back(j);
  if (L.Bd) {
    PackWJob d{};
    d.W = L.W + L.bd_col0, d.s_row = 1, d.s_col = L.ld, d.rows = L.bd_rows, d.K = L.N, d.NTILE = L.nt_d;
    d.NT = (d.rows + d.NTILE - 1) / d.NTILE, d.KC = (d.K + 63) / 64, d.out = L.Bd;
    jobs.push_back(d);
  }
}

// ------------------------------------------------------------------------------------------------ time-embedding MLP
__device__ __forceinline__ float mish_exact_u(float x) {
  const float sp = x > 20.f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float mish_grad_exact_u(float x) {
  const float sp = x > 20.f ? x : log1pf(expf(x));
  const float th = tanhf(sp);
  const float sg = 1.f / (1.f + expf(-x));
  return th + x * sg * (1.f - th * th);
}

// grid = (ft, ceil(H / 128)) blocks of 128 threads (every block recomputes the tiny MLP of its denoising index and owns
// 128 features of TB; blockIdx.y == 0 also stores the intermediates): SinusoidalPosEmb -> Linear -> Mish -> Linear (reference modules.py:14-27,
// mlp_diffusion.py:191-196) for the timestep of denoising index d, then TB[d][f] = b0[f] + W0[f, D:D+td] . temb
// (same arithmetic as time_bias_table_kernel in pack.cu; the intermediates are kept for the backward kernel)
__global__ void time_table_fwd_kernel(const float* __restrict__ tw1, const float* __restrict__ tb1,
                                      const float* __restrict__ tw2, const float* __restrict__ tb2,
                                      const float* __restrict__ W0, const float* __restrict__ b0, int td, int D, int in0,
                                      int H, const int* __restrict__ ts, float* __restrict__ emb, float* __restrict__ hpre,
                                      float* __restrict__ temb, float* __restrict__ TB) {
  __shared__ float s_emb[64], s_hid[128], s_out[64];
  const int d = blockIdx.x, t = ts[d];
  const int half = td / 2;
  if (threadIdx.x < half) {
    const float rate = float(log(10000.0) / double(half - 1));
    const float freq = expf(float(threadIdx.x) * -rate);
    const float ph = float(t) * freq;
    s_emb[threadIdx.x] = sinf(ph);
    s_emb[threadIdx.x + half] = cosf(ph);
  }
  __syncthreads();
  const bool keep = blockIdx.y == 0;
  if (keep && threadIdx.x < td) emb[d * td + threadIdx.x] = s_emb[threadIdx.x];
  if (threadIdx.x < 2 * td) {
    float acc = 0.f;
    for (int j = 0; j < td; ++j) acc += s_emb[j] * tw1[threadIdx.x * td + j];
    acc += tb1[threadIdx.x];
    if (keep) hpre[d * 2 * td + threadIdx.x] = acc;
    s_hid[threadIdx.x] = mish_exact_u(acc);
  }
  __syncthreads();
  if (threadIdx.x < td) {
    float acc = 0.f;
    for (int j = 0; j < 2 * td; ++j) acc += s_hid[j] * tw2[threadIdx.x * 2 * td + j];
    acc += tb2[threadIdx.x];
    s_out[threadIdx.x] = acc;
    if (keep) temb[d * td + threadIdx.x] = acc;
  }
  __syncthreads();
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  if (f < H) {
    float acc = 0.f;
    const float* w = W0 + size_t(f) * in0 + D;
    for (int j = 0; j < td; ++j) acc += w[j] * s_out[j];
    TB[size_t(d) * H + f] = acc + b0[f];
  }
}

// W0p[f][c]: c < Dc -> W0[f][D + td + c] (conditioning columns); c < Dc + D -> W0[f][c - Dc] (sample columns);
// c < Dc + D + ft -> TB[c - Dc - D][f] (one-hot(d) columns); else 0
__global__ void assemble_w0_kernel(const float* __restrict__ W0, const float* __restrict__ TB, int H, int in0, int D, int td,
                                   int Dc, int ft, int K0p, float* __restrict__ W0p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * K0p) return;
  const int f = i / K0p, c = i - f * K0p;
  float v = 0.f;
  if (c < Dc) v = W0[size_t(f) * in0 + D + td + c];
  else if (c < Dc + D) v = W0[size_t(f) * in0 + c - Dc];
  else if (c < Dc + D + ft) v = TB[size_t(c - Dc - D) * H + f];
  W0p[i] = v;
}

__global__ void scatter_dw0_kernel(const float* __restrict__ dW0p, int H, int in0, int D, int td, int Dc, int K0p,
                                   float* __restrict__ dW0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * (Dc + D)) return;
  const int f = i / (Dc + D), c = i - f * (Dc + D);
  const float g = dW0p[size_t(f) * K0p + c];
  if (c < Dc) dW0[size_t(f) * in0 + D + td + c] += g;
  else dW0[size_t(f) * in0 + c - Dc] += g;
}

// G[d][f] = dW0p[f][Dc + D + d] is the gradient w.r.t. TB[d][f] = b0[f] + W0[f, D:D+td] . temb[d]; three small kernels
// carry it back through the time MLP.  All outputs are accumulated (+=) into the parameter gradients.
// (a) grid = (ft, ceil(td / 8)) blocks of 8 warps: dtemb[d][j] = sum_f G[d][f] W0[f][D + j] (one warp per j, lanes over f);
//     G is also copied to a contiguous [ft][H] scratch for (c) by the blocks with blockIdx.y == 0
__global__ void __launch_bounds__(256) time_bwd_a_kernel(const float* __restrict__ dW0p, int K0p, int gcol, int H, int td, int D,
                                                         int in0, const float* __restrict__ W0, float* __restrict__ Gs,
                                                         float* __restrict__ dtemb) {
  extern __shared__ float s_g[];
  const int d = blockIdx.x;
  for (int f = threadIdx.x; f < H; f += blockDim.x) {
    const float v = dW0p[size_t(f) * K0p + gcol + d];
    s_g[f] = v;
    if (blockIdx.y == 0) Gs[size_t(d) * H + f] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = blockIdx.y * nw + warp; j < td; j += nw * gridDim.y) {
    float s = 0.f;
#pragma unroll 8
    for (int f = lane; f < H; f += 32) s += s_g[f] * W0[size_t(f) * in0 + D + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dtemb[d * td + j] = s;
  }
}

// (b) one block: the two Linear layers of the time MLP (ft x td x 2td work)
__global__ void __launch_bounds__(256) time_bwd_b_kernel(int ft, int td, const float* __restrict__ tw2,
                                                         const float* __restrict__ emb, const float* __restrict__ hpre,
                                                         const float* __restrict__ dtemb, float* __restrict__ hid,
                                                         float* __restrict__ dpre, float* __restrict__ dtw1,
                                                         float* __restrict__ dtb1, float* __restrict__ dtw2,
                                                         float* __restrict__ dtb2) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < ft * 2 * td; i += nt) hid[i] = mish_exact_u(hpre[i]);
  __threadfence_block();
  __syncthreads();
  for (int j = tid; j < td; j += nt) {
    float s = 0.f;
    for (int d = 0; d < ft; ++d) s += dtemb[d * td + j];
    dtb2[j] += s;
  }
  for (int i = tid; i < td * 2 * td; i += nt) {
    const int j = i / (2 * td), k = i - j * 2 * td;
    float s = 0.f;
    for (int d = 0; d < ft; ++d) s += dtemb[d * td + j] * hid[d * 2 * td + k];
    dtw2[i] += s;
  }
  for (int i = tid; i < ft * 2 * td; i += nt) {
    const int d = i / (2 * td), k = i - d * 2 * td;
    float s = 0.f;
    for (int j = 0; j < td; ++j) s += dtemb[d * td + j] * tw2[j * 2 * td + k];
    dpre[i] = s * mish_grad_exact_u(hpre[i]);
  }
  __threadfence_block();
  __syncthreads();
  for (int k = tid; k < 2 * td; k += nt) {
    float s = 0.f;
    for (int d = 0; d < ft; ++d) s += dpre[d * 2 * td + k];
    dtb1[k] += s;
  }
  for (int i = tid; i < 2 * td * td; i += nt) {
    const int k = i / td, j = i - k * td;
    float s = 0.f;
    for (int d = 0; d < ft; ++d) s += dpre[d * 2 * td + k] * emb[d * td + j];
    dtw1[i] += s;
  }
}

// (c) one thread per (feature, column j <= td): dW0[f][D + j] += sum_d G[d][f] temb[d][j]; j == td: db0[f] += sum_d G[d][f]
__global__ void time_bwd_c_kernel(const float* __restrict__ Gs, int ft, int H, int td, int D, int in0,
                                  const float* __restrict__ temb, float* __restrict__ dW0, float* __restrict__ db0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = i / (td + 1), j = i - f * (td + 1);
  if (f >= H) return;
  float s = 0.f;
  if (j == td) {
    for (int d = 0; d < ft; ++d) s += Gs[size_t(d) * H + f];
    db0[f] += s;
  } else {
    for (int d = 0; d < ft; ++d) s += Gs[size_t(d) * H + f] * temb[d * td + j];
    dW0[size_t(f) * in0 + D + j] += s;
  }
}

// ------------------------------------------------------------------------------------------------ residual MLP program
static int g_plan_sm_count = 0;  // SM count of the device, set by dppo_update_create (0: calls keep the packed tile width)
static RowGemmArgs gemm_args(const uint8_t* A, int FCa, const LayerW& L, bool dgrad, int R) {
  RowGemmArgs g{};
  g.A = A, g.FCa = FCa, g.R = R;
  if (!dgrad) {
    g.B = L.Bf, g.N = L.N, g.KC = (L.K + 63) / 64;
  } else {
    g.B = L.Bd, g.N = L.bd_rows, g.KC = (L.N + 63) / 64;
  }
  g.NTP = dgrad ? L.nt_d : L.nt_f;
  g.NTILE = call_ntile(g.NTP, g.N, R, g_plan_sm_count);
  g.NT = (g.N + g.NTILE - 1) / g.NTILE;
  return g;
}

static WgradArgs wgrad_args(const uint8_t* G, int FCg, const uint8_t* X, int FCx, const LayerW& L, int R) {
  WgradArgs w{};
  w.G = G, w.FCg = FCg, w.X = X, w.FCx = FCx, w.R = R, w.N_out = L.N, w.K_in = L.K, w.dW = L.dW, w.ld_dw = int(L.ld), w.db = L.db;
  return w;
}

#define URUN(call)                       \
  do {                                   \
    int rc__ = (call);                   \
    if (rc__ != DPPO_OK) return rc__;    \
  } while (0)

// layer 0 .. output layer; the input images m.X0 are already in place
static int resmlp_forward(const dppo_update* u, ResMlp& m, int R, float* out_override, cudaStream_t st) {
  const int sm = u->sm_count;
  const float eps = 1e-6f;
  if (m.ln) DPPO_CUDA(cudaMemsetAsync(m.stat_arena, 0, m.stat_arena_bytes, st));
  {
    RowGemmArgs g = gemm_args(m.X0, m.FC0, m.L0, false, R);
    g.bias = m.L0.b;
    g.out_f32 = m.Hf[0], g.ld_out = m.H, g.out_mode = m.fmode;
    if (!m.ln) g.out_op = m.A[0], g.FCo = m.FCH, g.act_out = m.act;
    if (m.relu_bits) g.mask_out = m.MH[0], g.mask_words = m.H / 32;
    if (m.ln) g.stat_out = m.sf1[0];
    URUN(launch_row_gemm(g, sm, st));
    if (m.ln) URUN(launch_ln_fwd_tiled(m.Hf[0], m.sf1[0], R, m.H, m.blk[0].g1, m.blk[0].be1, eps, m.act, m.st1[0], m.A[0], m.FCH, st));
  }
  for (int b = 0; b < m.nb; ++b) {
    const BlockW& B = m.blk[b];
    {
      RowGemmArgs g = gemm_args(m.A[b], m.FCH, B.l1, false, R);
      g.bias = B.l1.b;
      if (m.relu_bits) g.mask_out = m.MY[b], g.mask_words = m.H / 32;
      else g.out_f32 = m.Yf[b], g.ld_out = m.H, g.out_mode = m.fmode;
      if (!m.ln) g.out_op = m.A1[b], g.FCo = m.FCH, g.act_out = m.act;
      if (m.ln) g.stat_out = m.sf2[b];
      URUN(launch_row_gemm(g, sm, st));
      if (m.ln) URUN(launch_ln_fwd_tiled(m.Yf[b], m.sf2[b], R, m.H, B.g2, B.be2, eps, m.act, m.st2[b], m.A1[b], m.FCH, st));
    }
    {
      const bool last = b + 1 == m.nb;
      RowGemmArgs g = gemm_args(m.A1[b], m.FCH, B.l2, false, R);
      g.bias = B.l2.b;
      g.res = m.Hf[b], g.ld_res = m.H, g.res_mode = m.fmode;
      if (last) {
        g.out_op = m.HL, g.FCo = m.FCH, g.act_out = kUActNone;  // no activation between the last block and the output layer
      } else {
        g.out_f32 = m.Hf[b + 1], g.ld_out = m.H, g.out_mode = m.fmode;
        if (!m.ln) g.out_op = m.A[b + 1], g.FCo = m.FCH, g.act_out = m.act;
        if (m.relu_bits) g.mask_out = m.MH[b + 1], g.mask_words = m.H / 32;
        if (m.ln) g.stat_out = m.sf1[b + 1];
      }
      URUN(launch_row_gemm(g, sm, st));
      if (!last && m.ln)
        URUN(launch_ln_fwd_tiled(m.Hf[b + 1], m.sf1[b + 1], R, m.H, m.blk[b + 1].g1, m.blk[b + 1].be1, eps, m.act, m.st1[b + 1], m.A[b + 1], m.FCH, st));
    }
  }
  {
    RowGemmArgs g = gemm_args(m.HL, m.FCH, m.out, false, R);
    g.bias = m.out.b;
    g.out_f32 = out_override ? out_override : m.OUT, g.ld_out = m.Dout;
    URUN(launch_row_gemm(g, sm, st));
  }
  return DPPO_OK;
}

// m.GOUT holds the images of d loss / d OUT.  Leaves d loss / d h_0 in m.DHop (and m.DHf).
static int resmlp_backward(const dppo_update* u, ResMlp& m, int R, bool need_dh0_f32, cudaStream_t st) {
  const int sm = u->sm_count;
  (void)need_dh0_f32;
  URUN(launch_wgrad(wgrad_args(m.GOUT, m.FCout, m.HL, m.FCH, m.out, R), sm, st));
  {
    RowGemmArgs g = gemm_args(m.GOUT, m.FCout, m.out, true, R);
    g.out_f32 = m.DHf, g.ld_out = m.H, g.out_mode = m.fmode;
    g.out_op = m.DHop, g.FCo = m.FCH, g.act_out = kUActNone;
    URUN(launch_row_gemm(g, sm, st));
  }
  for (int b = m.nb - 1; b >= 0; --b) {
    const BlockW& B = m.blk[b];
    URUN(launch_wgrad(wgrad_args(m.DHop, m.FCH, m.A1[b], m.FCH, B.l2, R), sm, st));
    {
      RowGemmArgs g = gemm_args(m.DHop, m.FCH, B.l2, true, R);
      if (m.relu_bits) g.mask_in = m.MY[b], g.mask_words = m.H / 32;
      else g.pre = m.Yf[b], g.ld_pre = m.H, g.pre_mode = m.fmode, g.act_grad = m.act;
      if (m.ln) {
        g.ln_stats = m.st2[b], g.ln_g = B.g2, g.ln_b = B.be2;
        g.out_f32 = m.DZf, g.ld_out = m.H, g.out_mode = m.fmode, g.stat_out = m.sb2[b];
      } else {
        g.out_op = m.G1op, g.FCo = m.FCH, g.act_out = kUActNone;
      }
      URUN(launch_row_gemm(g, sm, st));
      if (m.ln)
        URUN(launch_ln_bwd_tiled(m.DZf, m.sb2[b], m.Yf[b], m.st2[b], B.g2, R, m.H, nullptr, nullptr, m.G1op, m.FCH, B.dg2, B.dbe2, st));
    }
    URUN(launch_wgrad(wgrad_args(m.G1op, m.FCH, m.A[b], m.FCH, B.l1, R), sm, st));
    {
      RowGemmArgs g = gemm_args(m.G1op, m.FCH, B.l1, true, R);
      if (m.relu_bits) g.mask_in = m.MH[b], g.mask_words = m.H / 32;
      else g.pre = m.Hf[b], g.ld_pre = m.H, g.pre_mode = m.fmode, g.act_grad = m.act;
      if (m.ln) {
        g.ln_stats = m.st1[b], g.ln_g = B.g1, g.ln_b = B.be1;
        g.out_f32 = m.DZf, g.ld_out = m.H, g.out_mode = m.fmode, g.stat_out = m.sb1[b];
      } else {
        g.res = m.DHf, g.ld_res = m.H, g.res_mode = m.fmode;
        g.out_f32 = m.DHf, g.ld_out = m.H, g.out_mode = m.fmode;
        g.out_op = m.DHop, g.FCo = m.FCH, g.act_out = kUActNone;
      }
      URUN(launch_row_gemm(g, sm, st));
      if (m.ln)
        URUN(launch_ln_bwd_tiled(m.DZf, m.sb1[b], m.Hf[b], m.st1[b], B.g1, R, m.H, m.DHf, m.DHf, m.DHop, m.FCH, B.dg1, B.dbe1, st));
    }
  }
  URUN(launch_wgrad(wgrad_args(m.DHop, m.FCH, m.X0, m.FC0, m.L0, R), sm, st));
  return DPPO_OK;
}

static int act_code(int dppo_act) { return dppo_act == DPPO_ACT_RELU ? kUActRelu : kUActMish; }

}  // namespace dppo

// ================================================================================================== C ABI
extern "C" int dppo_update_destroy(dppo_update* u) {
  if (!u) return DPPO_OK;
  for (void* p : u->allocs) cudaFree(p);
  delete u;
  return DPPO_OK;
}

extern "C" int dppo_update_create(dppo_update** out, dppo_ctx* ctx, const dppo_resmlp_desc* critic, int max_rows) {
  if (!out || !ctx || !critic) return set_error("dppo_update_create: null argument"), DPPO_ERR_INVALID;
  if (ctx->kind != 0) return set_error("dppo_update_create: the tensor-core update path covers DiffusionMLP actors"), DPPO_ERR_UNSUPPORTED;
  if (max_rows < 1) return set_error("dppo_update_create: max_rows=%d", max_rows), DPPO_ERR_INVALID;
  const MlpGeom& g = ctx->g;

  if (critic->hidden_dim % 8 || critic->hidden_dim < 16 || critic->n_blocks < 1 || critic->out_dim < 1 || critic->in_dim < 1 ||
      critic->in_dim != g.Dc_in)
    return set_error("dppo_update_create: critic geometry (%d -> %d x %d -> %d) unsupported", critic->in_dim, critic->hidden_dim,
                     critic->n_blocks, critic->out_dim), DPPO_ERR_UNSUPPORTED;
  if (g.CH && (g.CO % 64 || g.CH % 64))
    return set_error("dppo_update_create: cond_mlp widths (%d, %d) must be multiples of 64", g.CH, g.CO), DPPO_ERR_UNSUPPORTED;
  if ((g.ln && g.act != DPPO_ACT_MISH) || (critic->use_layernorm && critic->activation != DPPO_ACT_MISH))
    return set_error("dppo_update_create: LayerNorm nets are built for the Mish activation"), DPPO_ERR_UNSUPPORTED;
  if (g.H % 64 || critic->hidden_dim % 64)
    return set_error("dppo_update_create: hidden widths (%d, %d) must be multiples of 64", g.H, critic->hidden_dim), DPPO_ERR_UNSUPPORTED;
  if (ctx->ft < 1 || ctx->ft > 128) return set_error("dppo_update_create: ft_denoising_steps %d", ctx->ft), DPPO_ERR_UNSUPPORTED;
  DPPO_CUDA(cudaSetDevice(ctx->device));
  dppo_update* u = new dppo_update();
  u->ctx = ctx, u->max_rows = max_rows, u->sm_count = ctx->sm_count;
  g_plan_sm_count = ctx->sm_count;
  const int R = max_rows;
  int rc = [&]() -> int {
    ResMlp& a = u->actor;
    a.K0 = g.Dc + g.D + ctx->ft, a.H = g.H, a.nb = g.nb, a.Dout = g.D, a.act = act_code(g.act), a.ln = g.ln;
    a.blk.resize(g.nb);
    URUN(alloc_resmlp(u, a, R));
    u->K0p = (a.K0 + 63) / 64 * 64;
    UALLOC(u->W0p, size_t(g.H) * u->K0p * 4);
    UALLOC(u->dW0p, size_t(g.H) * u->K0p * 4);
    UALLOC(u->d_ts, size_t(ctx->ft) * 4);
    {
      std::vector<int> ts(ctx->ft);
      for (int d = 0; d < ctx->ft; ++d) ts[d] = ctx->rows[ctx->S - ctx->ft + d].t;
      DPPO_CUDA(cudaMemcpy(u->d_ts, ts.data(), ts.size() * 4, cudaMemcpyHostToDevice));
    }
    UALLOC(u->t_emb, size_t(ctx->ft) * g.td * 4);
    UALLOC(u->t_hpre, size_t(ctx->ft) * 2 * g.td * 4);
    UALLOC(u->t_temb, size_t(ctx->ft) * g.td * 4);
    UALLOC(u->TB, size_t(ctx->ft) * g.H * 4);
    UALLOC(u->t_dtemb, size_t(ctx->ft) * g.td * 4);
    UALLOC(u->t_dpre, size_t(ctx->ft) * 2 * g.td * 4);
    UALLOC(u->t_hid, size_t(ctx->ft) * 2 * g.td * 4);
    UALLOC(u->t_G, size_t(ctx->ft) * g.H * 4);
    // layer 0 runs on the assembled matrix; its wgrad lands in dW0p
    a.L0.W = u->W0p, a.L0.dW = u->dW0p, a.L0.N = g.H, a.L0.K = u->K0p, a.L0.ld = u->K0p;
    if (g.CH) a.L0.bd_rows = g.CO, a.L0.bd_col0 = 0;
    URUN(alloc_layer_tiles(u, a.L0, g.CH != 0));
    for (int b = 0; b < g.nb; ++b) {
      a.blk[b].l1.N = a.blk[b].l1.K = a.blk[b].l2.N = a.blk[b].l2.K = g.H;
      a.blk[b].l1.ld = a.blk[b].l2.ld = g.H;
      URUN(alloc_layer_tiles(u, a.blk[b].l1, true));
      URUN(alloc_layer_tiles(u, a.blk[b].l2, true));
    }
    a.out.N = g.D, a.out.K = g.H, a.out.ld = g.H;
    URUN(alloc_layer_tiles(u, a.out, true));
    // observation images: cond_mlp input and critic input
    u->FCobs = opmat_chunks(g.Dc_in);
    UALLOC(u->OBS, opmat_bytes(R, g.Dc_in));
    if (g.CH) {
      u->c0.N = g.CH, u->c0.K = g.Dc_in, u->c0.ld = g.Dc_in;
      u->c1.N = g.CO, u->c1.K = g.CH, u->c1.ld = g.CH;
      URUN(alloc_layer_tiles(u, u->c0, false));
      URUN(alloc_layer_tiles(u, u->c1, true));
      UALLOC(u->YC0f, f32_tiled_floats(R, g.CH) * 4);
      UALLOC(u->AC0, opmat_bytes(R, g.CH));
      UALLOC(u->GC1, opmat_bytes(R, g.CO));
      UALLOC(u->GC0, opmat_bytes(R, g.CH));
    }
    ResMlp& c = u->critic;
    c.K0 = critic->in_dim, c.H = critic->hidden_dim, c.nb = critic->n_blocks, c.Dout = critic->out_dim;
    c.act = act_code(critic->activation), c.ln = critic->use_layernorm;
    c.blk.resize(c.nb);
    c.own_x0 = false, c.X0 = u->OBS;
    URUN(alloc_resmlp(u, c, R));
    c.L0.N = c.H, c.L0.K = c.K0, c.L0.ld = c.K0;
    URUN(alloc_layer_tiles(u, c.L0, false));
    for (int b = 0; b < c.nb; ++b) {
      c.blk[b].l1.N = c.blk[b].l1.K = c.blk[b].l2.N = c.blk[b].l2.K = c.H;
      c.blk[b].l1.ld = c.blk[b].l2.ld = c.H;
      URUN(alloc_layer_tiles(u, c.blk[b].l1, true));
      URUN(alloc_layer_tiles(u, c.blk[b].l2, true));
    }
    c.out.N = c.Dout, c.out.K = c.H, c.out.ld = c.H;
    URUN(alloc_layer_tiles(u, c.out, true));
    UALLOC(u->geps, size_t(R) * g.D * 4);
    UALLOC(u->gv, size_t(R) * c.Dout * 4);
    UALLOC(u->d_jobs, sizeof(PackWJob) * 64);
    return DPPO_OK;
  }();
  if (rc != DPPO_OK) {
    dppo_update_destroy(u);
    return rc;
  }
  *out = u;
  return DPPO_OK;
}

// parameter / gradient pointers: actor_ft in dppo_pack_mlp order, critic as layers.0.{weight,bias}, per block
// l1.{weight,bias}, l2.{weight,bias} [, norm1.{weight,bias}, norm2.{weight,bias}], layers.last.{weight,bias}
extern "C" int dppo_update_bind(dppo_update* u, const float* const* ap, float* const* ag, int n_actor,
                                const float* const* cp, float* const* cg, int n_critic) {
  if (!u || !ap || !ag || !cp || !cg) return set_error("dppo_update_bind: null argument"), DPPO_ERR_INVALID;
  const MlpGeom& g = u->ctx->g;
  const int expect_a = 4 + (g.CH ? 4 : 0) + 2 + g.nb * (g.ln ? 8 : 4) + 2;
  const int expect_c = 2 + u->critic.nb * (u->critic.ln ? 8 : 4) + 2;
  if (n_actor != expect_a || n_critic != expect_c)
    return set_error("dppo_update_bind: expected %d actor / %d critic tensors, got %d / %d", expect_a, expect_c, n_actor, n_critic),
           DPPO_ERR_INVALID;
  for (int i = 0; i < n_actor; ++i)
    if (!ap[i] || !ag[i]) return set_error("dppo_update_bind: actor tensor %d is null", i), DPPO_ERR_INVALID;
  for (int i = 0; i < n_critic; ++i)
    if (!cp[i] || !cg[i]) return set_error("dppo_update_bind: critic tensor %d is null", i), DPPO_ERR_INVALID;
  int i = 0;
  u->tw1 = ap[i], u->dtw1 = ag[i++], u->tb1 = ap[i], u->dtb1 = ag[i++];
  u->tw2 = ap[i], u->dtw2 = ag[i++], u->tb2 = ap[i], u->dtb2 = ag[i++];
  if (g.CH) {
    u->c0.W = ap[i], u->c0.dW = ag[i++], u->c0.b = ap[i], u->c0.db = ag[i++];
    u->c1.W = ap[i], u->c1.dW = ag[i++], u->c1.b = ap[i], u->c1.db = ag[i++];
  }
  u->W0 = ap[i], u->dW0 = ag[i++], u->b0 = ap[i], u->db0 = ag[i++];
  auto bind_blocks = [](ResMlp& m, const float* const* p, float* const* gr, int& k) {
    for (int b = 0; b < m.nb; ++b) {
      BlockW& B = m.blk[b];
      B.l1.W = p[k], B.l1.dW = gr[k++], B.l1.b = p[k], B.l1.db = gr[k++];
      B.l2.W = p[k], B.l2.dW = gr[k++], B.l2.b = p[k], B.l2.db = gr[k++];
      if (m.ln) {
        B.g1 = p[k], B.dg1 = gr[k++], B.be1 = p[k], B.dbe1 = gr[k++];
        B.g2 = p[k], B.dg2 = gr[k++], B.be2 = p[k], B.dbe2 = gr[k++];
      }
    }
    m.out.W = p[k], m.out.dW = gr[k++], m.out.b = p[k], m.out.db = gr[k++];
  };
  bind_blocks(u->actor, ap, ag, i);
  int k = 0;
  u->critic.L0.W = cp[k], u->critic.L0.dW = cg[k++], u->critic.L0.b = cp[k], u->critic.L0.db = cg[k++];
  bind_blocks(u->critic, cp, cg, k);
  // job table of the per-minibatch weight repack (actor jobs first)
  std::vector<PackWJob> jobs;
  add_jobs(jobs, u->actor.L0);
  for (auto& B : u->actor.blk) add_jobs(jobs, B.l1), add_jobs(jobs, B.l2);
  add_jobs(jobs, u->actor.out);
  if (g.CH) add_jobs(jobs, u->c0), add_jobs(jobs, u->c1);
  u->n_jobs_actor = int(jobs.size());
  add_jobs(jobs, u->critic.L0);
  for (auto& B : u->critic.blk) add_jobs(jobs, B.l1), add_jobs(jobs, B.l2);
  add_jobs(jobs, u->critic.out);
  u->n_jobs = int(jobs.size());
  if (u->n_jobs > 64) return set_error("dppo_update_bind: %d weight matrices exceed the job table", u->n_jobs), DPPO_ERR_UNSUPPORTED;
  u->max_job = 0;
  for (auto& j : jobs) {
    const long long t = (long long)j.NT * j.KC * j.NTILE * 8;
    if (t > u->max_job) u->max_job = t;
  }
  DPPO_CUDA(cudaMemcpy(u->d_jobs, jobs.data(), sizeof(PackWJob) * jobs.size(), cudaMemcpyHostToDevice));
  u->bound = true;
  return DPPO_OK;
}

namespace dppo {

// the (b, d) rows of the minibatch slice -> operand images of layer 0's input and of the observation
static int gather_inputs(dppo_update* u, const dppo_update_batch* bt, cudaStream_t st) {
  const dppo_ctx* ctx = u->ctx;
  const MlpGeom& g = ctx->g;
  const int R = bt->n_rows;
  const bool gather = bt->inds_all != nullptr;
  PackArgs p{};
  p.inds = gather ? bt->inds_all + bt->row_begin : nullptr;
  p.dinds = gather ? nullptr : bt->denoising_inds;
  p.ft = ctx->ft, p.chain_stride = int64_t(ctx->ft + 1) * g.D, p.chain_d = g.D;
  p.R = R;
  // observation images (critic input, cond_mlp input)
  p.n_seg = 1;
  p.seg[0] = PackSeg{bt->obs, g.Dc_in, gather ? 1 : 0, g.Dc_in, 0};
  p.FCp = u->FCobs, p.out = u->OBS;
  URUN(launch_pack_rows(p, st));
  // layer-0 input: [cond | x | one-hot(d)]
  int n = 0;
  if (!g.CH) p.seg[n++] = PackSeg{bt->obs, g.Dc_in, gather ? 1 : 0, g.Dc_in, 0};
  p.seg[n++] = PackSeg{bt->chains, g.D, gather ? 2 : 0, g.D, g.Dc};
  p.seg[n++] = PackSeg{nullptr, 0, 0, ctx->ft, g.Dc + g.D};
  p.n_seg = n;
  p.FCp = u->actor.FC0, p.out = u->actor.X0;
  URUN(launch_pack_rows(p, st));
  return DPPO_OK;
}

static int actor_forward(dppo_update* u, int R, float* eps_out, cudaStream_t st) {
  const dppo_ctx* ctx = u->ctx;
  const MlpGeom& g = ctx->g;
  const int in0 = g.D + g.td + g.Dc;
  time_table_fwd_kernel<<<dim3(ctx->ft, (g.H + 127) / 128), 128, 0, st>>>(u->tw1, u->tb1, u->tw2, u->tb2, u->W0, u->b0, g.td, g.D, in0, g.H, u->d_ts,
                                                 u->t_emb, u->t_hpre, u->t_temb, u->TB);
  assemble_w0_kernel<<<(g.H * u->K0p + 255) / 256, 256, 0, st>>>(u->W0, u->TB, g.H, in0, g.D, g.td, g.Dc, ctx->ft, u->K0p, u->W0p);
  URUN(launch_pack_weights(u->d_jobs, u->n_jobs_actor, u->max_job, st));
  if (g.CH) {
    {
      RowGemmArgs a = gemm_args(u->OBS, u->FCobs, u->c0, false, R);
      a.bias = u->c0.b;
      a.out_f32 = u->YC0f, a.ld_out = g.CH, a.out_mode = 1;
      a.out_op = u->AC0, a.FCo = opmat_chunks(g.CH), a.act_out = u->actor.act;
      URUN(launch_row_gemm(a, u->sm_count, st));
    }
    {
      RowGemmArgs a = gemm_args(u->AC0, opmat_chunks(g.CH), u->c1, false, R);
      a.bias = u->c1.b;
      a.out_op = u->actor.X0, a.FCo = u->actor.FC0, a.op_col0 = 0, a.act_out = kUActNone;
      URUN(launch_row_gemm(a, u->sm_count, st));
    }
  }
  URUN(resmlp_forward(u, u->actor, R, eps_out, st));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "dppo_update actor forward");
}

static int actor_backward(dppo_update* u, int R, cudaStream_t st) {
  const dppo_ctx* ctx = u->ctx;
  const MlpGeom& g = ctx->g;
  const int in0 = g.D + g.td + g.Dc;
  DPPO_CUDA(cudaMemsetAsync(u->dW0p, 0, size_t(g.H) * u->K0p * 4, st));
  URUN(resmlp_backward(u, u->actor, R, g.CH != 0, st));
  if (g.CH) {
    {
      RowGemmArgs a = gemm_args(u->actor.DHop, u->actor.FCH, u->actor.L0, true, R);  // d cond = dh0 . W0[:, cond]
      a.out_op = u->GC1, a.FCo = opmat_chunks(g.CO), a.act_out = kUActNone;
      URUN(launch_row_gemm(a, u->sm_count, st));
    }
    URUN(launch_wgrad(wgrad_args(u->GC1, opmat_chunks(g.CO), u->AC0, opmat_chunks(g.CH), u->c1, R), u->sm_count, st));
    {
      RowGemmArgs a = gemm_args(u->GC1, opmat_chunks(g.CO), u->c1, true, R);
      a.pre = u->YC0f, a.ld_pre = g.CH, a.pre_mode = 1, a.act_grad = u->actor.act;
      a.out_op = u->GC0, a.FCo = opmat_chunks(g.CH), a.act_out = kUActNone;
      URUN(launch_row_gemm(a, u->sm_count, st));
    }
    URUN(launch_wgrad(wgrad_args(u->GC0, opmat_chunks(g.CH), u->OBS, u->FCobs, u->c0, R), u->sm_count, st));
  }
  scatter_dw0_kernel<<<(g.H * (g.Dc + g.D) + 255) / 256, 256, 0, st>>>(u->dW0p, g.H, in0, g.D, g.td, g.Dc, u->K0p, u->dW0);
  time_bwd_a_kernel<<<dim3(ctx->ft, (g.td + 7) / 8), 256, size_t(g.H) * 4, st>>>(u->dW0p, u->K0p, g.Dc + g.D, g.H, g.td, g.D, in0, u->W0, u->t_G,
                                                           u->t_dtemb);
  time_bwd_b_kernel<<<1, 256, 0, st>>>(ctx->ft, g.td, u->tw2, u->t_emb, u->t_hpre, u->t_dtemb, u->t_hid, u->t_dpre, u->dtw1,
                                       u->dtb1, u->dtw2, u->dtb2);
  time_bwd_c_kernel<<<(g.H * (g.td + 1) + 127) / 128, 128, 0, st>>>(u->t_G, ctx->ft, g.H, g.td, g.D, in0, u->t_temb, u->dW0, u->db0);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DPPO_OK : cuda_fail(e, "dppo_update actor backward");
}

static int pack_grad(const float* src, int ld, int width, int R, const float* scale, float scale_imm, uint8_t* out, int FCp,
                     cudaStream_t st) {
  PackArgs p{};
  p.n_seg = 1;
  p.seg[0] = PackSeg{src, ld, 0, width, 0};
  p.R = R, p.FCp = FCp, p.out = out, p.scale = scale, p.scale_imm = scale_imm, p.ft = 1;
  return launch_pack_rows(p, st);
}

}  // namespace dppo

extern "C" int dppo_update_forward(dppo_update* u, const dppo_update_batch* bt, float* eps_out, float* vpred_out,
                                   void* stream) {
  if (!u || !bt) return set_error("dppo_update_forward: null argument"), DPPO_ERR_INVALID;
  if (!u->bound) return set_error("dppo_update_forward: parameters not bound (dppo_update_bind)"), DPPO_ERR_STATE;
  if (bt->n_rows < 0 || bt->n_rows > u->max_rows)
    return set_error("dppo_update_forward: %d rows outside [0, %d]", bt->n_rows, u->max_rows), DPPO_ERR_INVALID;
  if (!bt->obs || !bt->chains || (!bt->inds_all && !bt->denoising_inds))
    return set_error("dppo_update_forward: missing input"), DPPO_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  u->n_rows = bt->n_rows;
  if (bt->n_rows == 0) return DPPO_OK;
  URUN(gather_inputs(u, bt, st));
  URUN(actor_forward(u, bt->n_rows, eps_out, st));
  URUN(launch_pack_weights(u->d_jobs + u->n_jobs_actor, u->n_jobs - u->n_jobs_actor, u->max_job, st));
  URUN(resmlp_forward(u, u->critic, bt->n_rows, vpred_out, st));
  return DPPO_OK;
}

// critic(obs) for n_rows observation rows [n_rows, cond_dim] (no gather): the value pass of the rollout buffer
// (train_ppo_diffusion_agent.py:197-206, 259-263) through the same row-GEMM kernels as the minibatch forward
extern "C" int dppo_update_values(dppo_update* u, const float* obs, int n_rows, float* vpred_out, void* stream) {
  if (!u || !obs || !vpred_out) return set_error("dppo_update_values: null argument"), DPPO_ERR_INVALID;
  if (!u->bound) return set_error("dppo_update_values: parameters not bound (dppo_update_bind)"), DPPO_ERR_STATE;
  if (n_rows < 0 || n_rows > u->max_rows)
    return set_error("dppo_update_values: %d rows outside [0, %d]", n_rows, u->max_rows), DPPO_ERR_INVALID;
  if (n_rows == 0) return DPPO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const MlpGeom& g = u->ctx->g;
  u->n_rows = 0;  // the saved activations no longer belong to a minibatch: a backward call needs a new forward
  PackArgs p{};
  p.ft = u->ctx->ft, p.chain_stride = int64_t(u->ctx->ft + 1) * g.D, p.chain_d = g.D;
  p.R = n_rows;
  p.n_seg = 1;
  p.seg[0] = PackSeg{obs, g.Dc_in, 0, g.Dc_in, 0};
  p.FCp = u->FCobs, p.out = u->OBS;
  URUN(launch_pack_rows(p, st));
  URUN(launch_pack_weights(u->d_jobs + u->n_jobs_actor, u->n_jobs - u->n_jobs_actor, u->max_job, st));
  URUN(resmlp_forward(u, u->critic, n_rows, vpred_out, st));
  return DPPO_OK;
}

extern "C" int dppo_update_backward(dppo_update* u, const float* grad_eps, const float* grad_vpred, const float* scale_pg,
                                    const float* scale_v, float vf_coef, int with_actor, int with_critic, void* stream) {
  if (!u) return set_error("dppo_update_backward: null argument"), DPPO_ERR_INVALID;
  if (!u->bound) return set_error("dppo_update_backward: parameters not bound"), DPPO_ERR_STATE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int R = u->n_rows;
  if (R == 0) return DPPO_OK;
  if (with_actor) {
    if (!grad_eps) return set_error("dppo_update_backward: grad_eps is null"), DPPO_ERR_INVALID;
    URUN(pack_grad(grad_eps, u->actor.Dout, u->actor.Dout, R, scale_pg, 1.f, u->actor.GOUT, u->actor.FCout, st));
    URUN(actor_backward(u, R, st));
    // the actor gradients are final here: a caller overlaps their all-reduce with the critic backward below
    if (u->actor_done) DPPO_CUDA(cudaEventRecord(u->actor_done, st));
  }
  if (with_critic) {
    if (!grad_vpred) return set_error("dppo_update_backward: grad_vpred is null"), DPPO_ERR_INVALID;
    URUN(pack_grad(grad_vpred, u->critic.Dout, u->critic.Dout, R, scale_v, vf_coef, u->critic.GOUT, u->critic.FCout, st));
    URUN(resmlp_backward(u, u->critic, R, false, st));
  }
  return DPPO_OK;
}

extern "C" int dppo_update_minibatch(dppo_update* u, const dppo_update_batch* bt, const dppo_loss_hp* hp, float vf_coef,
                                     int with_actor, float* scalars, void* workspace, void* stream) {
  if (!u || !bt || !hp || !scalars || !workspace) return set_error("dppo_update_minibatch: null argument"), DPPO_ERR_INVALID;
  if (u->critic.Dout != 1) return set_error("dppo_update_minibatch: the critic must have one output"), DPPO_ERR_INVALID;
  const float *eps = u->actor.OUT, *vpred = u->critic.OUT;
  URUN(dppo_update_forward(u, bt, nullptr, nullptr, stream));
  if (bt->inds_all) {
    URUN(dppo_ppo_loss_fwd_bwd(u->ctx, bt->chains, bt->old_logprobs, bt->returns, bt->old_values, bt->advantages, bt->inds_all,
                               bt->row_begin, eps, vpred, bt->n_rows, bt->global_rows, hp, u->geps, u->gv, scalars, workspace,
                               stream));
  } else {
    URUN(dppo_ppo_loss_rows(u->ctx, bt->chains, bt->x_next, bt->old_logprobs, bt->returns, bt->old_values, bt->advantages,
                            bt->denoising_inds, eps, vpred, bt->n_rows, hp, u->geps, u->gv, scalars, workspace, stream));
  }
  // loss = pg_loss + vf_coef * v_loss (the entropy term of a fixed-eta chain is a constant): grad_v is scaled by vf_coef
  return dppo_update_backward(u, u->geps, u->gv, nullptr, nullptr, vf_coef, with_actor, 1, stream);
}

// `event`: a cudaEvent_t (or NULL to switch it off) that every later dppo_update_backward / dppo_update_minibatch records
// behind the actor backward, before the critic backward is enqueued
extern "C" int dppo_update_set_actor_event(dppo_update* u, void* event) {
  if (!u) return set_error("dppo_update_set_actor_event: null argument"), DPPO_ERR_INVALID;
  u->actor_done = static_cast<cudaEvent_t>(event);
  return DPPO_OK;
}

// stream-ordered zero fill (the flat gradient buffer before a minibatch) without a framework elementwise kernel
extern "C" int dppo_memset_zero(void* ptr, size_t bytes, void* stream) {
  if (!ptr && bytes) return set_error("dppo_memset_zero: null pointer"), DPPO_ERR_INVALID;
  if (bytes) DPPO_CUDA(cudaMemsetAsync(ptr, 0, bytes, static_cast<cudaStream_t>(stream)));
  return DPPO_OK;
}

extern "C" int dppo_update_buffers(dppo_update* u, float** eps, float** vpred, float** grad_eps, float** grad_vpred) {
  if (!u) return set_error("dppo_update_buffers: null argument"), DPPO_ERR_INVALID;
  if (eps) *eps = u->actor.OUT;
  if (vpred) *vpred = u->critic.OUT;
  if (grad_eps) *grad_eps = u->geps;
  if (grad_vpred) *grad_vpred = u->gv;
  return DPPO_OK;
}

// ---- bring-up / unit tests (not in the public header) ------------------------------------------------------------
extern "C" int dppo_debug_set_mn_desc(unsigned lbo_bytes, unsigned sbo_bytes) {
  set_mn_desc_override(lbo_bytes, sbo_bytes);
  return DPPO_OK;
}

// out[R][N] = epilogue(x[R][K] . W^T) with W[N][K] (transposed = 0) or W[K][N] read as W^T (transposed = 1, the dgrad form)
extern "C" int dppo_debug_linear(const float* x, int R, int K, const float* W, int N, int transposed, const float* bias,
                                 const float* pre, int act_grad, const float* res, float* out_f32, int act_out,
                                 float* out_act_f32, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sm = kDefaultSmCount;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  uint8_t *xi = nullptr, *bt = nullptr, *oi = nullptr;
  const int NTILE = row_gemm_ntile(N);
  DPPO_CUDA(cudaMalloc(&xi, opmat_bytes(R, K)));
  DPPO_CUDA(cudaMalloc(&bt, packed_weight_bytes(N, K, NTILE)));
  DPPO_CUDA(cudaMalloc(&oi, opmat_bytes(R, N)));
  DPPO_CUDA(cudaMemsetAsync(oi, 0, opmat_bytes(R, N), st));
  PackArgs p{};
  p.n_seg = 1, p.seg[0] = PackSeg{x, K, 0, K, 0}, p.R = R, p.FCp = opmat_chunks(K), p.out = xi, p.ft = 1;
  int rc = launch_pack_rows(p, st);
  if (rc == DPPO_OK) rc = launch_pack_weight(W, transposed ? 1 : K, transposed ? N : 1, N, K, NTILE, bt, st);
  RowGemmArgs g{};
  g.A = xi, g.FCa = opmat_chunks(K), g.B = bt, g.R = R, g.KC = (K + 63) / 64, g.N = N, g.NTILE = NTILE, g.NT = (N + NTILE - 1) / NTILE;
  g.bias = bias, g.pre = pre, g.ld_pre = N, g.act_grad = act_grad, g.res = res, g.ld_res = N, g.out_f32 = out_f32, g.ld_out = N;
  g.out_op = out_act_f32 ? oi : nullptr, g.FCo = opmat_chunks(N), g.act_out = act_out;
  if (rc == DPPO_OK) rc = launch_row_gemm(g, sm, st);
  if (rc == DPPO_OK && out_act_f32) rc = launch_unpack_rows(oi, opmat_chunks(N), R, N, out_act_f32, N, st);
  cudaStreamSynchronize(st);
  cudaFree(xi), cudaFree(bt), cudaFree(oi);
  return rc;
}

// times `reps` launches of one row GEMM (R x K) . (N x K)^T on resident operands; flags: 1 = fp32 output (tiled),
// 2 = operand-image output (ReLU; 64: no activation, 128: Mish), 4 = fp32 residual input (tiled), 8 = bit-mask output,
// 16 = bias, 32 = bit-mask input, 256 = mish'(fp32 pre) input, 512 = force the generic epilogue.  prof (optional): [grid][16] role counters
// of the last launch.  Returns the average milliseconds per launch in *ms.
extern "C" int dppo_debug_gemm_bench(int R, int K, int N, int flags, int ntile_cap, int max_stages, int reps, float* ms,
                                     unsigned long long* prof, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sm = kDefaultSmCount, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  set_row_gemm_tuning(ntile_cap, max_stages);
  const int NTILE = row_gemm_ntile(N);
  uint8_t *xi = nullptr, *bt = nullptr, *oi = nullptr;
  float *of = nullptr, *rf = nullptr;
  uint32_t* mk = nullptr;
  DPPO_CUDA(cudaMalloc(&xi, opmat_bytes(R, K)));
  DPPO_CUDA(cudaMalloc(&bt, packed_weight_bytes(N, K, NTILE)));
  DPPO_CUDA(cudaMalloc(&oi, opmat_bytes(R, N)));
  DPPO_CUDA(cudaMalloc(&of, f32_tiled_floats(R, N) * 4));
  DPPO_CUDA(cudaMalloc(&rf, f32_tiled_floats(R, N) * 4));
  DPPO_CUDA(cudaMalloc(&mk, mask_words_total(R, N) * 4));
  DPPO_CUDA(cudaMemsetAsync(xi, 0, opmat_bytes(R, K), st));
  DPPO_CUDA(cudaMemsetAsync(bt, 0, packed_weight_bytes(N, K, NTILE), st));
  DPPO_CUDA(cudaMemsetAsync(rf, 0, f32_tiled_floats(R, N) * 4, st));
  RowGemmArgs g{};
  g.A = xi, g.FCa = opmat_chunks(K), g.B = bt, g.R = R, g.KC = (K + 63) / 64, g.N = N, g.NTILE = NTILE, g.NT = (N + NTILE - 1) / NTILE;
  if (flags & 1) g.out_f32 = of, g.ld_out = N, g.out_mode = 1;
  if (flags & 2) g.out_op = oi, g.FCo = opmat_chunks(N), g.act_out = (flags & 64) ? kUActNone : ((flags & 128) ? kUActMish : kUActRelu);
  if (flags & 4) g.res = rf, g.ld_res = N, g.res_mode = 1;
  if (flags & 8) g.mask_out = mk, g.mask_words = (N + 31) / 32;
  if (flags & 16) g.bias = rf;
  if (flags & 32) g.mask_in = mk, g.mask_words = (N + 31) / 32;
  if (flags & 256) g.pre = rf, g.ld_pre = N, g.pre_mode = 1, g.act_grad = kUActMish;
  set_row_gemm_fast_epilogue(!(flags & 512));
  int rc = launch_row_gemm(g, sm, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps && rc == DPPO_OK; ++i) {
    g.prof = (i == reps - 1) ? prof : nullptr;
    rc = launch_row_gemm(g, sm, st);
  }
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float t = 0.f;
  cudaEventElapsedTime(&t, e0, e1);
  if (ms) *ms = t / reps;
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(xi), cudaFree(bt), cudaFree(oi), cudaFree(of), cudaFree(rf), cudaFree(mk);
  set_row_gemm_tuning(0, 0);
  set_row_gemm_fast_epilogue(true);
  return rc;
}

// dW[N][K] += g[R][N]^T x[R][K], db[N] += column sums of g
extern "C" int dppo_debug_wgrad(const float* gm, const float* x, int R, int N, int K, float* dW, float* db, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sm = kDefaultSmCount;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  uint8_t *gi = nullptr, *xi = nullptr;
  DPPO_CUDA(cudaMalloc(&gi, opmat_bytes(R, N)));
  DPPO_CUDA(cudaMalloc(&xi, opmat_bytes(R, K)));
  PackArgs p{};
  p.n_seg = 1, p.R = R, p.ft = 1;
  p.seg[0] = PackSeg{gm, N, 0, N, 0}, p.FCp = opmat_chunks(N), p.out = gi;
  int rc = launch_pack_rows(p, st);
  p.seg[0] = PackSeg{x, K, 0, K, 0}, p.FCp = opmat_chunks(K), p.out = xi;
  if (rc == DPPO_OK) rc = launch_pack_rows(p, st);
  WgradArgs w{};
  w.G = gi, w.FCg = opmat_chunks(N), w.X = xi, w.FCx = opmat_chunks(K), w.R = R, w.N_out = N, w.K_in = K, w.dW = dW, w.ld_dw = K, w.db = db;
  if (rc == DPPO_OK) rc = launch_wgrad(w, sm, st);
  cudaStreamSynchronize(st);
  cudaFree(gi), cudaFree(xi);
  return rc;
}
