// Latency-bound denoise chain for a handful of environments (E <= 64): weights-stationary across a 16-CTA cluster.
//
// Replaces the same reference code as chain_mlp.cu (VPGDiffusion.forward / p_mean_var, diffusion_vpg.py:139-315, and
// DiffusionMLP.forward, mlp_diffusion.py:218-250) for the Hopper-sized case north_star singles out: 40 envs, K = 20.
// With so few rows every one of the S x (2 + 2 nb) dependent layers is pure latency: the tcgen05 kernel needs one
// weight pass per step and SM (>= 7.5 k cycles per 512 x 512 layer at the 34.7 B/cycle L2 -> shared-memory ingest of
// one SM), while the arithmetic is 40 x 512 x 512 MACs = 0.3 us of one SM's fp32 lanes.  So here the WEIGHTS never
// move: a cluster of 16 CTAs keeps one network resident in shared memory for the whole launch (CTA r owns output
// features [r H/16, (r+1) H/16) of every hidden layer, 138 KB at H = 512), each cluster serves up to 8 environments,
// and per layer only the activation vector moves: every CTA writes its 32 x EPC outputs into the next layer's input
// buffer of all 16 CTAs with st.shared::cluster (128-byte coalesced rows) and one hardware cluster barrier
// (barrier.cluster arrive.release / wait.acquire) closes the layer.  Arithmetic is exact fp32 FMA on the CUDA cores (no
// operand split needed); the residual stream stays in a register of the thread that owns (feature, env).
//
// Applicability (else the tcgen05 kernel runs): sampling mode, no cond_mlp, no LayerNorm, the per-CTA weight slice
// fits in shared memory (H = 512 with one block: hopper / walker2d), E <= 8 clusters x 8 environments.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace dppo {

namespace {

constexpr int kCS = 16;        // CTAs per cluster (non-portable size; one cluster per GPC)
constexpr int kThreadsS = 256;  // 512 threads (half the dot product per thread, 128 registers) measured slower: 0.165 vs 0.146 ms
constexpr int kMaxBlocks = 4;

struct SmallArgs {
  int D, Dc, td, H, nb, act, K0, K0p, in0, FS, OR, E, epc;
  const float* W0[2];
  const float* TB[2];                 // [K][H] layer-0 bias incl. time embedding (packed side table)
  const float* w1[2][kMaxBlocks];
  const float* b1[2][kMaxBlocks];
  const float* w2[2][kMaxBlocks];
  const float* b2[2][kMaxBlocks];
  const float* wo[2];
  const float* bo[2];
  const StepRow* rows;
  int S, ft, use_ddim, deterministic, use_base;
  float min_std, x0_clip, randn_clip, final_clip, eps_clip;
  const float* state;
  const float* noise;
  float* traj;
  float* chain;
  uint64_t seed, offset;
  int64_t env_offset;
  int* nonfinite;  // OR-ed with 1 when a final action element is NaN / Inf
  // completion ticket of the host-buffer call (dppo_sample_chain_host), nullptr otherwise: the last cluster to finish
  // stores `done_seq` into page-locked host memory behind a system-scope fence, so the host sees the results ~ a kernel
  // tear-down + stream-synchronise wake-up earlier than through the driver
  unsigned* done_counter;
  unsigned* done_flag;
  unsigned done_seq;
};

__device__ __forceinline__ float philox_normal_s(uint64_t seed, uint64_t offset, uint64_t elem, uint32_t slot) {
  uint32_t c0 = uint32_t(elem), c1 = uint32_t(elem >> 32), c2 = slot, c3 = uint32_t(offset);
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32) ^ uint32_t(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  const float u1 = (float(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = (float(c1 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__device__ __forceinline__ float act_s(int act, float x) { return act == DPPO_ACT_RELU ? fmaxf(x, 0.f) : mish_f(x); }

// floats of shared memory
__host__ __device__ inline size_t small_smem_floats(int FS, int K0p, int H, int nb, int OR, int EPC, int S) {
  const int KS = kThreadsS / (FS / 2);  // k-slices of a hidden layer (two features per thread)
  return size_t(FS) * K0p + size_t(2 * nb) * FS * (H + 4) + size_t(OR) * (H + 4)  // resident weight slices
         + size_t(2) * EPC * (H + 4 * KS)                                         // double-buffered layer input (padded slices)
         + size_t(EPC) * K0p                                                      // layer-0 input [x | obs]
         + size_t(KS) * EPC * FS                                                  // k-slice partial sums
         + size_t(2 * nb) * FS + size_t(OR)                                       // bias slices (b1 / b2 per block, output)
         + size_t(S) * (sizeof(StepRow) / 4) + 4                                  // the schedule rows
         + 12;                                                                    // four mbarriers (8-byte aligned)
}

template <int EPC>
__global__ void __launch_bounds__(kThreadsS, 1) chain_small_kernel(const SmallArgs a) {
  extern __shared__ __align__(16) float sm[];
  // Data hand-off between the 16 CTAs: every published value is an async remote store that completes its 4 bytes on the
  // RECEIVER's mbarrier (st.async ... mbarrier::complete_tx), and every CTA waits on its own barrier for the byte count of a
  // full buffer.  [0], [1]: the two layer-input buffers, [2]: the layer-0 input x.  (The first version closed every layer
  // with barrier.cluster.arrive.release / wait.acquire: MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR on the critical path of each
  // of the 80 dependent layers of a launch.)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x / kCS;
  const int env0 = cluster * a.epc;                      // first environment of this cluster
  const int ne = min(a.epc, a.E - env0);                 // environments this cluster serves (<= EPC)
  const int H = a.H, FS = a.FS, K0p = a.K0p, HS = H + 4; // HS: padded weight-row stride (bank spread for LDS.128)
  float* w0 = sm;                                        // [FS][K0p]
  float* wh = w0 + size_t(FS) * K0p;                     // [2 nb][FS][HS]
  float* wos = wh + size_t(2 * a.nb) * FS * HS;          // [OR][HS]
  float* xb = wos + size_t(a.OR) * HS;                   // [2][EPC][HP], column c at c + 4 (c / KQ)
  // Hidden layers: a thread owns TWO features (fa, fa + FS/2) x one k-slice, so every activation it loads feeds two
  // FMAs: the dot products are bound by shared-memory wavefronts (a broadcast read of x delivers 16 bytes per wavefront),
  // and with one feature per thread 96 of the 160 wavefronts per warp were those reads.  A warp then spans two or more
  // k-slices; 4 floats of padding per slice put their broadcast addresses into different banks.
  const int FP = FS / 2;                                 // feature pairs
  const int KS = kThreadsS / FP;                         // k-slices of a hidden layer
  const int KQ = H / KS;                                 // columns per slice (multiple of 4)
  const int HP = H + 4 * KS;                             // padded length of one environment's activation row
  float* x0 = xb + size_t(2) * EPC * HP;                 // [EPC][K0p]
  float* red = x0 + size_t(EPC) * K0p;                   // [KS][EPC][FS]
  float* bh = red + size_t(KS) * EPC * FS;               // [2 nb][FS] hidden biases of the owned features
  float* bos = bh + size_t(2 * a.nb) * FS;               // [OR] output biases of the owned rows
  // the whole schedule lives in shared memory: no global load at a step start
  StepRow* s_rows = reinterpret_cast<StepRow*>(bos + ((a.OR + 3) & ~3));
  const int fa = t % FP, kq = t / FP;                    // compute role: (feature pair fa / fa + FP, k-slice)
  const int f = t % FS;                                  // reduce role: owned feature
  const int F = int(rank) * FS + f;                      // global feature index
  const bool owner = t < FS * EPC;                       // reduce role: thread (f, e = t / FS) owns one output
  const int oe = t / FS;

  uint64_t* sbar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(s_rows + a.S) + 7) & ~uintptr_t(7));
  const uint32_t xbytes = uint32_t(EPC) * H * 4u;     // one layer-input buffer: FS x EPC values from each of the 16 CTAs
  const uint32_t x0bytes = uint32_t(ne) * a.D * 4u;   // the next sample: one value per (environment, action element)
  uint32_t bph = 0;                                   // phase bits of sbar[0..2]
  // Launch latency of the weights-stationary scheme: the first network's hidden-layer slices (2 nb x FS rows of H floats,
  // 128 KB at H = 512) are fetched by bulk async copies issued here, one row per copy, completing on sbar[3]; they run
  // under the prologue (observation read over PCIe when `state` is page-locked host memory, x_T draws, the cluster
  // barrier) and layer 0 waits for them only where it used to load them through registers.
  const int net_first = (a.rows[0].ft && !a.use_base) ? 1 : 0;
  const int wh_rows = 2 * a.nb * FS;
  // the observation values of this cluster are requested now and stored into the layer-0 input behind the weight fetch
  const int n_state = ne * a.Dc;
  const bool state_early = n_state <= kThreadsS;
  float state_pre = 0.f;
  if (state_early && t < n_state) state_pre = a.state[size_t(env0) * a.Dc + t];
  if (warp == 0) {
    if (t == 0) {
      for (int i = 0; i < 4; ++i) mbar_init(&sbar[i], 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(&sbar[0], xbytes);
      mbar_arrive_expect_tx(&sbar[1], xbytes);
      mbar_arrive_expect_tx(&sbar[2], x0bytes);
      mbar_arrive_expect_tx(&sbar[3], uint32_t(wh_rows) * uint32_t(H) * 4u);
    }
    __syncwarp();
    for (int i = lane; i < wh_rows; i += 32) {
      const int l = i / FS, r = i % FS;
      const float* src = ((l & 1) ? a.w2[net_first][l >> 1] : a.w1[net_first][l >> 1]) + size_t(int(rank) * FS + r) * H;
      bulk_g2s(wh + size_t(l) * FS * HS + size_t(r) * HS, src, uint32_t(H) * 4u, &sbar[3]);
    }
  }
  // wait for a full buffer, then (one thread) expect the bytes of its next use.  Nobody can complete that next phase
  // before every thread here has passed this wait: it needs this CTA's own stores of a later layer, behind a __syncthreads.
  auto wait_full = [&](int i, uint32_t bytes) {
    mbar_wait_spin(&sbar[i], (bph >> i) & 1u);  // polling: the CTA has nothing else to issue (suspending wait: +1.3 %)
    bph ^= 1u << i;
    if (t == 0) mbar_arrive_expect_tx(&sbar[i], bytes);
  };
  // ---- prologue: layer-0 input [x_T | obs] of this cluster's environments (every CTA builds its own full copy)
  for (int i = t; i < a.S; i += kThreadsS) s_rows[i] = a.rows[i];
  for (int i = t; i < EPC * K0p; i += kThreadsS) x0[i] = 0.f;
  __syncthreads();
  if (!state_early)
    for (int i = t; i < n_state; i += kThreadsS) {
      const int e = i / a.Dc, k = i % a.Dc;
      x0[e * K0p + a.D + k] = a.state[size_t(env0 + e) * a.Dc + k];
    }
  for (int i = t; i < ne * a.D; i += kThreadsS) {
    const int e = i / a.D, j = i % a.D, env = env0 + e;
    const float x = a.noise ? a.noise[size_t(env) * a.D + j]
                            : philox_normal_s(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + j, 0u);
    x0[e * K0p + j] = x;
    if (rank == 0 && a.chain && a.ft == a.S) a.chain[(size_t(env) * (a.ft + 1)) * a.D + j] = x;
  }
  cluster_sync_all();

  int cur_net = -1, buf = 0;
  float hreg = 0.f;  // residual stream h[oe][F] of the owning thread
  float tb_next = 0.f, z_first = 0.f;
  for (int step = 0; step < a.S; ++step) {
    const StepRow row = s_rows[step];
    const int net = (row.ft && !a.use_base) ? 1 : 0;
    if (net != cur_net) {
      // ---- (re)load this CTA's weight slices: rows [rank FS, (rank+1) FS) of every hidden Linear, rows rank + 16 i of
      // the output Linear; layer 0 keeps the columns [x | obs] (its time columns live in the TB table)
      __syncthreads();
      for (int i = t; i < FS * a.K0; i += kThreadsS) {
        const int r = i / a.K0, k = i % a.K0;
        w0[r * K0p + k] = a.W0[net][size_t(int(rank) * FS + r) * a.in0 + (k < a.D ? k : k + a.td)];
      }
      const int q4 = H / 4;
      if (step == 0) {
        // hidden-layer slices of the first network: the bulk copies issued at kernel start; the observation values
        // requested there land in the layer-0 input now (every thread observes the barrier: async-proxy writes)
        if (state_early && t < n_state) x0[(t / a.Dc) * K0p + a.D + t % a.Dc] = state_pre;
        mbar_wait(&sbar[3], 0);
      } else {
        for (int l = 0; l < 2 * a.nb; ++l) {
          const float* src = (l & 1) ? a.w2[net][l >> 1] : a.w1[net][l >> 1];
          float* dst = wh + size_t(l) * FS * HS;
          for (int i = t; i < FS * q4; i += kThreadsS) {
            const int r = i / q4, c = i % q4;
            *reinterpret_cast<float4*>(dst + r * HS + c * 4) =
                *reinterpret_cast<const float4*>(src + size_t(int(rank) * FS + r) * H + c * 4);
          }
        }
      }
      for (int i = t; i < a.OR * q4; i += kThreadsS) {
        const int r = i / q4, c = i % q4, j = int(rank) + kCS * r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < a.D) v = *reinterpret_cast<const float4*>(a.wo[net] + size_t(j) * H + c * 4);
        *reinterpret_cast<float4*>(wos + r * HS + c * 4) = v;
      }
      for (int i = t; i < 2 * a.nb * FS; i += kThreadsS) {
        const int l = i / FS, r = i % FS;
        bh[i] = ((l & 1) ? a.b2[net][l >> 1] : a.b1[net][l >> 1])[int(rank) * FS + r];
      }
      for (int i = t; i < a.OR; i += kThreadsS) bos[i] = int(rank) + kCS * i < a.D ? a.bo[net][int(rank) + kCS * i] : 0.f;
      cur_net = net;
      __syncthreads();
    }
    // layer-0 bias row of this step: fetched one step ahead (the L2 latency of the load sat in front of layer 0's hand-off)
    const float tb = step == 0 ? (owner ? a.TB[net][size_t(row.t) * H + F] : 0.f) : tb_next;
    if (step + 1 < a.S) {
      const StepRow nrow = s_rows[step + 1];
      const int nnet = (nrow.ft && !a.use_base) ? 1 : 0;
      tb_next = owner ? a.TB[nnet][size_t(nrow.t) * H + F] : 0.f;
    }

    // every owning thread publishes one value into the next input buffer of all 16 CTAs (rows of FS consecutive floats).
    // (Measured and dropped: gathering four neighbouring features by shuffles into one 16-byte async store per quad and
    // every fourth CTA - a quarter of the transaction updates - is slower, 0.1606 vs 0.1544 ms.)
    auto publish = [&](float v) {
      if (!owner) return;
      float* dstp = xb + size_t(buf) * EPC * HP + oe * HP + F + 4 * (F / KQ);
#pragma unroll
      for (uint32_t p = 0; p < uint32_t(kCS); ++p) st_async_f32(dstp, p, v, &sbar[buf]);
    };

    // ---- layer 0: h = W0 [x | obs] + TB[t]
    {
      float v0 = 0.f;
      if (owner) {
        const float* wr = w0 + f * K0p;
        const float* xr = x0 + oe * K0p;
        float acc = 0.f;
        for (int k = 0; k < a.K0; ++k) acc = fmaf(wr[k], xr[k], acc);
        hreg = acc + tb;
        v0 = act_s(a.act, hreg);
      }
      publish(v0);
    }
    wait_full(buf, xbytes);

    // ---- residual blocks: y = W1 act(h) + b1 ; h += W2 act(y) + b2
    for (int l = 0; l < 2 * a.nb; ++l) {
      const float* xin = xb + size_t(buf) * EPC * HP + kq * (KQ + 4);
      const float* wra = wh + size_t(l) * FS * HS + fa * HS + kq * KQ;
      const float* wrb = wra + size_t(FP) * HS;
      float acca[EPC], accb[EPC];
#pragma unroll
      for (int e = 0; e < EPC; ++e) acca[e] = 0.f, accb[e] = 0.f;
#pragma unroll 8
      for (int i = 0; i < KQ; i += 4) {
        const float4 wa = *reinterpret_cast<const float4*>(wra + i);
        const float4 wb = *reinterpret_cast<const float4*>(wrb + i);
#pragma unroll
        for (int e = 0; e < EPC; ++e) {
          const float4 xv = *reinterpret_cast<const float4*>(xin + e * HP + i);
          acca[e] = fmaf(wa.x, xv.x, acca[e]), accb[e] = fmaf(wb.x, xv.x, accb[e]);
          acca[e] = fmaf(wa.y, xv.y, acca[e]), accb[e] = fmaf(wb.y, xv.y, accb[e]);
          acca[e] = fmaf(wa.z, xv.z, acca[e]), accb[e] = fmaf(wb.z, xv.z, accb[e]);
          acca[e] = fmaf(wa.w, xv.w, acca[e]), accb[e] = fmaf(wb.w, xv.w, accb[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < EPC; ++e) red[(kq * EPC + e) * FS + fa] = acca[e], red[(kq * EPC + e) * FS + fa + FP] = accb[e];
      __syncthreads();
      buf ^= 1;
      {
        float pv = 0.f;
        if (owner) {
          float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;  // KS is a multiple of 4; four chains instead of one of KS dependent adds
#pragma unroll 2
          for (int s = 0; s < KS; s += 4) {
            v0 += red[(s * EPC + oe) * FS + f], v1 += red[((s + 1) * EPC + oe) * FS + f];
            v2 += red[((s + 2) * EPC + oe) * FS + f], v3 += red[((s + 3) * EPC + oe) * FS + f];
          }
          float v = (v0 + v1) + (v2 + v3);
          const int b = l >> 1;
          v += bh[l * FS + f];
          if (!(l & 1)) {
            pv = act_s(a.act, v);
          } else {
            hreg += v;
            pv = b + 1 < a.nb ? act_s(a.act, hreg) : hreg;  // no activation between the last block and the output layer
          }
        }
        publish(pv);
      }
      if (l == 2 * a.nb - 1 && warp < a.OR * ne) {
        // the draw of this warp's first (row, env) pair of the output layer depends on nothing the network computes:
        // Philox + Box-Muller while the last hidden layer's values are still in flight
        const int r = warp / ne, e = warp % ne, j = int(rank) + kCS * r;
        if (j < a.D) {
          const int env = env0 + e;
          z_first = a.noise ? a.noise[(size_t(step + 1) * a.E + env) * a.D + j]
                            : philox_normal_s(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + j, uint32_t(step + 1));
        }
      }
      wait_full(buf, xbytes);
    }

    // ---- output layer (rows rank + 16 i) + posterior mean / noise injection (diffusion_vpg.py:165-224, 279-311);
    // one warp per (row, env) pair, lanes split K; x_next goes into the layer-0 input of every CTA
    {
      const bool last = step == a.S - 1;
      float stdv, f2 = row.f2, f3 = row.f3;
      if (a.deterministic) {
        f2 = row.f2_det, f3 = row.f3_det;
        stdv = a.use_ddim ? 0.f : (row.t == 0 ? 0.f : fmaxf(row.std_train, 1e-3f));
      } else {
        stdv = fmaxf(row.std_train, a.min_std);
      }
      const float* xin = xb + size_t(buf) * EPC * HP;
      for (int pq = warp; pq < a.OR * ne; pq += kThreadsS / 32) {
        const int r = pq / ne, e = pq % ne, j = int(rank) + kCS * r;
        if (j >= a.D) continue;
        const float* wr = wos + r * HS;
        const float* xr = xin + e * HP;
        const int env = env0 + e;
        // the draw does not depend on the network output: issue it (global load or Philox) ahead of the dot product
        float z = pq == warp ? z_first
                             : (a.noise ? a.noise[(size_t(step + 1) * a.E + env) * a.D + j]
                                        : philox_normal_s(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + j, uint32_t(step + 1)));
        float acc = 0.f;
#pragma unroll 4
        for (int k = lane * 4; k < H; k += 128) {
          const float4 wv = *reinterpret_cast<const float4*>(wr + k);
          const float4 xv = *reinterpret_cast<const float4*>(xr + k + 4 * (k / KQ));
          acc = fmaf(wv.x, xv.x, acc), acc = fmaf(wv.y, xv.y, acc), acc = fmaf(wv.z, xv.z, acc), acc = fmaf(wv.w, xv.w, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        float eps = acc + bos[r];
        const float x = x0[e * K0p + j];
        float xz, mu;
        if (!a.use_ddim) {
          xz = row.f0 * x - row.f1 * eps;
          if (a.x0_clip >= 0.f) xz = fminf(fmaxf(xz, -a.x0_clip), a.x0_clip);
          mu = f2 * xz + f3 * x;
        } else {
          xz = (x - row.f1 * eps) / row.f0;
          if (a.x0_clip >= 0.f) {
            xz = fminf(fmaxf(xz, -a.x0_clip), a.x0_clip);
            eps = (x - row.f0 * xz) / row.f1;
          }
          if (a.eps_clip >= 0.f) eps = fminf(fmaxf(eps, -a.eps_clip), a.eps_clip);
          mu = f2 * xz + f3 * eps;
        }
        z = fminf(fmaxf(z, -a.randn_clip), a.randn_clip);
        float xn = mu + stdv * z;
        if (last && a.final_clip >= 0.f) xn = fminf(fmaxf(xn, -a.final_clip), a.final_clip);
        if (lane == 0) {
          if (a.chain && row.slot >= 0) a.chain[(size_t(env) * (a.ft + 1) + row.slot) * a.D + j] = xn;
          if (last) {
            a.traj[size_t(env) * a.D + j] = xn;
            if (!(fabsf(xn) <= 3.0e38f)) atomicOr(a.nonfinite, 1);
          }
        }
        if (lane < kCS) st_async_f32(x0 + e * K0p + j, uint32_t(lane), xn, &sbar[2]);
      }
    }
    wait_full(2, x0bytes);
  }
  cluster_sync_all();  // nobody leaves while a peer's stores into it may still be in flight
  if (a.done_flag && rank == 0 && threadIdx.x == 0) {
    // the release / acquire cluster barrier above orders every result store of this cluster before this point; the fence
    // makes them visible system-wide (cumulative) before the ticket, the last ticket holder publishes the sequence number
    __threadfence_system();
    const unsigned clusters = gridDim.x / kCS;
    if (atomicAdd(a.done_counter, 1u) == clusters - 1) {
      *a.done_counter = 0;  // the next launch is stream-ordered behind this one
      __threadfence_system();
      *reinterpret_cast<volatile unsigned*>(a.done_flag) = a.done_seq;
    }
  }
}

template <int EPC>
int launch_small(const SmallArgs& a, int clusters, cudaStream_t st) {
  auto kfn = chain_small_kernel<EPC>;
  const size_t smem = small_smem_floats(a.FS, a.K0p, a.H, a.nb, a.OR, EPC, a.S) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(chain_small_kernel)");
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(clusters * kCS)), cfg.blockDim = dim3(kThreadsS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCS, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, a);
  if (e != cudaSuccess) return cuda_fail(e, "chain_small_kernel launch");
  return DPPO_OK;
}

}  // namespace

// How many environments one launch of the weights-stationary kernel can take (0 = geometry outside the kernel).
int small_chain_capacity(const dppo_ctx* ctx) {
  const MlpGeom& g = ctx->g;
  if (ctx->kind != 0 || g.CH || g.ln || g.nb > kMaxBlocks || g.H % (kCS * 4) || kThreadsS % (g.H / kCS)) return 0;
  {
    const int FP = g.H / kCS / 2;  // feature pairs per CTA: a thread owns a pair x one k-slice
    if (FP <= 0 || kThreadsS % FP || (g.H / (kThreadsS / FP)) % 4) return 0;
  }
  if (ctx->small_clusters <= 0) return 0;
  const int FS = g.H / kCS, K0p = ((g.D + g.Dc + 3) & ~3) + 1, OR = (g.D + kCS - 1) / kCS;
  int epc = kThreadsS / FS < 8 ? kThreadsS / FS : 8;
  while (epc > 0 && small_smem_floats(FS, K0p, g.H, g.nb, OR, epc, ctx->S) * sizeof(float) > 232448) --epc;
  return epc * ctx->small_clusters;
}

int sample_chain_small_impl(dppo_ctx* ctx, const float* state, int E, const float* noise, uint64_t seed, uint64_t offset,
                            int64_t env_offset, int deterministic, int use_base, float min_std, float* traj, float* chain,
                            cudaStream_t st) {
  const MlpGeom& g = ctx->g;
  SmallArgs a{};
  a.D = g.D, a.Dc = g.Dc, a.td = g.td, a.H = g.H, a.nb = g.nb, a.act = g.act;
  a.K0 = g.D + g.Dc, a.K0p = ((a.K0 + 3) & ~3) + 1, a.in0 = g.D + g.td + g.Dc;  // odd row stride: conflict-free scalar reads
  a.FS = g.H / kCS, a.OR = (g.D + kCS - 1) / kCS, a.E = E;
  for (int w = 0; w < 2; ++w) {
    const std::vector<const float*>& p = ctx->nets[w].raw;
    if (p.empty()) {
      if (w == 1 && use_base) continue;
      return set_error("small chain kernel: weights of network %d not packed", w), DPPO_ERR_STATE;
    }
    int i = 4;  // time_embedding.{1,3}.{weight,bias} are folded into the TB table
    a.W0[w] = p[i], i += 2;
    for (int b = 0; b < g.nb; ++b) a.w1[w][b] = p[i], a.b1[w][b] = p[i + 1], a.w2[w][b] = p[i + 2], a.b2[w][b] = p[i + 3], i += 4;
    a.wo[w] = p[i], a.bo[w] = p[i + 1];
    a.TB[w] = ctx->nets[w].side + g.off_tb;
  }
  a.rows = ctx->d_rows, a.S = ctx->S, a.ft = ctx->ft, a.use_ddim = ctx->use_ddim;
  a.deterministic = deterministic, a.use_base = use_base;
  a.min_std = min_std, a.x0_clip = ctx->x0_clip, a.randn_clip = ctx->randn_clip, a.final_clip = ctx->final_clip;
  a.eps_clip = ctx->eps_clip;
  a.state = state, a.noise = noise, a.traj = traj, a.chain = chain;
  a.seed = seed, a.offset = offset, a.env_offset = env_offset;
  a.nonfinite = ctx->d_nonfinite;
  if (ctx->done_want && ctx->h_done) {
    a.done_counter = reinterpret_cast<unsigned*>(ctx->d_nonfinite) + 1;
    a.done_flag = ctx->h_done, a.done_seq = ++ctx->done_seq;
    ctx->done_armed = 1;
  }
  // spread the environments over as many clusters as the device co-schedules
  const int clusters = E < ctx->small_clusters ? E : ctx->small_clusters;
  const int epc = (E + clusters - 1) / clusters;
  a.epc = epc;
  const int used = (E + epc - 1) / epc;
  switch (epc) {
    case 1: return launch_small<1>(a, used, st);
    case 2: return launch_small<2>(a, used, st);
    case 3: return launch_small<3>(a, used, st);
    case 4: return launch_small<4>(a, used, st);
    case 5: return launch_small<5>(a, used, st);
    case 6: return launch_small<6>(a, used, st);
    case 7: return launch_small<7>(a, used, st);
    case 8: return launch_small<8>(a, used, st);
  }
  return set_error("small chain kernel: %d environments per cluster", epc), DPPO_ERR_INVALID;
}

// number of 16-CTA clusters of this kernel the device can co-schedule (queried once per context)
int small_chain_query_clusters(int H, int D, int Dc, int nb, int S) {
  auto kfn = chain_small_kernel<8>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess ||
      cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCS * 8), cfg.blockDim = dim3(kThreadsS);
  const int FS = H / kCS, K0p = ((D + Dc + 3) & ~3) + 1, OR = (D + kCS - 1) / kCS;
  size_t smem = small_smem_floats(FS, K0p, H, nb, OR, 8, S) * sizeof(float);
  if (smem > 232448) smem = 232448;
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCS, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

}  // namespace dppo
