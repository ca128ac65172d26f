// Shared by the DiffusionMLP chain kernels (chain_mlp.cu: feature-split clusters, cta_group::1; chain_pair.cu: CTA pairs
// with cta_group::2 MMAs): the kernel argument block, the Philox / Box-Muller draw and the activation.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace dppo {

struct ChainArgs {
  // geometry
  int D, Dc_in, Dc, H, nb, act, ln, CH, CO, MT, KCH, KC0, KCc, MTc, nsplit, nstage;
  int C;  // CTAs per cluster (feature split), divides MT
  int early;  // block layers run in the early-first-tile order (two M tiles per CTA, no LayerNorm): see run_layer
  uint32_t off_tb, off_blk, blk_stride, off_bout, off_bc0, off_bc1;
  const uint8_t* tiles[2];
  const float* side[2];
  uint32_t n_cond_tiles, n_step_tiles;
  size_t off_step_tiles;
  // schedule
  const StepRow* rows;
  int S, ft, first_step, eval_mode, use_ddim;
  int deterministic, use_base;
  float min_std, x0_clip, randn_clip, final_clip, eps_clip;
  // io
  const float* state;
  int E;
  const float* noise;
  float* traj;
  float* chain;
  const float* chains_in;
  float* logp;
  uint64_t seed, offset;
  int64_t env_offset;
  unsigned long long* prof;  // optional [grid][16] cycle counters (bring-up / profiling), nullptr in production
  int* nonfinite;            // OR-ed with 1 when a final action element is NaN / Inf (NaN observations, diverged weights)
};

// ---------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t offset, uint64_t elem, uint32_t slot) {
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

template <int ACT>
__device__ __forceinline__ float activate(float x) {
  if (ACT == DPPO_ACT_RELU) return fmaxf(x, 0.f);
  return mish_f(x);
}

// store one activation value into a K-major SWIZZLE_128B operand (rows = NE environments)
template <int NE>
__device__ __forceinline__ void store_operand(uint8_t* hi, uint8_t* lo, int row, int k, float v, bool split) {
  const uint32_t off = sw128_offset(uint32_t(row), uint32_t(k), NE);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
  if (split) *reinterpret_cast<__nv_bfloat16*>(lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
}

// chain_pair.cu: applicability test and launch of the CTA-pair kernel (NE = 64 environments per pair, two CTAs, no
// LayerNorm, no cond_mlp); `a` comes from sample_chain_impl with C = 2.  Opt-in: DPPO_B200_PAIR=1 or a forced shape
// (measured 2 % behind the cta_group::1 kernel on the headline shape).
bool chain_pair_applicable(const MlpGeom& g, int NE, int C, bool forced);
int launch_chain_pair(ChainArgs a, const MlpGeom& g, cudaStream_t st);

}  // namespace dppo
