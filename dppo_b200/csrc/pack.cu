// Weight repacking for the denoise-chain kernel.
//   pack_linear_tiles : fp32 nn.Linear weight [out, in] -> bf16 (hi [+ lo]) 128x64 tiles in the SWIZZLE_128B K-major
//                       image tcgen05.mma reads, laid out in the order the kernel streams them (m-tile major, k-chunk
//                       minor, hi tile then lo tile).
//   time_bias_table   : folds SinusoidalPosEmb -> Linear -> Mish -> Linear (reference dppo/model/diffusion/modules.py:14-27,
//                       mlp_diffusion.py:191-196) and the time columns of the trunk's first Linear into a per-timestep bias
//                       table TB[t][f] = b0[f] + W0[f, D:D+td] . temb(t)  (t is uniform over the batch while sampling).
#include "common.cuh"
#include "internal.h"

namespace dppo {

__global__ void pack_linear_tiles_kernel(const float* __restrict__ W, int out_f, int in_f, int skip_at, int skip_n,
                                         int MTl, int KCl, int nsplit, uint8_t* __restrict__ dst) {
  // one thread per (row, 8-column group)
  const int groups_per_row = KCl * 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)MTl * 128 * groups_per_row;
  if (idx >= total) return;
  const int r = int(idx / groups_per_row);
  const int cgx = int(idx % groups_per_row);
  const int kc = cgx >> 3, cg = cgx & 7;
  const int mt = r >> 7, ri = r & 127;
  const int in_eff = in_f - skip_n;
  __align__(16) __nv_bfloat16 hi[8];
  __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = kc * 64 + cg * 8 + j;
    float v = 0.f;
    if (r < out_f && k < in_eff) {
      const int src = k < skip_at ? k : k + skip_n;
      v = W[(size_t)r * in_f + src];
    }
    split_bf16(v, hi[j], lo[j]);
  }
  const size_t group = (size_t)mt * KCl + kc;
  uint8_t* tile_hi = dst + group * nsplit * 16384;
  const uint32_t off = ri * 128u + (uint32_t((cg ^ (ri & 7))) << 4);
  *reinterpret_cast<uint4*>(tile_hi + off) = *reinterpret_cast<const uint4*>(hi);
  if (nsplit == 2) *reinterpret_cast<uint4*>(tile_hi + 16384 + off) = *reinterpret_cast<const uint4*>(lo);
}

__device__ __forceinline__ float mish_exact(float x) {
  // torch: x * tanh(softplus(x)), softplus(x) = x for x > 20 else log1p(exp(x))
  const float sp = x > 20.f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}

// grid = K timesteps, block = 128 threads.  tw1 [2td, td], tb1 [2td], tw2 [td, 2td], tb2 [td], W0 [H, in0], b0 [H]
__global__ void time_bias_table_kernel(const float* __restrict__ tw1, const float* __restrict__ tb1,
                                       const float* __restrict__ tw2, const float* __restrict__ tb2,
                                       const float* __restrict__ W0, const float* __restrict__ b0, int td, int D,
                                       int in0, int H, float* __restrict__ TB) {
  __shared__ float s_emb[64];
  __shared__ float s_hid[128];
  __shared__ float s_out[64];
  const int t = blockIdx.x;
  const int half = td / 2;
  if (threadIdx.x < half) {
    const float rate = float(log(10000.0) / double(half - 1));
    const float freq = expf(float(threadIdx.x) * -rate);
    const float ph = float(t) * freq;
    s_emb[threadIdx.x] = sinf(ph);
    s_emb[threadIdx.x + half] = cosf(ph);
  }
  __syncthreads();
  if (threadIdx.x < 2 * td) {
    float acc = 0.f;
    for (int j = 0; j < td; ++j) acc += s_emb[j] * tw1[threadIdx.x * td + j];
    s_hid[threadIdx.x] = mish_exact(acc + tb1[threadIdx.x]);
  }
  __syncthreads();
  if (threadIdx.x < td) {
    float acc = 0.f;
    for (int j = 0; j < 2 * td; ++j) acc += s_hid[j] * tw2[threadIdx.x * 2 * td + j];
    s_out[threadIdx.x] = acc + tb2[threadIdx.x];
  }
  __syncthreads();
  for (int f = threadIdx.x; f < H; f += blockDim.x) {
    float acc = 0.f;
    const float* w = W0 + (size_t)f * in0 + D;
    for (int j = 0; j < td; ++j) acc += w[j] * s_out[j];
    TB[(size_t)t * H + f] = acc + b0[f];
  }
}

__global__ void copy_pad_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = i < n ? src[i] : 0.f;
}

static int launch_pack_linear(const float* W, int out_f, int in_f, int skip_at, int skip_n, int MTl, int KCl,
                              int nsplit, uint8_t* dst, cudaStream_t st) {
  const long long total = (long long)MTl * 128 * KCl * 8;
  const int threads = 256;
  pack_linear_tiles_kernel<<<unsigned((total + threads - 1) / threads), threads, 0, st>>>(W, out_f, in_f, skip_at,
                                                                                           skip_n, MTl, KCl, nsplit, dst);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// dst = src + prev (prev may be null): running sum of the l2 biases over the residual blocks
__global__ void add_vec_kernel(const float* __restrict__ src, const float* __restrict__ prev, int n, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] + (prev ? prev[i] : 0.f);
}

static void copy_pad(const float* src, int n, float* dst, int n_pad, cudaStream_t st) {
  copy_pad_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(src, n, dst, n_pad);
}

int pack_mlp_impl(dppo_ctx* ctx, int which, const float* const* p, int n_params, cudaStream_t st) {
  const MlpGeom& g = ctx->g;
  const int expect = 4 + (g.CH ? 4 : 0) + 2 + g.nb * (g.ln ? 8 : 4) + 2;
  if (n_params != expect) {
    set_error("dppo_pack_mlp: expected %d parameter tensors for this geometry, got %d", expect, n_params);
    return DPPO_ERR_INVALID;
  }
  PackedNet& net = ctx->nets[which];
  int i = 0;
  const float *tw1 = p[i++], *tb1 = p[i++], *tw2 = p[i++], *tb2 = p[i++];
  const float *cw0 = nullptr, *cb0 = nullptr, *cw1 = nullptr, *cb1 = nullptr;
  if (g.CH) {
    cw0 = p[i++], cb0 = p[i++], cw1 = p[i++], cb1 = p[i++];
  }
  const float *W0 = p[i++], *b0 = p[i++];
  const int in0 = g.D + g.td + g.Dc;
  const size_t tile_group = (size_t)g.nsplit * 16384;

  // layer-0 bias table over all K timesteps
  time_bias_table_kernel<<<ctx->K, 128, 0, st>>>(tw1, tb1, tw2, tb2, W0, b0, g.td, g.D, in0, g.H, net.side + g.off_tb);

  uint8_t* cur = net.tiles + g.off_cond_tiles;
  if (g.CH) {
    if (launch_pack_linear(cw0, g.CH, g.Dc_in, 0, 0, g.MTc, g.KCc, g.nsplit, cur, st)) goto fail;
    cur += (size_t)g.MTc * g.KCc * tile_group;
    if (launch_pack_linear(cw1, g.CO, g.CH, 0, 0, 1, g.CH / 64, g.nsplit, cur, st)) goto fail;
    cur += (size_t)(g.CH / 64) * tile_group;
    copy_pad(cb0, g.CH, net.side + g.off_bc0, g.CH, st);
    copy_pad(cb1, g.CO, net.side + g.off_bc1, 128, st);
  }
  cur = net.tiles + g.off_step_tiles;
  // layer 0: columns [x (D) | time (td, folded into TB) | cond (Dc)] -> [x | cond]
  if (launch_pack_linear(W0, g.H, in0, g.D, g.td, g.MT, g.KC0, g.nsplit, cur, st)) goto fail;
  cur += (size_t)g.MT * g.KC0 * tile_group;
  for (int b = 0; b < g.nb; ++b) {
    const float *w1 = p[i++], *b1 = p[i++], *w2 = p[i++], *b2 = p[i++];
    float* side = net.side + g.off_blk + (size_t)b * g.blk_stride;
    if (launch_pack_linear(w1, g.H, g.H, 0, 0, g.MT, g.KCH, g.nsplit, cur, st)) goto fail;
    cur += (size_t)g.MT * g.KCH * tile_group;
    if (launch_pack_linear(w2, g.H, g.H, 0, 0, g.MT, g.KCH, g.nsplit, cur, st)) goto fail;
    cur += (size_t)g.MT * g.KCH * tile_group;
    copy_pad(b1, g.H, side, g.H, st);
    // the residual stream's biases are added by the epilogue, not stored back into TMEM: slot 1 = b2_0 + ... + b2_b
    add_vec_kernel<<<(g.H + 255) / 256, 256, 0, st>>>(b2, b > 0 ? side - g.blk_stride + g.H : nullptr, g.H, side + g.H);
    if (g.ln) {
      for (int q = 0; q < 4; ++q) copy_pad(p[i++], g.H, side + (size_t)(2 + q) * g.H, g.H, st);  // g1 be1 g2 be2
    }
  }
  {
    const float *wo = p[i++], *bo = p[i++];
    if (launch_pack_linear(wo, g.D, g.H, 0, 0, 1, g.KCH, g.nsplit, cur, st)) goto fail;
    copy_pad(bo, g.D, net.side + g.off_bout, 128, st);
  }
  if (cudaGetLastError() != cudaSuccess) goto fail;
  net.raw.assign(p, p + n_params);
  net.packed = true;
  return DPPO_OK;
fail:
  return cuda_fail(cudaGetLastError(), "dppo_pack_mlp launch");
}

}  // namespace dppo
