// Host plan of the lowered Unet1D (see unet_plan.h), weight / side-table packing kernels, and the host-only
// inspection entry points the CPU tests use to check the lowering against the oracle without a GPU.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "internal.h"
#include "unet_plan.h"

namespace dppo {

static int cdiv(int a, int b) { return (a + b - 1) / b; }
static int round_up(int a, int b) { return cdiv(a, b) * b; }
static int ilog2_exact(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return (1 << s) == v ? s : -1;
}

namespace {

struct Seg {
  int chunk, C;
};

struct Builder {
  UnetPlan& P;
  const dppo_unet_desc& d;
  int pc = 0;          // parameter cursor
  uint32_t tile = 0;   // tile cursor (16 KiB units)
  size_t side = 0;     // side-table cursor (floats)
  int err = 0;

  Builder(UnetPlan& p, const dppo_unet_desc& dd) : P(p), d(dd) {}

  // side[dst + f] = param[channel of f] for f < n * rep, zero padded to pad_to.  tmajor: channel = f % n, else f / rep.
  int side_copy(int param, int n, int rep, int tmajor, int pad_to) {
    USideJob j{};
    j.kind = U_SIDE_COPY, j.param = param, j.n = n, j.dst = int(side), j.pad_to = pad_to, j.lin_in = rep, j.act_first = tmajor;
    P.side_jobs.push_back(j);
    side += pad_to;
    return j.dst;
  }
  int side_time(int wparam, int n, int pad_to, int lin_in, int act_first) {
    USideJob j{};
    j.kind = U_SIDE_TIME, j.param = wparam, j.n = n, j.dst = int(side), j.pad_to = pad_to, j.lin_in = lin_in, j.act_first = act_first;
    P.side_jobs.push_back(j);
    side += size_t(P.K) * pad_to;
    return j.dst;
  }

  UGemm finish(UPackJob& j, int acc_tile, const int* chunks) {
    j.MT = cdiv(j.out_f, 128);
    j.KC = j.seg_chunks[0] + (j.nseg > 1 ? j.seg_chunks[1] : 0);
    j.tile_off = tile;
    tile += uint32_t(j.MT) * j.KC * P.nsplit;
    P.jobs.push_back(j);
    UGemm g{};
    g.tile_off = j.tile_off, g.mt = uint16_t(j.MT), g.kc = uint16_t(j.KC);
    g.src_chunk[0] = uint16_t(chunks[0]), g.src_n[0] = uint16_t(j.seg_chunks[0]);
    g.src_chunk[1] = uint16_t(j.nseg > 1 ? chunks[1] : 0), g.src_n[1] = uint16_t(j.nseg > 1 ? j.seg_chunks[1] : 0);
    g.acc_tile = uint16_t(acc_tile);
    return g;
  }

  UGemm lin_gemm(int wparam, int out_f, int in_valid, int lin_in, int col0, int src_chunk, int acc_tile) {
    UPackJob j{};
    j.param = wparam, j.kind = U_PACK_LINEAR, j.out_f = out_f, j.nseg = 1;
    j.seg_f[0] = in_valid, j.seg_chunks[0] = cdiv(in_valid, 64);
    j.lin_in = lin_in, j.lin_col0 = col0;
    P.macs_dense += double(out_f) * in_valid;
    return finish(j, acc_tile, &src_chunk);
  }

  UGemm conv_gemm(int wparam, int kind, const Seg* segs, int nseg, int Tin, int Cout, int Tout, int ks, int stride,
                  int pad, int out_tmajor, int acc_tile) {
    UPackJob j{};
    j.param = wparam, j.kind = kind, j.out_f = Cout * Tout, j.Cout = Cout, j.Tout = Tout, j.out_tmajor = out_tmajor;
    j.nseg = nseg, j.Tin = Tin, j.ks = ks, j.stride = stride, j.pad = pad;
    int chunks[2] = {0, 0}, c0 = 0;
    for (int s = 0; s < nseg; ++s) {
      j.seg_f[s] = segs[s].C * Tin, j.seg_c0[s] = c0, j.seg_chunks[s] = cdiv(segs[s].C * Tin, 64);
      chunks[s] = segs[s].chunk;
      c0 += segs[s].C;
    }
    j.Cin_total = c0;
    int taps = 0;
    for (int to = 0; to < Tout; ++to)
      for (int ti = 0; ti < Tin; ++ti) {
        const int q = kind == U_PACK_CONV ? ti - to * stride + pad : to - ti * stride + pad;
        taps += q >= 0 && q < ks;
      }
    P.macs_dense += double(taps) * Cout * c0;
    return finish(j, acc_tile, chunks);
  }

  ULayer blank() {
    ULayer L;
    memset(&L, 0, sizeof(L));
    L.res_chunk = L.dst_chunk = -1;
    return L;
  }

  int gn_group(int C, int T) {
    // a group is a run of gs consecutive lowered features.  Powers of two never straddle a warp's 32 TMEM lanes (shuffle
    // butterflies); any other size <= 32 (robomimic can / lift: 5 channels x 4 positions = 20) takes the kernel's
    // segmented path, whose groups may span two warps or two M tiles and exchange partial sums through shared memory
    const int gs = C / d.n_groups * T;
    if (C % d.n_groups || gs < 1 || gs > 32) {
      set_error("unet: GroupNorm group of %d channels x %d positions = %d features; the kernel needs groups of <= 32 features",
                d.n_groups ? C / d.n_groups : 0, T, gs);
      err = DPPO_ERR_UNSUPPORTED;
      return 1;
    }
    if (ilog2_exact(gs) < 0) P.gn_segmented = 1;
    return gs;
  }

  // Conv1dBlock epilogue fields (bias, GroupNorm, activation); parameters conv w,b at pw, norm w,b at pw+2
  void conv_block_epi(ULayer& L, int pw, int C, int T) {
    const int nf = C * T, nfp = round_up(nf, 128);
    L.acc_tile = 0, L.mt = nfp / 128, L.nf = nf;
    L.bias_off = side_copy(pw + 1, C, T, 0, nfp);
    L.gn_size = gn_group(C, T);
    L.gamma_off = side_copy(pw + 2, C, T, 0, nfp);
    L.beta_off = side_copy(pw + 3, C, T, 0, nfp);
    L.gn_eps = d.groupnorm_eps;
    L.act = 1;
  }

  // ResidualBlock1D (unet.py:27-118): input = channel concatenation of `segs` at length T; output -> dst_chunk
  void res_block(const Seg* segs, int nseg, int Cout, int T, int dst_chunk, int y_chunk) {
    int Cin = 0;
    for (int s = 0; s < nseg; ++s) Cin += segs[s].C;
    const int p0 = pc;
    const int n_cond = d.larger_encoder ? 3 : 1;
    const bool res_conv = Cin != Cout;
    pc += 8 + 2 * n_cond + (res_conv ? 2 : 0);
    const int pcond = p0 + 8, pres = p0 + 8 + 2 * n_cond;
    const int fd = d.cond_predict_scale ? 2 * Cout : Cout, fdp = round_up(fd, 128);
    const int gin = d.time_dim + d.cond_dim;
    // ---- FiLM encoder
    if (d.larger_encoder) {
      ULayer L1 = blank();
      L1.n_gemm = 1, L1.g[0] = lin_gemm(pcond, fd, d.cond_dim, gin, d.time_dim, P.chunk_state, 0);
      L1.kind = U_EPI_OPERAND, L1.acc_tile = 0, L1.mt = fdp / 128, L1.nf = fd;
      L1.bias_off = side_time(pcond, fd, fdp, gin, 0), L1.bias_tstride = fdp;
      L1.act = 1, L1.dst_chunk = P.chunk_condh, L1.track = 1;
      P.layers.push_back(L1);
      ULayer L2 = blank();
      L2.n_gemm = 1, L2.g[0] = lin_gemm(pcond + 2, fd, fd, fd, 0, P.chunk_condh, 0);
      L2.kind = U_EPI_OPERAND, L2.acc_tile = 0, L2.mt = fdp / 128, L2.nf = fd;
      L2.bias_off = side_copy(pcond + 3, fd, 1, 0, fdp), L2.act = 1, L2.dst_chunk = P.chunk_condh, L2.track = 1;
      P.layers.push_back(L2);
      ULayer L3 = blank();
      L3.n_gemm = 1, L3.g[0] = lin_gemm(pcond + 4, fd, fd, fd, 0, P.chunk_condh, 0);
      L3.kind = U_EPI_FILM, L3.acc_tile = 0, L3.mt = fdp / 128, L3.nf = fd;
      L3.bias_off = side_copy(pcond + 5, fd, 1, 0, fdp), L3.track = 1;
      P.layers.push_back(L3);
    } else {
      ULayer L = blank();
      L.n_gemm = 1, L.g[0] = lin_gemm(pcond, fd, d.cond_dim, gin, d.time_dim, P.chunk_state_act, 0);
      L.kind = U_EPI_FILM, L.acc_tile = 0, L.mt = fdp / 128, L.nf = fd;
      L.bias_off = side_time(pcond, fd, fdp, gin, 1), L.bias_tstride = fdp, L.track = 1;
      P.layers.push_back(L);
    }
    // ---- blocks[0]: conv -> GroupNorm -> act, then FiLM
    {
      ULayer L = blank();
      L.n_gemm = 1;
      L.g[0] = conv_gemm(p0, U_PACK_CONV, segs, nseg, T, Cout, T, d.kernel_size, 1, d.kernel_size / 2, 0, 0);
      L.kind = U_EPI_OPERAND;
      conv_block_epi(L, p0, Cout, T);
      L.film = d.cond_predict_scale ? 2 : 1, L.film_c = Cout, L.film_tshift = ilog2_exact(T);
      L.dst_chunk = y_chunk;
      P.layers.push_back(L);
    }
    // ---- blocks[1] + residual
    {
      ULayer L = blank();
      const Seg ys{y_chunk, Cout};
      L.n_gemm = 1;
      L.g[0] = conv_gemm(p0 + 4, U_PACK_CONV, &ys, 1, T, Cout, T, d.kernel_size, 1, d.kernel_size / 2, 0, 0);
      L.kind = U_EPI_OPERAND;
      conv_block_epi(L, p0 + 4, Cout, T);
      if (res_conv) {
        L.n_gemm = 2;
        L.g[1] = conv_gemm(pres, U_PACK_CONV, segs, nseg, T, Cout, T, 1, 1, 0, 0, P.MTmax);
        L.res = U_RES_ACC, L.res_acc_tile = P.MTmax;
        L.res_bias_off = side_copy(pres + 1, Cout, T, 0, round_up(Cout * T, 128));
      } else {
        if (nseg != 1) {
          set_error("unet: identity residual over a concatenated input is not supported");
          err = DPPO_ERR_UNSUPPORTED;
        }
        L.res = U_RES_SLOT, L.res_chunk = segs[0].chunk;
      }
      L.dst_chunk = dst_chunk;
      P.layers.push_back(L);
    }
  }
};

}  // namespace

int unet_build_plan(const dppo_unet_desc& d, int K, int precision, UnetPlan* out) {
  UnetPlan& P = *out;
  P = UnetPlan{};
  P.d = d, P.K = K, P.nsplit = precision == DPPO_PRECISION_SPLIT3 ? 2 : 1;
  P.D = d.action_dim * d.horizon_steps, P.e = d.time_dim;
  const int nl = d.n_levels, Ta = d.horizon_steps;
  if (nl < 1 || nl > DPPO_UNET_MAX_LEVELS) return set_error("unet: n_levels %d outside [1,%d]", nl, DPPO_UNET_MAX_LEVELS), DPPO_ERR_INVALID;
  if (P.D < 1 || P.D > 128) return set_error("unet: Ta*Da = %d outside [1,128]", P.D), DPPO_ERR_INVALID;
  if (ilog2_exact(Ta) < 0 || (Ta >> (nl - 1)) < 1)
    return set_error("unet: horizon_steps %d must be a power of two >= 2^(levels-1)", Ta), DPPO_ERR_UNSUPPORTED;
  if (d.kernel_size < 1 || d.kernel_size % 2 == 0 || d.kernel_size > 9) return set_error("unet: kernel_size %d", d.kernel_size), DPPO_ERR_INVALID;
  if (d.time_dim % 2 || d.time_dim < 4 || d.time_dim > 32) return set_error("unet: time_dim %d unsupported", d.time_dim), DPPO_ERR_INVALID;
  if (d.n_groups < 1) return set_error("unet: n_groups must be set (GroupNorm)"), DPPO_ERR_UNSUPPORTED;
  if (d.cond_dim < 1 || d.cond_dim > 512) return set_error("unet: cond_dim %d", d.cond_dim), DPPO_ERR_INVALID;
  if (d.activation != DPPO_ACT_RELU && d.activation != DPPO_ACT_MISH) return set_error("unet: activation %d", d.activation), DPPO_ERR_INVALID;
  int widths[DPPO_UNET_MAX_LEVELS + 1];
  widths[0] = d.action_dim;
  for (int l = 0; l < nl; ++l) {
    widths[l + 1] = d.dim * d.dim_mults[l];
    if (widths[l + 1] < 1) return set_error("unet: level %d has no channels", l), DPPO_ERR_INVALID;
  }
  // widest activation / FiLM vector
  int wmax = 0, fdmax = 0;
  for (int l = 0; l < nl; ++l) {
    const int T = Ta >> l;
    wmax = wmax > widths[l + 1] * T ? wmax : widths[l + 1] * T;
    const int fd = d.cond_predict_scale ? 2 * widths[l + 1] : widths[l + 1];
    fdmax = fdmax > fd ? fdmax : fd;
    if (l >= 1) wmax = wmax > widths[l] * T * 2 ? wmax : widths[l] * T * 2;  // up path: dims[l] channels at T_l and 2 T_l
  }
  P.KA = round_up(cdiv(wmax, 64), 2);
  P.MTmax = P.KA / 2;
  P.film_dim = fdmax;
  P.KX = cdiv(P.D, 64), P.KS = cdiv(d.cond_dim, 64);
  // every slot a layer epilogue writes starts on an even chunk: an M tile of 128 features is a chunk PAIR, the unit of
  // the tile hand-off (the sample / observation slots are written element-wise and only need to come first)
  int c = 0;
  P.chunk_x = c, c += P.KX;
  P.chunk_state = c, c += P.KS;
  if (!d.larger_encoder) P.chunk_state_act = c, c += P.KS;
  c = round_up(c, 2);
  P.n_condh = round_up(cdiv(fdmax, 64), 2);
  P.chunk_condh = c, c += P.n_condh;
  const int chunk_h = c;
  c += P.KA;
  const int chunk_y = c;
  c += P.KA;
  int chunk_skip[DPPO_UNET_MAX_LEVELS] = {0, 0, 0, 0};
  for (int l = 1; l < nl; ++l) chunk_skip[l] = c, c += P.KA;
  P.total_chunks = c;

  Builder B(P, d);
  B.pc = 4;  // time_mlp.{1,3}.{weight,bias}
  int T = Ta;
  Seg cur{P.chunk_x, widths[0]};
  // ---- down path (unet.py:286-300)
  for (int l = 0; l < nl; ++l) {
    const int C = widths[l + 1];
    B.res_block(&cur, 1, C, T, chunk_h, chunk_y);
    const Seg h{chunk_h, C};
    const int out2 = l >= 1 ? chunk_skip[l] : chunk_h;
    B.res_block(&h, 1, C, T, out2, chunk_y);
    cur = Seg{out2, C};
    if (l < nl - 1) {  // Downsample1d: Conv1d(C, C, 3, 2, 1)
      ULayer L = B.blank();
      L.n_gemm = 1, L.g[0] = B.conv_gemm(B.pc, U_PACK_CONV, &cur, 1, T, C, T / 2, 3, 2, 1, 0, 0);
      const int nf = C * (T / 2), nfp = round_up(nf, 128);
      L.kind = U_EPI_OPERAND, L.acc_tile = 0, L.mt = nfp / 128, L.nf = nf;
      L.bias_off = B.side_copy(B.pc + 1, C, T / 2, 0, nfp);
      L.dst_chunk = chunk_h;
      P.layers.push_back(L);
      B.pc += 2;
      T /= 2;
      cur = Seg{chunk_h, C};
    }
  }
  // ---- middle (unet.py:302-303)
  {
    const int C = widths[nl];
    B.res_block(&cur, 1, C, T, chunk_h, chunk_y);
    const Seg h{chunk_h, C};
    B.res_block(&h, 1, C, T, chunk_h, chunk_y);
  }
  // ---- up path (unet.py:305-317): level j consumes the skip of down level j, always upsamples
  for (int j = nl - 1; j >= 1; --j) {
    const int Cs = widths[j + 1], Co = widths[j];
    const Seg in2[2] = {{chunk_h, Cs}, {chunk_skip[j], Cs}};
    B.res_block(in2, 2, Co, T, chunk_h, chunk_y);
    const Seg h{chunk_h, Co};
    B.res_block(&h, 1, Co, T, chunk_h, chunk_y);
    ULayer L = B.blank();  // Upsample1d: ConvTranspose1d(Co, Co, 4, 2, 1)
    L.n_gemm = 1, L.g[0] = B.conv_gemm(B.pc, U_PACK_CONVT, &h, 1, T, Co, T * 2, 4, 2, 1, 0, 0);
    const int nf = Co * T * 2, nfp = round_up(nf, 128);
    L.kind = U_EPI_OPERAND, L.acc_tile = 0, L.mt = nfp / 128, L.nf = nf;
    L.bias_off = B.side_copy(B.pc + 1, Co, T * 2, 0, nfp);
    L.dst_chunk = chunk_h;
    P.layers.push_back(L);
    B.pc += 2;
    T *= 2;
  }
  // ---- final_conv (unet.py:226-237): Conv1dBlock(dim, dim) then Conv1d(dim, Da, 1); rows of the last map are
  // time-major (t * Da + d) = the flat sample index the posterior step uses
  {
    const int C = widths[1];
    const Seg h{chunk_h, C};
    ULayer L = B.blank();
    L.n_gemm = 1, L.g[0] = B.conv_gemm(B.pc, U_PACK_CONV, &h, 1, T, C, T, d.kernel_size, 1, d.kernel_size / 2, 0, 0);
    L.kind = U_EPI_OPERAND;
    B.conv_block_epi(L, B.pc, C, T);
    L.dst_chunk = chunk_y;
    P.layers.push_back(L);
    B.pc += 4;
    const Seg y{chunk_y, C};
    ULayer O = B.blank();
    O.n_gemm = 1, O.g[0] = B.conv_gemm(B.pc, U_PACK_CONV, &y, 1, T, d.action_dim, T, 1, 1, 0, 1, 0);
    O.kind = U_EPI_EPS, O.acc_tile = 0, O.mt = 1, O.nf = P.D;
    O.bias_off = B.side_copy(B.pc + 1, d.action_dim, T, 1, 128);
    P.layers.push_back(O);
    B.pc += 2;
  }
  if (B.err) return B.err;
  if (T != Ta) return set_error("unet: internal: final length %d != horizon %d", T, Ta), DPPO_ERR_INVALID;
  P.n_params = B.pc;
  P.n_side = B.side;
  P.n_tiles = B.tile;
  // Main-path layers alternate between two accumulator sets (TMEM tiles [0, 2 MTmax) and [2 MTmax, 4 MTmax)): the MMAs
  // of layer l+1 may then start while the epilogue of layer l is still reading its accumulators.  Each main-path layer
  // also records which chunks its predecessor writes.
  {
    int n_main = 0, prev = -1, last_main = -1;
    for (size_t i = 0; i < P.layers.size(); ++i)
      if (P.layers[i].track == 0) last_main = int(i);
    prev = last_main;  // cyclic: the output layer (posterior) precedes the first layer of the next step
    for (size_t i = 0; i < P.layers.size(); ++i) {
      ULayer& L = P.layers[i];
      if (L.track != 0) continue;
      if (n_main & 1) {
        const int base = 2 * P.MTmax;
        L.acc_tile += base;
        if (L.res == U_RES_ACC) L.res_acc_tile += base;
        for (int gi = 0; gi < L.n_gemm; ++gi) L.g[gi].acc_tile = uint16_t(L.g[gi].acc_tile + base);
      }
      const ULayer& Q = P.layers[prev];
      if (Q.kind == U_EPI_EPS) {
        L.wait_chunk = P.chunk_x, L.wait_tiles = 1;
      } else {
        L.wait_chunk = Q.dst_chunk, L.wait_tiles = Q.mt;
      }
      prev = int(i);
      ++n_main;
    }
  }
  for (const ULayer& L : P.layers) {
    ++P.track_layers[L.track];
    for (int gi = 0; gi < L.n_gemm; ++gi) P.track_tiles[L.track] += size_t(L.g[gi].mt) * L.g[gi].kc * P.nsplit;
  }
  if (4 * P.MTmax * 16 > 512) return set_error("unet: %d features per activation exceed the TMEM budget", P.MTmax * 128), DPPO_ERR_UNSUPPORTED;
  return DPPO_OK;
}

// ---------------------------------------------------------------------------------------------- packing kernels
// grid (ceil(max threads / 256), n_jobs): one thread per (output row, 8-column group) of job blockIdx.y
__global__ void unet_pack_tiles_kernel(const UPackJob* __restrict__ jobs, const float* const* __restrict__ params,
                                       int nsplit, uint8_t* __restrict__ tiles) {
  const UPackJob j = jobs[blockIdx.y];
  const int groups_per_row = j.KC * 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= j.MT * 128 * groups_per_row) return;
  const float* W = params[j.param];
  const int r = idx / groups_per_row, cgx = idx % groups_per_row;
  const int kc = cgx >> 3, cg = cgx & 7;
  const int mt = r >> 7, ri = r & 127;
  __align__(16) __nv_bfloat16 hi[8];
  __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) split_bf16(unet_dense_at(j, W, r, kc * 64 + cg * 8 + q), hi[q], lo[q]);
  uint8_t* tile_hi = tiles + (size_t(j.tile_off) + (size_t(mt) * j.KC + kc) * nsplit) * 16384;
  const uint32_t off = ri * 128u + (uint32_t(cg ^ (ri & 7)) << 4);
  *reinterpret_cast<uint4*>(tile_hi + off) = *reinterpret_cast<const uint4*>(hi);
  if (nsplit == 2) *reinterpret_cast<uint4*>(tile_hi + 16384 + off) = *reinterpret_cast<const uint4*>(lo);
}

__device__ __forceinline__ float unet_act_exact(int act, float x) {
  if (act == DPPO_ACT_RELU) return fmaxf(x, 0.f);
  const float sp = x > 20.f ? x : log1pf(expf(x));  // torch: softplus threshold 20
  return x * tanhf(sp);
}

// grid (ceil(max pad_to / 256), n_side_jobs); TIME jobs are handled by unet_time_table_kernel
__global__ void unet_side_copy_kernel(const USideJob* __restrict__ jobs, const float* const* __restrict__ params,
                                      float* __restrict__ side) {
  const USideJob j = jobs[blockIdx.y];
  if (j.kind != U_SIDE_COPY) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.pad_to) return;
  const int rep = j.lin_in, tmajor = j.act_first;
  float v = 0.f;
  if (i < j.n * rep) v = params[j.param][tmajor ? i % j.n : i / rep];
  side[j.dst + i] = v;
}

// grid (K, n_side_jobs), 128 threads: time_mlp(t) = Lin(4e,e)(Mish(Lin(e,4e)(sinusoidal(t)))) (unet.py:145-150,
// modules.py:14-27), then row t of the per-timestep bias table of one FiLM Linear
__global__ void unet_time_table_kernel(const USideJob* __restrict__ jobs, const float* const* __restrict__ params,
                                       int e, int act, float* __restrict__ side) {
  const USideJob j = jobs[blockIdx.y];
  if (j.kind != U_SIDE_TIME) return;
  __shared__ float s_emb[32], s_hid[128], s_out[32];
  const int t = blockIdx.x, half = e / 2;
  const float *tw1 = params[0], *tb1 = params[1], *tw2 = params[2], *tb2 = params[3];
  if (threadIdx.x < half) {
    const float rate = float(log(10000.0) / double(half - 1));
    const float ph = float(t) * expf(float(threadIdx.x) * -rate);
    s_emb[threadIdx.x] = sinf(ph), s_emb[threadIdx.x + half] = cosf(ph);
  }
  __syncthreads();
  if (threadIdx.x < 4 * e) {
    float acc = 0.f;
    for (int q = 0; q < e; ++q) acc += s_emb[q] * tw1[threadIdx.x * e + q];
    s_hid[threadIdx.x] = unet_act_exact(DPPO_ACT_MISH, acc + tb1[threadIdx.x]);
  }
  __syncthreads();
  if (threadIdx.x < e) {
    float acc = 0.f;
    for (int q = 0; q < 4 * e; ++q) acc += s_hid[q] * tw2[threadIdx.x * 4 * e + q];
    acc += tb2[threadIdx.x];
    s_out[threadIdx.x] = j.act_first ? unet_act_exact(act, acc) : acc;
  }
  __syncthreads();
  const float *W = params[j.param], *b = params[j.param + 1];
  for (int f = threadIdx.x; f < j.pad_to; f += blockDim.x) {
    float acc = 0.f;
    if (f < j.n) {
      for (int q = 0; q < e; ++q) acc += W[(size_t)f * j.lin_in + q] * s_out[q];
      acc += b[f];
    }
    side[j.dst + (size_t)t * j.pad_to + f] = acc;
  }
}

int pack_unet_impl(dppo_ctx* ctx, int which, const float* const* p, int n_params, cudaStream_t st) {
  const UnetPlan& P = *ctx->unet;
  if (n_params != P.n_params) {
    set_error("dppo_pack_unet: expected %d parameter tensors for this geometry, got %d", P.n_params, n_params);
    return DPPO_ERR_INVALID;
  }
  PackedNet& net = ctx->nets[which];
  const float** d_params = ctx->d_unet_params[which];
  DPPO_CUDA(cudaMemcpyAsync(d_params, p, sizeof(float*) * n_params, cudaMemcpyHostToDevice, st));
  int max_threads = 0, max_pad = 0;
  for (const UPackJob& j : P.jobs) max_threads = max_threads > j.MT * 128 * j.KC * 8 ? max_threads : j.MT * 128 * j.KC * 8;
  for (const USideJob& j : P.side_jobs) max_pad = max_pad > j.pad_to ? max_pad : j.pad_to;
  unet_pack_tiles_kernel<<<dim3((max_threads + 255) / 256, unsigned(P.jobs.size())), 256, 0, st>>>(ctx->d_unet_jobs, d_params, P.nsplit,
                                                                                             net.tiles);
  unet_side_copy_kernel<<<dim3((max_pad + 255) / 256, unsigned(P.side_jobs.size())), 256, 0, st>>>(ctx->d_unet_side_jobs, d_params,
                                                                                                net.side);
  unet_time_table_kernel<<<dim3(P.K, unsigned(P.side_jobs.size())), 128, 0, st>>>(ctx->d_unet_side_jobs, d_params, P.e, P.d.activation,
                                                                                 net.side);
  DPPO_CUDA(cudaGetLastError());
  net.packed = true;
  return DPPO_OK;
}

}  // namespace dppo

using namespace dppo;

// ---------------------------------------------------------------------------------------------- host-only inspection
// (not in the public header; used by tests/test_unet_plan.py to validate the lowering against the oracle on CPU)
extern "C" int dppo_unet_plan_create(const dppo_unet_desc* d, int K, int precision, void** out) {
  if (!d || !out) return set_error("dppo_unet_plan_create: null argument"), DPPO_ERR_INVALID;
  UnetPlan* P = new UnetPlan();
  const int rc = unet_build_plan(*d, K, precision, P);
  if (rc != DPPO_OK) {
    delete P;
    return rc;
  }
  *out = P;
  return DPPO_OK;
}
extern "C" int dppo_unet_plan_destroy(void* plan) {
  delete static_cast<UnetPlan*>(plan);
  return DPPO_OK;
}
// info[0..15]: n_layers, n_jobs, n_side, total_chunks, chunk_x, KX, chunk_state, KS, chunk_state_act, film_dim, MTmax,
//              n_params, n_tiles, sizeof(ULayer), macs_dense (low 31 bits), nsplit
extern "C" int dppo_unet_plan_info(void* plan, int64_t* info) {
  const UnetPlan& P = *static_cast<UnetPlan*>(plan);
  const int64_t v[16] = {int64_t(P.layers.size()), int64_t(P.jobs.size()), int64_t(P.n_side), P.total_chunks, P.chunk_x, P.KX,
                         P.chunk_state, P.KS, P.chunk_state_act, P.film_dim, P.MTmax, P.n_params, int64_t(P.n_tiles),
                         int64_t(sizeof(ULayer)), int64_t(P.macs_dense), P.nsplit};
  memcpy(info, v, sizeof(v));
  return DPPO_OK;
}
extern "C" int dppo_unet_plan_layers(void* plan, void* out) {
  const UnetPlan& P = *static_cast<UnetPlan*>(plan);
  memcpy(out, P.layers.data(), P.layers.size() * sizeof(ULayer));
  return DPPO_OK;
}
// dense [MT*128, KC*64] fp32 of job `job` from HOST parameter pointers
extern "C" int dppo_unet_plan_dense(void* plan, int job, const float* const* host_params, float* out) {
  const UnetPlan& P = *static_cast<UnetPlan*>(plan);
  if (job < 0 || job >= int(P.jobs.size())) return set_error("dppo_unet_plan_dense: job %d", job), DPPO_ERR_INVALID;
  const UPackJob& j = P.jobs[job];
  const float* W = host_params[j.param];
  const int K = j.KC * 64;
  for (int r = 0; r < j.MT * 128; ++r)
    for (int k = 0; k < K; ++k) out[(size_t)r * K + k] = unet_dense_at(j, W, r, k);
  return DPPO_OK;
}
static float host_act(int act, float x) {
  if (act == DPPO_ACT_RELU) return x > 0.f ? x : 0.f;
  const float sp = x > 20.f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
// the fp32 side table from HOST parameter pointers (same arithmetic as the packing kernels)
extern "C" int dppo_unet_plan_side(void* plan, const float* const* hp, float* side) {
  const UnetPlan& P = *static_cast<UnetPlan*>(plan);
  const int e = P.e, half = e / 2;
  for (const USideJob& j : P.side_jobs) {
    if (j.kind == U_SIDE_COPY) {
      const int rep = j.lin_in, tmajor = j.act_first;
      for (int i = 0; i < j.pad_to; ++i) side[j.dst + i] = i < j.n * rep ? hp[j.param][tmajor ? i % j.n : i / rep] : 0.f;
      continue;
    }
    for (int t = 0; t < P.K; ++t) {
      float emb[32], hid[128], outv[32];
      const float rate = float(log(10000.0) / double(half - 1));
      for (int q = 0; q < half; ++q) {
        const float ph = float(t) * expf(float(q) * -rate);
        emb[q] = sinf(ph), emb[q + half] = cosf(ph);
      }
      for (int r = 0; r < 4 * e; ++r) {
        float acc = 0.f;
        for (int q = 0; q < e; ++q) acc += emb[q] * hp[0][r * e + q];
        hid[r] = host_act(DPPO_ACT_MISH, acc + hp[1][r]);
      }
      for (int r = 0; r < e; ++r) {
        float acc = 0.f;
        for (int q = 0; q < 4 * e; ++q) acc += hid[q] * hp[2][r * 4 * e + q];
        acc += hp[3][r];
        outv[r] = j.act_first ? host_act(P.d.activation, acc) : acc;
      }
      for (int f = 0; f < j.pad_to; ++f) {
        float acc = 0.f;
        if (f < j.n) {
          for (int q = 0; q < e; ++q) acc += hp[j.param][(size_t)f * j.lin_in + q] * outv[q];
          acc += hp[j.param + 1][f];
        }
        side[j.dst + (size_t)t * j.pad_to + f] = acc;
      }
    }
  }
  return DPPO_OK;
}
