// Persistent denoise-chain kernel for the Unet1D denoiser (sm_100a: tcgen05.mma + TMEM + bulk TMA copies).
//
// Replaces, for one tile of NE environments per CTA and ALL S denoising steps in one launch:
//   VPGDiffusion.forward / p_mean_var / get_logprobs   reference dppo/model/diffusion/diffusion_vpg.py:139-396
//   Unet1D.forward + ResidualBlock1D                    reference dppo/model/diffusion/unet.py:27-118,267-327
//   Conv1dBlock / Downsample1d / Upsample1d             reference dppo/model/diffusion/modules.py:30-95
//
// The network arrives as a program of dense layers (unet_plan.h): every conv over the short action horizon was
// lowered at pack time to the banded block-Toeplitz matrix it is.  Per layer: 1-2 swap-AB GEMMs (weights = MMA A
// operand streamed from L2 through a ring of 16 KiB pre-swizzled tiles, activations of the NE environments = B operand
// resident in shared memory as bf16 hi [+ lo]) accumulate into TMEM, then the epilogue warps apply bias, GroupNorm
// (a group is a run of <= 32 consecutive features = TMEM lanes of one warp: shuffle reductions, no shared memory),
// activation, FiLM (scale / bias per channel from this block's conditioning encoder, kept in a small fp32 buffer) and
// the residual, and write the next operand.  Same warp roles and barrier protocol as chain_mlp.cu.
//
// Track-split CTA pairs (C = 2): the FiLM conditioning encoders (24 of the 44 layers of cfg5, 40 % of the weights)
// depend on the timestep and the observation only, never on x.  A cluster of two CTAs shares one tile of environments:
// rank 0 runs the main path (x -> eps -> posterior), rank 1 runs the encoders, up to two blocks ahead and across step
// boundaries, and stores each block's (scale, bias) vectors straight into a double-buffered fp32 FiLM buffer in rank
// 0's shared memory (st.shared::cluster + releasing remote mbarrier arrive).  Each CTA streams only its own track's
// weights, so the per-SM weight ingest - the bound of this kernel family - is split between two SMs.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "unet_plan.h"

namespace dppo {

namespace {

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr uint32_t kTile = 16384;
constexpr int kMaxStages = 10;

struct UArgs {
  int D, Da, Ta, Dc, nsplit, nstage;
  int C;  // 1 = one CTA runs both tracks, 2 = track-split CTA pair
  int n_layers;
  const ULayer* layers;
  const uint8_t* tiles[2];
  const float* side[2];
  int chunk_x, chunk_state, chunk_state_act, KS, total_chunks, film_dim;
  int gn_blocks;  // 32-lane feature blocks of the widest layer when some GroupNorm group is not a power of two, else 0
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
  int* nonfinite;            // OR-ed with 1 when a final action element is NaN / Inf
};

// Philox4x32-10 keyed exactly like chain_mlp.cu (same draws for the same (seed, offset, element, slot))
__device__ __forceinline__ float philox_normal_u(uint64_t seed, uint64_t offset, uint64_t elem, uint32_t slot) {
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
__device__ __forceinline__ float act_f(float x) {
  if (ACT == DPPO_ACT_RELU) return fmaxf(x, 0.f);
  return mish_f(x);
}

struct USmem {
  uint8_t *op_hi, *op_lo, *ring;
  float *film, *eps, *gn_part;
  uint64_t *full, *empty, *layer_done, *x_full, *xt_full, *film_full, *film_free;
  uint32_t* tmem_slot;
  uint32_t film_stride;  // floats between the two FiLM buffers
};

// segmented GroupNorm scratch: per 32-lane feature block {head, tail} x {sum, sum of squares} x NE environments
__host__ __device__ inline size_t ugn_part_bytes(int NE, int gn_blocks) { return size_t(gn_blocks) * 4 * NE * sizeof(float); }

__host__ __device__ inline size_t usmem_fixed_bytes(int NE, int total_chunks, int nsplit, int film_dim, int D, int C,
                                                    int gn_blocks) {
  const size_t op = size_t(total_chunks) * NE * 128 * nsplit;
  const size_t film = C * ((size_t(NE) * film_dim * 4 + 127) & ~size_t(127));
  const size_t eps = (size_t(NE) * D * 4 + 127) & ~size_t(127);
  return op + film + eps + ugn_part_bytes(NE, gn_blocks) + 16 * kMaxStages + 160 + 1024 /* alignment slack */;
}

template <int NE>
__device__ __forceinline__ USmem ucarve(uint8_t* base, const UArgs& a) {
  USmem s;
  const size_t opb = size_t(a.total_chunks) * NE * 128;
  uint8_t* p = base;
  s.op_hi = p, p += opb;
  s.op_lo = p, p += (a.nsplit == 2 ? opb : 0);
  s.ring = p, p += size_t(a.nstage) * kTile;
  const size_t film_bytes = (size_t(NE) * a.film_dim * 4 + 127) & ~size_t(127);
  s.film = reinterpret_cast<float*>(p), p += film_bytes * a.C;
  s.film_stride = uint32_t(film_bytes / 4);
  s.eps = reinterpret_cast<float*>(p), p += (size_t(NE) * a.D * 4 + 127) & ~size_t(127);
  s.gn_part = reinterpret_cast<float*>(p), p += ugn_part_bytes(NE, a.gn_blocks);
  s.full = reinterpret_cast<uint64_t*>(p), p += 8 * kMaxStages;
  s.empty = reinterpret_cast<uint64_t*>(p), p += 8 * kMaxStages;
  s.layer_done = reinterpret_cast<uint64_t*>(p), p += 8;
  s.x_full = reinterpret_cast<uint64_t*>(p), p += 8;
  s.xt_full = reinterpret_cast<uint64_t*>(p), p += 64;  // per-M-tile hand-off of the main path (track-split pairs)
  s.film_full = reinterpret_cast<uint64_t*>(p), p += 16;
  s.film_free = reinterpret_cast<uint64_t*>(p), p += 16;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

// ============================================================================================== the kernel
// SEG: the segmented GroupNorm path (group sizes that are not powers of two) is compiled in; the SEG = false
// instantiations are the kernels every power-of-two geometry runs (same registers / code as before the path existed)
template <int NE, int ACT, bool SEG>
__global__ void __launch_bounds__(kThreads, 1) chain_unet_kernel(const UArgs a) {
  constexpr int CPT = NE / 2;  // accumulator columns (environments) per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const USmem s = ucarve<NE>(smem, a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = a.nsplit == 2;
  const int C = a.C;
  const int rank = C > 1 ? int(cluster_ctarank()) : 0;
  const int my_track = C > 1 ? rank : -1;  // -1: this CTA runs every layer
  const int env0 = (blockIdx.x / C) * NE;
  // Main-path CTA of a pair: its layers form a strict chain (every layer consumes its predecessor's output), so the
  // operand is handed over M tile by M tile (xt_full[t]) and consecutive layers use different accumulator sets
  // (unet_plan.cu): the MMAs of layer l+1 start on tile 0 while the epilogue of layer l is still writing tile 1.
  // The encoder CTA and the single-CTA mode keep the whole-operand hand-off (x_full): their layers are not a chain.
  const bool tiled = C > 1 && rank == 0;
  constexpr uint32_t kChunk = NE * 128u;  // bytes of one 64-feature operand chunk (one half)

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstage; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.layer_done, 1);
    mbar_init(s.x_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&s.xt_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.film_full[i], 1);  // the encoder CTA announces the byte count; its values arrive as async stores
      mbar_init(&s.film_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (C > 1) cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  const uint32_t tmem = *s.tmem_slot;

  if (warp == 0) {
    // ======================================================================================= weight-tile producer
    // walks the same layer list as the MMA warp; a GEMM's tiles are contiguous (m-tile major, k-chunk minor, hi then lo)
    uint32_t stage = 0, phase = 0;
    long long p_wait = 0;
    const long long p_t0 = clock64();
    const uint32_t p_leader = elect_one() ? 1u : 0u;  // copies are predicated on it: no divergent region per tile
    for (int step = a.first_step; step < a.S; ++step) {
      const int net = (a.rows[step].ft && !a.use_base) ? 1 : 0;
      for (int li = 0; li < a.n_layers; ++li) {
        const ULayer* L = a.layers + li;
        if (my_track >= 0 && L->track != my_track) continue;
        const int n_gemm = L->n_gemm;
        for (int gi = 0; gi < n_gemm; ++gi) {
          const UGemm G = L->g[gi];
          const uint8_t* src = a.tiles[net] + size_t(G.tile_off) * kTile;
          // chunk-major (every output tile consumes K chunk kc before the next chunk is touched); in memory the tiles of
          // a GEMM are m-tile major, k-chunk minor, hi then lo
          for (uint32_t kc = 0; kc < uint32_t(G.kc); ++kc)
            for (uint32_t mt = 0; mt < uint32_t(G.mt); ++mt)
              for (uint32_t h = 0; h < uint32_t(a.nsplit); ++h) {
                const size_t i = (size_t(mt) * G.kc + kc) * a.nsplit + h;
#ifdef DPPO_CHAIN_PROF
                const long long tw = clock64();
#endif
                mbar_wait(&s.empty[stage], phase ^ 1);
#ifdef DPPO_CHAIN_PROF
                p_wait += clock64() - tw;
#endif
                bulk_g2s_expect_p(s.ring + size_t(stage) * kTile, src + i * kTile, kTile, &s.full[stage], p_leader);
                if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
              }
        }
      }
    }
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 0] = p_wait, a.prof[blockIdx.x * 16 + 1] = clock64() - p_t0;
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    const uint32_t idesc = umma_idesc_bf16(128, NE);
    uint32_t stage = 0, phase = 0, xr_phase = 0, xt_phase = 0;
    const uint32_t ring_lo = umma_desc_lo(smem_u32(s.ring));
    const uint32_t op_hi = umma_desc_lo(smem_u32(s.op_hi)), op_lo = umma_desc_lo(smem_u32(s.op_lo));
    long long m_wait_x = 0, m_wait_full = 0;
    const long long m_t0 = clock64();
    // Same issue discipline as chain_mlp.cu: no divergent region around the MMAs (they are predicated on the lane
    // elected once, here) and the wait counters exist only in a -DDPPO_CHAIN_PROF build.
    const uint32_t leader = elect_one() ? 1u : 0u;
#ifdef DPPO_CHAIN_PROF
#define DPPO_UMMA_T0() tw = clock64()
#define DPPO_UMMA_T1(acc_) acc_ += clock64() - tw
#else
#define DPPO_UMMA_T0()
#define DPPO_UMMA_T1(acc_)
#endif
    for (int step = a.first_step; step < a.S; ++step) {
      for (int li = 0; li < a.n_layers; ++li) {
        const ULayer* L = a.layers + li;
        if (my_track >= 0 && L->track != my_track) continue;
        const int n_gemm = L->n_gemm;
        const UGemm G0 = L->g[0], G1 = L->g[1];  // fetched while the previous epilogue is still running
        const uint32_t wait_chunk = uint32_t(L->wait_chunk), wait_tiles = uint32_t(L->wait_tiles);
        uint32_t waited = 0;
        [[maybe_unused]] long long tw = 0;
        DPPO_UMMA_T0();
        if (!tiled) {
          mbar_wait(s.x_full, xr_phase);
          xr_phase ^= 1;
          tc_fence_after();
        }
        DPPO_UMMA_T1(m_wait_x);
        for (int gi = 0; gi < n_gemm; ++gi) {
          const UGemm G = gi == 0 ? G0 : G1;
          for (int kc = 0; kc < int(G.kc); ++kc) {
            const uint32_t chunk = kc < int(G.src_n[0]) ? uint32_t(G.src_chunk[0]) + kc
                                                        : uint32_t(G.src_chunk[1]) + (kc - int(G.src_n[0]));
            if (tiled && chunk >= wait_chunk && chunk < wait_chunk + 2 * wait_tiles) {
              const uint32_t t = (chunk - wait_chunk) >> 1;  // tile of the predecessor's output this chunk belongs to
              if (!((waited >> t) & 1u)) {
                DPPO_UMMA_T0();
                mbar_wait(&s.xt_full[t], (xt_phase >> t) & 1u);
                DPPO_UMMA_T1(m_wait_x);
                tc_fence_after();
                waited |= 1u << t;
              }
            }
            const uint32_t boff = chunk * (kChunk / 16);
            const uint32_t bh = op_hi + boff, bl = op_lo + boff;
            for (int mt = 0; mt < int(G.mt); ++mt) {
              const uint32_t d = tmem + uint32_t(G.acc_tile + mt) * NE;
              DPPO_UMMA_T0();
              mbar_wait(&s.full[stage], phase);
              DPPO_UMMA_T1(m_wait_full);
              tc_fence_after();
              {
                const uint32_t wa = ring_lo + stage * (kTile / 16);
                const uint32_t rel = smem_u32(&s.empty[stage]);
                umma_bf16_lo_x4_p(d, wa, bh, idesc, kc > 0 ? 1u : 0u, leader, split ? 0u : rel);
                if (split) umma_bf16_lo_x4_p(d, wa, bl, idesc, 1u, leader, rel);
              }
              if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
              if (split) {
                DPPO_UMMA_T0();
                mbar_wait(&s.full[stage], phase);
                DPPO_UMMA_T1(m_wait_full);
                tc_fence_after();
                umma_bf16_lo_x4_p(d, ring_lo + stage * (kTile / 16), bh, idesc, 1u, leader, smem_u32(&s.empty[stage]));
                if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
              }
            }
          }
        }
        if (tiled) {
          // keep the tile barriers in step even if this layer did not read every tile of its predecessor's output
          for (uint32_t t = 0; t < wait_tiles; ++t)
            if (!((waited >> t) & 1u)) {
              mbar_wait(&s.xt_full[t], (xt_phase >> t) & 1u);
              waited |= 1u << t;
            }
          xt_phase ^= waited;
        }
        umma_commit_p(s.layer_done, leader);
      }
    }
    if (a.prof && lane == 0) {
      a.prof[blockIdx.x * 16 + 2] = m_wait_x, a.prof[blockIdx.x * 16 + 3] = m_wait_full;
      a.prof[blockIdx.x * 16 + 4] = clock64() - m_t0;
    }
  } else {
    // ======================================================================================= epilogue warps
    const int et = threadIdx.x - 64;   // 0..255
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the NE columns
    const int fl = q * 32 + lane;      // feature (TMEM lane) within an m-tile
    const int col0 = half * CPT;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t odd = lane & 1;
    uint32_t ld_phase = 0;
    uint32_t film_n = 0;  // FiLM blocks produced (encoder CTA) / consumed (main CTA) so far; buffer = film_n & 1
    float xreg[CPT];
    float zreg[CPT];  // the step's injected noise / stored next sample, fetched while the output layer's MMAs run
    const int nxe = NE * a.D;  // sample elements of this tile (<= 256 * CPT because D <= 128)

    long long e_wait = 0, e_film = 0;
    const long long e_t0 = clock64();
    auto wait_layer = [&]() {
      const long long tw = clock64();
      mbar_wait(s.layer_done, ld_phase);
      e_wait += clock64() - tw;
      ld_phase ^= 1;
      tc_fence_after();
    };
    // whole-operand hand-off, or (main path of a pair) release of tile `t` of this layer's output
    auto signal_x = [&](int t) {
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, kEpiThreads);
      if (et == 0) mbar_arrive(tiled ? &s.xt_full[t] : s.x_full);
    };
    // operand store of one value (prologue / posterior; the layer epilogues use packed pairs)
    auto store_op = [&](int chunk0, int row, int k, float v) {
      const uint32_t off = uint32_t(chunk0) * kChunk + sw128_offset(uint32_t(row), uint32_t(k), NE);
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      *reinterpret_cast<__nv_bfloat16*>(s.op_hi + off) = h;
      if (split) *reinterpret_cast<__nv_bfloat16*>(s.op_lo + off) = __float2bfloat16_rn(v - __bfloat162float(h));
    };
    // flat sample index (t * Da + d, the reference's (Ta, Da) layout) -> channel-major operand feature d * Ta + t
    auto x_feature = [&](int f) { return (f % a.Da) * a.Ta + f / a.Da; };

    // ------------------------------------------------------------------------------------ prologue
    {
      const uint32_t op_bytes = uint32_t(a.total_chunks) * kChunk;
      for (uint32_t i = et * 16; i < op_bytes; i += kEpiThreads * 16) {
        *reinterpret_cast<uint4*>(s.op_hi + i) = make_uint4(0, 0, 0, 0);
        if (split) *reinterpret_cast<uint4*>(s.op_lo + i) = make_uint4(0, 0, 0, 0);
      }
      named_bar_sync(1, kEpiThreads);
      for (int i = et; i < NE * a.Dc; i += kEpiThreads) {
        const int e = i / a.Dc, k = i % a.Dc;
        const int env = env0 + e;
        const float v = env < a.E ? a.state[size_t(env) * a.Dc + k] : 0.f;
        store_op(a.chunk_state, e, k, v);
        if (a.chunk_state_act >= 0) store_op(a.chunk_state_act, e, k, act_f<ACT>(v));
      }
#pragma unroll 1
      for (int j = 0; j < CPT; ++j) {
        const int i = et + j * kEpiThreads;
        float x = 0.f;
        if (i < nxe && rank == 0) {
          const int e = i / a.D, f = i - e * a.D;
          const int env = env0 + e;
          if (env < a.E) {
            if (a.eval_mode)
              x = a.chains_in[(size_t(env) * (a.ft + 1)) * a.D + f];
            else if (a.noise)
              x = a.noise[size_t(env) * a.D + f];
            else
              x = philox_normal_u(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + f, 0u);
            if (!a.eval_mode && a.chain && a.ft == a.S) a.chain[(size_t(env) * (a.ft + 1)) * a.D + f] = x;
          }
          store_op(a.chunk_x, e, x_feature(f), x);
        }
        xreg[j] = x;
      }
      signal_x(0);
    }

    // ------------------------------------------------------------------------------------ step loop
    for (int step = a.first_step; step < a.S; ++step) {
      const StepRow row = a.rows[step];
      const int net = (row.ft && !a.use_base) ? 1 : 0;
      const float* side = a.side[net];
      for (int li = 0; li < a.n_layers; ++li) {
        const ULayer* L = a.layers + li;
        if (my_track >= 0 && L->track != my_track) continue;
        const int kind = L->kind, acc_tile = L->acc_tile, MTl = L->mt, nf = L->nf;
        const float* bias = side + L->bias_off + size_t(L->bias_tstride) * row.t;
        const int gs = L->gn_size, do_act = L->act, film = L->film, film_c = L->film_c, tshift = L->film_tshift;
        const int res = L->res;
        const float* gamma = side + L->gamma_off;
        const float* beta = side + L->beta_off;
        const float gn_eps = L->gn_eps;
        const float* res_bias = side + L->res_bias_off;
        // per-feature constants of the first two M tiles are fetched (L2 latency) while the MMAs of this layer still run
        float pb[2], pg[2], pbe[2], prb[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = (i < MTl ? i : 0) * 128 + fl;
          pb[i] = bias[f], pg[i] = gamma[f], pbe[i] = beta[f], prb[i] = res_bias[f];
        }
        if (kind != U_EPI_OPERAND) {
          // the posterior is a serial stretch of the chain; its Philox + Box-Muller draws depend on nothing the MMAs
          // produce, so they happen here, where this warp would otherwise wait
#pragma unroll 1
          for (int j = 0; j < CPT; ++j) {
            const int i = et + j * kEpiThreads;
            if (i >= nxe) break;
            const int e = i / a.D, f = i - e * a.D;
            const int env = env0 + e;
            float z = 0.f;
            if (env < a.E && a.eval_mode) {
              z = a.chains_in[(size_t(env) * (a.ft + 1) + (step - a.first_step) + 1) * a.D + f];
            } else if (env < a.E) {
              if (a.noise)
                z = a.noise[(size_t(step + 1) * a.E + env) * a.D + f];
              else
                z = philox_normal_u(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + f, uint32_t(step + 1));
              z = fminf(fmaxf(z, -a.randn_clip), a.randn_clip);
            }
            zreg[j] = z;
          }
        }
        wait_layer();
        if (kind == U_EPI_OPERAND) {
          const uint32_t dst_base = uint32_t(L->dst_chunk) * kChunk;
          const uint32_t res_base = uint32_t(res == U_RES_SLOT ? L->res_chunk : 0) * kChunk;
          const uint32_t res_tile = uint32_t(L->res_acc_tile);
          const float* film_buf = s.film;
          if (film && C > 1) {
            // this block's (scale, bias) vectors arrive from the encoder CTA
            const long long tw = clock64();
            mbar_wait(&s.film_full[film_n & 1], (film_n >> 1) & 1);
            e_film += clock64() - tw;
            film_buf = s.film + (film_n & 1) * s.film_stride;
          }
          // GroupNorm groups whose size is not a power of two (5 channels x 4 positions = 20 lowered features) straddle
          // warps and M tiles: every warp reduces the runs of its 32 lanes by a segmented shuffle scan and the two pieces
          // of a straddling group meet through shared memory (first pass: partial sums of the runs that touch lane 0 /
          // lane 31 of each 32-feature block; second pass, below: statistics = own run + the neighbour block's piece)
          const bool gn_seg = SEG && gs && (gs & (gs - 1));
          struct SegInfo {
            uint32_t same;     // bit k: lane + 2^k is in this warp and in the same group
            int first;         // first lane of this lane's run
            bool prev, next;   // the group began in the previous 32-feature block / continues in the next one
          };
          auto seg_info = [&](int mt) {
            SegInfo si;
            const int f = mt * 128 + fl, blk0 = f - lane;
            const bool valid = f < nf;
            const int g = valid ? f / gs : -1;
            si.same = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              const int fn = f + (1 << k);
              if (valid && lane + (1 << k) < 32 && fn < nf && fn / gs == g) si.same |= 1u << k;
            }
            si.first = valid ? max(g * gs - blk0, 0) : lane;
            si.prev = valid && g * gs < blk0;
            si.next = valid && (g + 1) * gs > blk0 + 32;
            return si;
          };
          // after the scan the first lane of every run holds the run's sum (runs are at most 32 lanes long)
          auto seg_scan = [&](float (&x)[CPT], const SegInfo& si) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              const bool take = (si.same >> k) & 1u;
#pragma unroll
              for (int c = 0; c < CPT; ++c) {
                const float t = __shfl_down_sync(0xffffffffu, x[c], 1 << k);
                if (take) x[c] += t;
              }
            }
          };
          if (gn_seg) {
            for (int mt = 0; mt < MTl; ++mt) {
              float v[CPT], sq[CPT];
              tmem_ld(tmem + lane_addr + uint32_t(acc_tile + mt) * NE + col0, v);
              const float b = mt < 2 ? pb[mt & 1] : bias[mt * 128 + fl];
#pragma unroll
              for (int c = 0; c < CPT; ++c) v[c] += b, sq[c] = v[c] * v[c];
              const SegInfo si = seg_info(mt);
              seg_scan(v, si);
              seg_scan(sq, si);
              float* part = s.gn_part + size_t(mt * 4 + q) * 4 * NE + col0;
              if (lane == 0 && si.prev) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) part[c] = v[c], part[NE + c] = sq[c];
              }
              if (lane == si.first && si.next) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) part[2 * NE + c] = v[c], part[3 * NE + c] = sq[c];
              }
            }
            named_bar_sync(1, kEpiThreads);
          }
          for (int mt = 0; mt < MTl; ++mt) {
            float v[CPT];
            tmem_ld(tmem + lane_addr + uint32_t(acc_tile + mt) * NE + col0, v);
            const int f = mt * 128 + fl;
            const float b = mt < 2 ? pb[mt & 1] : bias[f];
#pragma unroll
            for (int c = 0; c < CPT; ++c) v[c] += b;
            if (gn_seg) {
              const float inv = 1.f / float(gs);
              const float g = mt < 2 ? pg[mt & 1] : gamma[f], be = mt < 2 ? pbe[mt & 1] : beta[f];
              const SegInfo si = seg_info(mt);
              float sm[CPT], sq[CPT];
#pragma unroll
              for (int c = 0; c < CPT; ++c) sm[c] = v[c], sq[c] = v[c] * v[c];
              seg_scan(sm, si);
              seg_scan(sq, si);
              const float* part = s.gn_part + size_t(mt * 4 + q) * 4 * NE + col0;
#pragma unroll
              for (int c = 0; c < CPT; ++c) {
                float a1 = __shfl_sync(0xffffffffu, sm[c], si.first), a2 = __shfl_sync(0xffffffffu, sq[c], si.first);
                if (si.prev) a1 += part[c - 4 * NE + 2 * NE], a2 += part[c - 4 * NE + 3 * NE];  // tail of the previous block
                if (si.next) a1 += part[c + 4 * NE], a2 += part[c + 4 * NE + NE];                // head of the next block
                const float mean = a1 * inv, var = fmaxf(a2 * inv - mean * mean, 0.f);
                v[c] = (v[c] - mean) * rsqrtf(var + gn_eps) * g + be;
              }
            } else if (gs) {
              // GroupNorm over gs consecutive features (lanes) of each environment column: mean, then centred variance.
              // Stage-major loops keep CPT independent shuffles in flight per butterfly stage.
              const float inv = 1.f / float(gs);
              const float g = mt < 2 ? pg[mt & 1] : gamma[f], be = mt < 2 ? pbe[mt & 1] : beta[f];
              float sm[CPT];
#pragma unroll
              for (int c = 0; c < CPT; ++c) sm[c] = v[c];
              for (int o = gs >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) sm[c] += __shfl_xor_sync(0xffffffffu, sm[c], o);
              }
#pragma unroll
              for (int c = 0; c < CPT; ++c) v[c] -= sm[c] * inv, sm[c] = v[c] * v[c];
              for (int o = gs >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) sm[c] += __shfl_xor_sync(0xffffffffu, sm[c], o);
              }
#pragma unroll
              for (int c = 0; c < CPT; ++c) v[c] = v[c] * rsqrtf(sm[c] * inv + gn_eps) * g + be;
            }
            if (do_act) {
#pragma unroll
              for (int c = 0; c < CPT; ++c) v[c] = act_f<ACT>(v[c]);
            }
            if (film && f < nf) {
              const float* fr = film_buf + size_t(col0) * a.film_dim + (f >> tshift);
#pragma unroll
              for (int c = 0; c < CPT; ++c) {
                const float* fe = fr + size_t(c) * a.film_dim;
                v[c] = film == 2 ? fe[0] * v[c] + fe[film_c] : v[c] + fe[0];
              }
            }
            if (res == U_RES_ACC) {
              float r[CPT];
              tmem_ld(tmem + lane_addr + (res_tile + uint32_t(mt)) * NE + col0, r);
              const float rb = mt < 2 ? prb[mt & 1] : res_bias[f];
#pragma unroll
              for (int c = 0; c < CPT; ++c) v[c] += r[c] + rb;
            }
            if (f >= nf) {
#pragma unroll
              for (int c = 0; c < CPT; ++c) v[c] = 0.f;
            }
            // neighbouring lanes swap one value per column pair: every thread then owns two consecutive features of one
            // environment row -> one packed bf16x2 convert and one 4-byte store per operand half
            const uint32_t kp = uint32_t(f) & ~1u;
            const uint32_t j16 = (kp & 63u) >> 3;
            const uint32_t rel = (kp >> 6) * kChunk + ((kp & 7u) << 1) + (uint32_t(col0) + odd) * 128u;
#pragma unroll
            for (int c = 0; c < CPT; c += 2) {
              const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[c] : v[c + 1], 1);
              float fa = odd ? recv : v[c], fb = odd ? v[c + 1] : recv;  // features kp, kp + 1 of env row col0 + c + odd
              const uint32_t off = rel + uint32_t(c) * 128u + ((j16 ^ ((uint32_t(c) + odd) & 7u)) << 4);
              if (res == U_RES_SLOT) {
                const __nv_bfloat162 rh = *reinterpret_cast<const __nv_bfloat162*>(s.op_hi + res_base + off);
                fa += __low2float(rh), fb += __high2float(rh);
                if (split) {
                  const __nv_bfloat162 rl = *reinterpret_cast<const __nv_bfloat162*>(s.op_lo + res_base + off);
                  fa += __low2float(rl), fb += __high2float(rl);
                }
              }
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(fa, fb);
              *reinterpret_cast<__nv_bfloat162*>(s.op_hi + dst_base + off) = h2;
              if (split)
                *reinterpret_cast<__nv_bfloat162*>(s.op_lo + dst_base + off) =
                    __floats2bfloat162_rn(fa - __low2float(h2), fb - __high2float(h2));
            }
            if (tiled && mt + 1 < MTl) signal_x(mt);  // tile mt is complete: the next layer's MMAs may consume it
          }
          if (film && C > 1) {
            // every epilogue thread is done reading the buffer: hand it back to the encoder CTA
            named_bar_sync(1, kEpiThreads);
            if (et == 0) mbar_arrive_remote_nodata(&s.film_free[film_n & 1], 1);
            ++film_n;
          }
        } else if (kind == U_EPI_FILM) {
          float* film_buf = s.film;
          if (C > 1) {
            // double-buffered in the MAIN CTA's shared memory; block n may be written once block n - 2 was consumed
            const long long tw = clock64();
            if (film_n >= 2) mbar_wait(&s.film_free[film_n & 1], ((film_n >> 1) - 1) & 1);
            e_film += clock64() - tw;
            film_buf = s.film + (film_n & 1) * s.film_stride;
          }
          for (int mt = 0; mt < MTl; ++mt) {
            float v[CPT];
            tmem_ld(tmem + lane_addr + uint32_t(acc_tile + mt) * NE + col0, v);
            const int f = mt * 128 + fl;
            if (f < nf) {
              const float b = mt < 2 ? pb[mt & 1] : bias[f];
              if (C > 1) {
#pragma unroll
                for (int c = 0; c < CPT; ++c)
                  st_async_f32(film_buf + size_t(col0 + c) * a.film_dim + f, 0, v[c] + b, &s.film_full[film_n & 1]);
              } else {
#pragma unroll
                for (int c = 0; c < CPT; ++c) film_buf[size_t(col0 + c) * a.film_dim + f] = v[c] + b;
              }
            }
          }
          if (C > 1) {
            // nf features x NE environments, 4 bytes each, complete on the main CTA's barrier (async stores: no release
            // fence per thread, which was MEMBAR.ALL.GPU x 256 threads per block)
            if (et == 0) mbar_arrive_expect_tx_remote(&s.film_full[film_n & 1], 0, uint32_t(nf) * NE * 4u);
            ++film_n;
          }
        } else {
          // output layer + posterior (diffusion_vpg.py:165-224, 279-311): eps -> [env][flat element] fp32 tile, then all
          // epilogue threads run the element-wise update on the flat mapping (coalesced global access)
          if (q * 32 < a.D) {
            float v[CPT];
            tmem_ld(tmem + lane_addr + uint32_t(acc_tile) * NE + col0, v);
            if (fl < a.D) {
              const float bo = pb[0];
#pragma unroll
              for (int c = 0; c < CPT; ++c) s.eps[(col0 + c) * a.D + fl] = v[c] + bo;
            }
          }
          named_bar_sync(1, kEpiThreads);
          const bool last = step == a.S - 1;
          const int d_eval = step - a.first_step;
          float stdv, f2 = row.f2, f3 = row.f3;
          if (a.eval_mode) {
            stdv = fmaxf(row.std_train, a.min_std);
          } else if (a.deterministic) {
            f2 = row.f2_det, f3 = row.f3_det;
            stdv = a.use_ddim ? 0.f : (row.t == 0 ? 0.f : fmaxf(row.std_train, 1e-3f));
          } else {
            stdv = fmaxf(row.std_train, a.min_std);
          }
          const float inv_2var = 1.f / (2.f * (stdv * stdv)), log_std = logf(stdv);
#pragma unroll 1
          for (int j = 0; j < CPT; ++j) {
            const int i = et + j * kEpiThreads;
            if (i >= nxe) break;
            const int e = i / a.D, f = i - e * a.D;
            const int env = env0 + e;
            float eps = s.eps[i];
            const float x = xreg[j];
            float x0, mu;
            if (!a.use_ddim) {
              x0 = row.f0 * x - row.f1 * eps;
              if (a.x0_clip >= 0.f) x0 = fminf(fmaxf(x0, -a.x0_clip), a.x0_clip);
              mu = f2 * x0 + f3 * x;
            } else {
              x0 = (x - row.f1 * eps) / row.f0;
              if (a.x0_clip >= 0.f) {
                x0 = fminf(fmaxf(x0, -a.x0_clip), a.x0_clip);
                eps = (x - row.f0 * x0) / row.f1;
              }
              if (a.eps_clip >= 0.f) eps = fminf(fmaxf(eps, -a.eps_clip), a.eps_clip);
              mu = f2 * x0 + f3 * eps;
            }
            float xn = 0.f;
            if (env < a.E) {
              if (a.eval_mode) {
                xn = zreg[j];
                const float diff = xn - mu;
                a.logp[(size_t(env) * a.ft + d_eval) * a.D + f] = -(diff * diff) * inv_2var - log_std - 0.91893853320467274f;
              } else {
                xn = mu + stdv * zreg[j];
                if (last && a.final_clip >= 0.f) xn = fminf(fmaxf(xn, -a.final_clip), a.final_clip);
                if (a.chain && row.slot >= 0) a.chain[(size_t(env) * (a.ft + 1) + row.slot) * a.D + f] = xn;
                if (last) {
                  a.traj[size_t(env) * a.D + f] = xn;
                  if (!(fabsf(xn) <= 3.0e38f)) atomicOr(a.nonfinite, 1);
                }
              }
            }
            xreg[j] = xn;
            store_op(a.chunk_x, e, x_feature(f), xn);
          }
        }
        signal_x(kind == U_EPI_OPERAND ? MTl - 1 : 0);
      }
    }
    if (a.prof && et == 0) a.prof[blockIdx.x * 16 + 5] = e_wait, a.prof[blockIdx.x * 16 + 6] = clock64() - e_t0, a.prof[blockIdx.x * 16 + 7] = e_film;
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 8 + (warp - 2)] = (clock64() - e_t0) - e_wait - e_film;
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (C > 1) cluster_sync_all();  // no CTA leaves while its peer can still signal it or store into it
}

template <int NE, int ACT, bool SEG>
int ulaunch(const UArgs& a, size_t smem_bytes, cudaStream_t st) {
  auto kfn = chain_unet_kernel<NE, ACT, SEG>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(chain_unet_kernel)");
    configured = true;
  }
  const int tiles = (a.E + NE - 1) / NE;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(tiles * a.C)), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(a.C), attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, a);
  if (e != cudaSuccess) return cuda_fail(e, "chain_unet_kernel launch");
  return DPPO_OK;
}

}  // namespace

// Launch shape (NE environments per tile, C = 1 | 2 CTAs per tile).  The kernel is bound by the weight ingest of one SM
// (its whole track, every step, whatever NE is) plus a per-layer epilogue / hand-off latency, so the smallest tile that
// still fits every environment in one wave wins, and the track-split pair wins whenever the pairs fit in one wave.
struct UShape {
  int NE, C, nstage;
};
static UShape pick_unet_shape(const dppo_ctx* ctx, int E) {
  const UnetPlan& P = *ctx->unet;
  static int env_ne = -1, env_c = -1;
  if (env_ne < 0) {
    const char* e = getenv("DPPO_B200_TILE_ENVS");
    env_ne = e ? atoi(e) : 0;
    e = getenv("DPPO_B200_CLUSTER");
    env_c = e ? atoi(e) : 0;
  }
  const int forced_ne = ctx->force_ne ? ctx->force_ne : env_ne, forced_c = ctx->force_c ? ctx->force_c : env_c;
  const size_t budget = 232448;
  UShape best{0, 1, 0};
  double best_t = 1e30;
  for (int C = 1; C <= 2; ++C) {
    if (forced_c > 0 && C != forced_c) continue;
    if (C == 2 && P.track_layers[1] == 0) continue;
    for (int NE = 16; NE <= 64; NE *= 2) {
      if (forced_ne > 0 && NE != forced_ne) continue;
      if (2 * P.MTmax * NE > 512) continue;
      const size_t fixed = usmem_fixed_bytes(NE, P.total_chunks, P.nsplit, P.film_dim, P.D, C, P.gn_segmented ? P.MTmax * 4 : 0);
      if (fixed + 2 * kTile > budget) continue;
      int nstage = int((budget - fixed) / kTile);
      if (nstage > kMaxStages) nstage = kMaxStages;
      const int tiles = (E + NE - 1) / NE;
      const int slots = ctx->sm_count / C;
      const int waves = (tiles + slots - 1) / slots;
      const double rate = nstage >= 3 ? 34.7 : 20.0;
      const double per_layer = double(NE) * P.MTmax * 128 / 256.0 * 37.0 + 4000.0;
      double t;
      if (C == 1) {
        t = double(P.n_tiles) * 16384.0 / rate + double(P.layers.size()) * per_layer;
      } else {
        const double t0 = double(P.track_tiles[0]) * 16384.0 / rate + P.track_layers[0] * per_layer;
        const double t1 = double(P.track_tiles[1]) * 16384.0 / rate + P.track_layers[1] * per_layer;
        t = t0 > t1 ? t0 : t1;
      }
      t *= waves;
      if (t < best_t) best_t = t, best = UShape{NE, C, nstage};
    }
  }
  return best;
}

int sample_chain_unet_impl(dppo_ctx* ctx, const float* state, int E, const float* noise, uint64_t seed, uint64_t offset,
                           int64_t env_offset, int deterministic, int use_base, float min_std, float* traj, float* chain,
                           const float* chains_in, float* logp, cudaStream_t st) {
  const UnetPlan& P = *ctx->unet;
  UArgs a{};
  a.D = P.D, a.Da = P.d.action_dim, a.Ta = P.d.horizon_steps, a.Dc = P.d.cond_dim, a.nsplit = P.nsplit;
  a.n_layers = int(P.layers.size()), a.layers = ctx->d_unet_layers;
  for (int w = 0; w < 2; ++w) a.tiles[w] = ctx->nets[w].tiles, a.side[w] = ctx->nets[w].side;
  a.chunk_x = P.chunk_x, a.chunk_state = P.chunk_state, a.chunk_state_act = P.chunk_state_act, a.KS = P.KS;
  a.total_chunks = P.total_chunks, a.film_dim = P.film_dim;
  a.gn_blocks = P.gn_segmented ? P.MTmax * 4 : 0;
  a.rows = ctx->d_rows, a.S = ctx->S, a.ft = ctx->ft, a.use_ddim = ctx->use_ddim;
  a.eval_mode = chains_in != nullptr;
  a.first_step = a.eval_mode ? ctx->S - ctx->ft : 0;
  a.deterministic = deterministic, a.use_base = use_base;
  a.min_std = min_std, a.x0_clip = ctx->x0_clip, a.randn_clip = ctx->randn_clip, a.final_clip = ctx->final_clip;
  a.eps_clip = ctx->eps_clip;
  a.state = state, a.E = E, a.noise = noise, a.traj = traj, a.chain = chain, a.chains_in = chains_in, a.logp = logp;
  a.seed = seed, a.offset = offset, a.env_offset = env_offset;
  a.prof = ctx->d_prof;
  a.nonfinite = ctx->d_nonfinite;

  const UShape shape = pick_unet_shape(ctx, E);
  const int NE = shape.NE;
  if (NE == 0) return set_error("unet chain kernel: no tile size fits this geometry in shared memory / TMEM"), DPPO_ERR_UNSUPPORTED;
  a.nstage = shape.nstage, a.C = shape.C;
  const size_t smem_bytes = usmem_fixed_bytes(NE, P.total_chunks, P.nsplit, P.film_dim, P.D, a.C, a.gn_blocks) + size_t(a.nstage) * kTile;
#define DPPO_ULAUNCH_A(NE_, ACT_) (a.gn_blocks ? ulaunch<NE_, ACT_, true>(a, smem_bytes, st) : ulaunch<NE_, ACT_, false>(a, smem_bytes, st))
#define DPPO_ULAUNCH(NE_) (P.d.activation == DPPO_ACT_RELU ? DPPO_ULAUNCH_A(NE_, DPPO_ACT_RELU) : DPPO_ULAUNCH_A(NE_, DPPO_ACT_MISH))
  if (NE == 64) return DPPO_ULAUNCH(64);
  if (NE == 32) return DPPO_ULAUNCH(32);
  return DPPO_ULAUNCH(16);
#undef DPPO_ULAUNCH
#undef DPPO_ULAUNCH_A
}

}  // namespace dppo
