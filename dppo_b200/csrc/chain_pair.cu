// Denoise-chain kernel for CTA PAIRS: tcgen05.mma.cta_group::2 (M = 256 over two SMs, N = 64 environments).
//
// Replaces the same reference code as chain_mlp.cu (VPGDiffusion.forward / p_mean_var / get_logprobs,
// dppo/model/diffusion/diffusion_vpg.py:139-396; DiffusionMLP.forward, dppo/model/diffusion/mlp_diffusion.py:218-250)
// for the launch shape the headline workload runs in: 64 environments per cluster of two CTAs, feature split in halves,
// no LayerNorm, no cond_mlp.
//
// What changes against the cta_group::1 kernel of the same shape:
//   * ONE instruction stream for the pair.  The leader CTA's MMA warp issues M = 256 MMAs whose A operand is one weight
//     tile from EACH CTA's ring (CTA r streams the M tiles [r MT/2, (r+1) MT/2) exactly as before) and whose B operand is
//     split by environments: each CTA holds the activations of 32 of the pair's 64 environments.  Each CTA's TMEM
//     receives its own 128 features x all 64 environments - the accumulator layout of the old kernel - but the pair
//     issues half the instructions (43 instead of 2 x 49 cycles per K step, profiles/r1k_microbench_mma_variants.txt).
//   * Half the activation operand per SM (32 x H instead of 64 x H bf16 hi + lo): the freed shared memory deepens the
//     weight ring (7 instead of 5 stages at H = 512), and the kernel is bound by how far the producer can run ahead of the
//     serial stretches of the layer chain.
//   * The epilogue routes by environment half: warps that hold the CTA's own 32 environments write X in place, the other
//     four warps write a staging block that one bulk copy per M tile moves into the peer's X (half the exchange bytes).
//     The posterior step runs on the CTA's own 32 environments only.
//   * No hand-shake before a layer's epilogue: tcgen05.commit multicasts the layer's completion to both CTAs, and a
//     completed layer implies that every block pushed for it has landed and that nobody reads either copy of X.
// The follower CTA's second warp relays what only it can observe (its ring stages and the leader's blocks landing in its
// copy of X) to barriers in the leader's shared memory; its own tiles and its layer-0 operand are announced by its
// epilogue directly.
#include <stdlib.h>

#include "chain_mlp.cuh"

namespace dppo {

namespace {

constexpr int kPN = 64;             // environments per CTA pair = N of every MMA
constexpr int kPL = 32;             // environments (rows of B) per CTA
#ifndef DPPO_PAIR_EPI_WARPS
#define DPPO_PAIR_EPI_WARPS 8
#endif
constexpr int kPEpiWarps = DPPO_PAIR_EPI_WARPS;  // 8 or 16: warps 2.. ; each lane quarter of TMEM is served by kPEpiWarps / 4 warps
constexpr int kPEpiThreads = kPEpiWarps * 32;
constexpr int kPCols = 4 * kPN / kPEpiWarps;     // accumulator columns (environments) per epilogue thread: 32 or 16
constexpr int kPThreads = 64 + kPEpiThreads;
constexpr uint32_t kPTile = 16384;  // one weight tile (128 features x 64 K, bf16, SWIZZLE_128B image)
constexpr uint32_t kPChunk = kPL * 128;  // one 64-wide K chunk of a 32-row activation operand
constexpr int kPMaxStages = 12;
constexpr int kPX = kPL * 128 / kPEpiThreads;  // sample elements per epilogue thread: 32 environments x D <= 128
constexpr size_t kPBarBytes = 512;

struct PSmem {
  uint8_t *x_hi, *x_lo, *x0_hi, *x0_lo, *stg_hi, *stg_lo, *ring;
  uint64_t *full, *empty, *pfull, *layer_done, *x_full, *px_full, *x0_full, *px0_full;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ PSmem pcarve(uint8_t* base, const ChainArgs& a) {
  PSmem s;
  const uint32_t xb = uint32_t(a.KCH) * kPChunk, x0b = uint32_t(a.KC0) * kPChunk, sb = uint32_t(a.MT / 2) * 2u * kPChunk;
  const bool split = a.nsplit == 2;
  uint8_t* p = base;
  s.x_hi = p, p += xb;
  s.x_lo = p, p += split ? xb : 0;
  s.x0_hi = p, p += x0b;
  s.x0_lo = p, p += split ? x0b : 0;
  s.stg_hi = p, p += sb;
  s.stg_lo = p, p += split ? sb : 0;
  s.ring = p, p += size_t(a.nstage) * kPTile;
  s.full = reinterpret_cast<uint64_t*>(p), p += 8 * kPMaxStages;
  s.empty = reinterpret_cast<uint64_t*>(p), p += 8 * kPMaxStages;
  s.pfull = reinterpret_cast<uint64_t*>(p), p += 8 * kPMaxStages;  // leader: the follower's ring stage has landed
  s.layer_done = reinterpret_cast<uint64_t*>(p), p += 8;
  s.x_full = reinterpret_cast<uint64_t*>(p), p += 64;              // per M tile (= K-chunk pair) of this CTA's copy of X
  s.px_full = reinterpret_cast<uint64_t*>(p), p += 64;             // leader: the same tile of the follower's copy
  s.x0_full = reinterpret_cast<uint64_t*>(p), p += 8;
  s.px0_full = reinterpret_cast<uint64_t*>(p), p += 8;
  s.tmem_slot = reinterpret_cast<uint32_t*>(p);
  return s;
}

size_t pair_fixed_bytes(const MlpGeom& g) {
  const size_t xb = size_t(g.KCH) * kPChunk, x0b = size_t(g.KC0) * kPChunk, sb = size_t(g.MT / 2) * 2 * kPChunk;
  return (xb + x0b + sb) * g.nsplit + kPBarBytes + 1024 /* alignment slack */;
}

// ---- cta_group::2 forms of the tcgen05 helpers (one .cta_group per kernel)
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of every MMA issued so far -> the barrier at this CTA-relative address in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2_p(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(leader), "h"(uint16_t(3))
      : "memory");
}
// the four K = 16 steps of one 64-wide swizzle atom in one statement, optional ring release behind them (see
// umma_bf16_lo_x4_p in common.cuh: same issue discipline, pair form)
__device__ __forceinline__ void umma2_bf16_lo_x4_p(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                   uint32_t accumulate_first, uint32_t leader, uint32_t commit_bar) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p, q, t, c;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\tsetp.eq.u32 t, 1, 1;\n\t"
      "setp.ne.and.b32 c, %6, 0, q;\n\t"
      "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
      "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, t;\n\t"
      "@c tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %7;\n\t}" ::"r"(
          d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate_first), "r"(leader), "r"(commit_bar), "h"(uint16_t(3))
      : "memory");
}

// K chunks of X are visited tile by tile, the follower's tile j in front of the leader's tile j: the follower announces its
// own tiles to the leader directly, while a leader tile is usable only once the relay has seen its copy land in the
// follower - that hop hides behind the MMAs of the chunks in front of it.
__device__ __forceinline__ int pair_chunk(int i, int MTo) {
  const int slot = i >> 1;
  const int t = (slot & 1) ? (slot >> 1) : MTo + (slot >> 1);
  return 2 * t + (i & 1);
}

// ============================================================================================== the kernel
template <int ACT>
__global__ void __launch_bounds__(kPThreads, 1) chain_pair_kernel(const ChainArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const PSmem s = pcarve(smem, a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = a.nsplit == 2;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const int env0 = (blockIdx.x >> 1) * kPN + int(rank) * kPL;  // first environment of this CTA's half of the pair's tile
  const int MTo = a.MT / 2;         // M tiles of every hidden layer this CTA streams and post-processes
  const int mt0 = int(rank) * MTo;  // first one

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstage; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
      mbar_init(&s.pfull[i], 1);
    }
    mbar_init(s.layer_done, 1);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&s.x_full[i], 1);
      mbar_init(&s.px_full[i], 1);
    }
    mbar_init(s.x0_full, 1);
    mbar_init(s.px0_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(s.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM are in place before anybody signals or issues
  tc_fence_after();
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t col_h = 0, col_y = uint32_t(MTo) * kPN;

  if (warp == 0) {
    // ======================================================================================= weight-tile producer
    // Both CTAs run the same sequence on their own M tiles, so ring stage i of the two CTAs always holds the two halves of
    // one M = 256 A operand.
    uint32_t stage = 0, phase = 0;
    const uint32_t p_leader = elect_one() ? 1u : 0u;
    [[maybe_unused]] long long p_wait = 0, p_t0 = clock64();
    auto stream = [&](const uint8_t* base, int m_begin, int m_end, int KCl, bool tiled, uint32_t bytes) {
      for (int i = 0; i < KCl; ++i) {
        const int kc = tiled ? pair_chunk(i, MTo) : i;
        for (int mt = m_begin; mt < m_end; ++mt) {
          const uint8_t* src = base + size_t(mt) * KCl * a.nsplit * kPTile;
          for (int h = 0; h < a.nsplit; ++h) {
#ifdef DPPO_CHAIN_PROF
            const long long tw = clock64();
#endif
            mbar_wait(&s.empty[stage], phase ^ 1);
#ifdef DPPO_CHAIN_PROF
            p_wait += clock64() - tw;
#endif
            bulk_g2s_expect_p(s.ring + size_t(stage) * kPTile, src + size_t(kc * a.nsplit + h) * kPTile, bytes,
                              &s.full[stage], p_leader);
            if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
          }
        }
      }
    };
    const uint32_t out_bytes = uint32_t((a.D + 7) / 8) * 8u * 128u;  // the output layer has D <= 128 real rows
    const size_t lin0 = size_t(a.MT) * a.KC0 * a.nsplit * kPTile, linh = size_t(a.MT) * a.KCH * a.nsplit * kPTile;
    for (int step = a.first_step; step < a.S; ++step) {
      const int net = (a.rows[step].ft && !a.use_base) ? 1 : 0;
      const uint8_t* base = a.tiles[net] + a.off_step_tiles;
      stream(base, mt0, mt0 + MTo, a.KC0, false, kPTile);
      base += lin0;
      for (int b = 0; b < 2 * a.nb; ++b, base += linh) stream(base, mt0, mt0 + MTo, a.KCH, true, kPTile);
      stream(base, 0, 1, a.KCH, true, out_bytes);
    }
#ifdef DPPO_CHAIN_PROF
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 0] = p_wait, a.prof[blockIdx.x * 16 + 1] = clock64() - p_t0;
#endif
  } else if (warp == 1 && rank == 0) {
    // ======================================================================================= MMA issuer (leader CTA)
    const uint32_t idesc = umma_idesc_bf16(256, kPN);
    uint32_t stage = 0, phase = 0, x0_phase = 0, xf_phase = 0;
    const uint32_t ring_lo = umma_desc_lo(smem_u32(s.ring));
    const uint32_t leader = elect_one() ? 1u : 0u;
#ifdef DPPO_CHAIN_PROF
    long long pw[4] = {0, 0, 0, 0}, tw = 0;  // waits: own ring stage, follower's ring stage, own X / x0, follower's X / x0
    const long long m_t0 = clock64();
#define DPPO_PT0() tw = clock64()
#define DPPO_PT1(i_) pw[i_] += clock64() - tw
#else
#define DPPO_PT0()
#define DPPO_PT1(i_)
#endif
    auto stage_ready = [&]() {
      DPPO_PT0();
      mbar_wait(&s.full[stage], phase);
      DPPO_PT1(0);
      DPPO_PT0();
      mbar_wait(&s.pfull[stage], phase);
      DPPO_PT1(1);
      tc_fence_after();
    };
    auto run_layer = [&](const uint8_t* b_hi, const uint8_t* b_lo, int MTl, int KCl, uint32_t d_col, bool acc, bool tiled) {
      if (!tiled) {
        DPPO_PT0();
        mbar_wait(s.x0_full, x0_phase);
        DPPO_PT1(2);
        DPPO_PT0();
        mbar_wait(s.px0_full, x0_phase);
        DPPO_PT1(3);
        x0_phase ^= 1;
        tc_fence_after();
      }
      uint32_t waited = 0;
      const uint32_t bh = umma_desc_lo(smem_u32(b_hi)), bl = umma_desc_lo(smem_u32(b_lo));
      for (int i = 0; i < KCl; ++i) {
        const int kc = tiled ? pair_chunk(i, MTo) : i;
        if (tiled) {
          const uint32_t t = uint32_t(kc) >> 1;
          if (!((waited >> t) & 1u)) {
            DPPO_PT0();
            mbar_wait(&s.x_full[t], (xf_phase >> t) & 1u);
            DPPO_PT1(2);
            DPPO_PT0();
            mbar_wait(&s.px_full[t], (xf_phase >> t) & 1u);
            DPPO_PT1(3);
            tc_fence_after();
            waited |= 1u << t;
          }
        }
        const uint32_t boff = uint32_t(kc) * (kPChunk / 16);
        const uint32_t first_acc = (acc || i > 0) ? 1u : 0u;
        for (int mt = 0; mt < MTl; ++mt) {
          const uint32_t d = tmem + d_col + uint32_t(mt) * kPN;
          stage_ready();
          {
            const uint32_t wa = ring_lo + stage * (kPTile / 16);
            const uint32_t rel = smem_u32(&s.empty[stage]);
            umma2_bf16_lo_x4_p(d, wa, bh + boff, idesc, first_acc, leader, split ? 0u : rel);
            if (split) umma2_bf16_lo_x4_p(d, wa, bl + boff, idesc, 1u, leader, rel);
          }
          if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
          if (split) {
            stage_ready();
            umma2_bf16_lo_x4_p(d, ring_lo + stage * (kPTile / 16), bh + boff, idesc, 1u, leader, smem_u32(&s.empty[stage]));
            if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
          }
        }
      }
      xf_phase ^= waited;
      umma_commit2_p(s.layer_done, leader);
    };
    for (int step = a.first_step; step < a.S; ++step) {
      run_layer(s.x0_hi, s.x0_lo, MTo, a.KC0, col_h, false, false);
      for (int b = 0; b < a.nb; ++b) {
        run_layer(s.x_hi, s.x_lo, MTo, a.KCH, col_y, false, true);
        run_layer(s.x_hi, s.x_lo, MTo, a.KCH, col_h, true, true);
      }
      run_layer(s.x_hi, s.x_lo, 1, a.KCH, col_y, false, true);
    }
#ifdef DPPO_CHAIN_PROF
    if (a.prof && lane == 0) {
      a.prof[blockIdx.x * 16 + 2] = pw[2] + pw[3], a.prof[blockIdx.x * 16 + 3] = pw[0] + pw[1];
      a.prof[blockIdx.x * 16 + 4] = clock64() - m_t0;
      for (int i = 0; i < 4; ++i) a.prof[(4096 + blockIdx.x) * 16 + i] = pw[i];
    }
#endif
  } else if (warp == 1) {
    // ======================================================================================= relay (follower CTA)
    // Walks the leader's wait sequence and forwards what only this CTA can observe: a ring stage has landed, a block the
    // leader pushed has landed in this CTA's copy of X.
    uint32_t stage = 0, phase = 0, xf_phase = 0;
    [[maybe_unused]] long long r_wait = 0, r_t0 = clock64();
    auto relay_layer = [&](int MTl, int KCl, bool tiled) {
      uint32_t waited = 0;
      for (int i = 0; i < KCl; ++i) {
        if (tiled) {
          const uint32_t t = uint32_t(pair_chunk(i, MTo)) >> 1;
          if (int(t) < MTo && !((waited >> t) & 1u)) {  // the leader's tiles; this CTA's own are announced by its epilogue
            mbar_wait(&s.x_full[t], (xf_phase >> t) & 1u);
            if (lane == 0) mbar_arrive_remote_nodata(&s.px_full[t], 0);
            waited |= 1u << t;
          }
        }
        for (int k = 0; k < MTl * a.nsplit; ++k) {
#ifdef DPPO_CHAIN_PROF
          const long long tw = clock64();
#endif
          mbar_wait(&s.full[stage], phase);
#ifdef DPPO_CHAIN_PROF
          r_wait += clock64() - tw;
#endif
          if (lane == 0) mbar_arrive_remote_nodata(&s.pfull[stage], 0);
          if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
        }
      }
      xf_phase ^= waited;
    };
    for (int step = a.first_step; step < a.S; ++step) {
      relay_layer(MTo, a.KC0, false);
      for (int b = 0; b < 2 * a.nb; ++b) relay_layer(MTo, a.KCH, true);
      relay_layer(1, a.KCH, true);
    }
#ifdef DPPO_CHAIN_PROF
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 3] = r_wait, a.prof[blockIdx.x * 16 + 4] = clock64() - r_t0;
#endif
  } else {
    // ======================================================================================= epilogue warps
    const int et = threadIdx.x - 64;            // 0..255
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const uint32_t cgrp = uint32_t(warp - 2) >> 2;  // column group of this warp: kPCols environments
    const uint32_t half = cgrp * kPCols / kPL;      // environment half (= destination CTA) of those columns
    const uint32_t row0 = cgrp * kPCols % kPL;      // first environment row inside that half
    const bool mine = half == rank;             // the columns are this CTA's own environments
    const int fl = q * 32 + lane;               // feature (TMEM lane) within an M tile / action element index
    const uint32_t col0 = cgrp * kPCols;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t odd = lane & 1;
    uint32_t ld_phase = 0;
    float xreg[kPX], zreg[kPX];
    const int nxl = kPL * a.D;  // sample elements of this CTA's environments

    auto wait_layer = [&]() {
      mbar_wait(s.layer_done, ld_phase);
      ld_phase ^= 1;
      tc_fence_after();
    };
    // this CTA's barriers of the peer's tiles expect the blocks the peer will push (the bytes may land first)
    auto expect_peer_tiles = [&]() {
      if (et == 0)
        for (int t = int(peer) * MTo; t < int(peer) * MTo + MTo; ++t)
          mbar_arrive_expect_tx(&s.x_full[t], 2u * kPChunk * uint32_t(a.nsplit));
    };
    const uint32_t tile_bytes = 2u * kPChunk;                 // one M tile = two K chunks of a 32-row operand
    const uint32_t blk_off = uint32_t(mt0) * tile_bytes;      // this CTA's feature block inside a copy of X
    // tile `mt` of this CTA's features is complete (own environments in X, the peer's in the staging block)
    auto publish_tile = [&](int mt) {
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, kPEpiThreads);
      if (et == 0) {
        const uint32_t so = uint32_t(mt) * tile_bytes, xo = blk_off + so;
        bulk_s2peer(s.x_hi + xo, s.stg_hi + so, tile_bytes, &s.x_full[mt0 + mt], peer);
        if (split) bulk_s2peer(s.x_lo + xo, s.stg_lo + so, tile_bytes, &s.x_full[mt0 + mt], peer);
        mbar_arrive(&s.x_full[mt0 + mt]);
        if (rank != 0) mbar_arrive_remote_nodata(&s.px_full[mt0 + mt], 0);
      }
    };
    auto signal_x0 = [&]() {
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, kPEpiThreads);
      if (et == 0) {
        mbar_arrive(s.x0_full);
        if (rank != 0) mbar_arrive_remote_nodata(s.px0_full, 0);
      }
    };
    float pre_b[4];
    // hidden-layer epilogue: v = acc + bias; activation; bf16 hi / lo -> X (own environments) or staging (the peer's)
    auto epi_hidden = [&](uint32_t region, const float* bias_a, const float* bias_b, bool identity) {
      uint8_t* dst_hi = mine ? s.x_hi + blk_off : s.stg_hi;
      uint8_t* dst_lo = mine ? s.x_lo + blk_off : s.stg_lo;
      for (int mt = 0; mt < MTo; ++mt) {
        float v[kPCols];
        tmem_ld(tmem + lane_addr + region + uint32_t(mt) * kPN + col0, v);
        const int f = (mt0 + mt) * 128 + fl;
        const float b = mt < 4 ? pre_b[mt & 3] : bias_a[f] + (bias_b ? bias_b[f] : 0.f);
        const uint32_t kp = uint32_t(mt * 128 + fl) & ~1u;  // first feature of the pair this thread stores, block-relative
        const uint32_t j16 = (kp & 63u) >> 3;
        const uint32_t base = (kp >> 6) * kPChunk + ((kp & 7u) << 1) + (row0 + odd) * 128u;  // row0 is a multiple of 8
#pragma unroll
        for (int c = 0; c < kPCols; c += 2) {
          float x0 = v[c] + b, x1 = v[c + 1] + b;
          if (!identity) x0 = activate<ACT>(x0), x1 = activate<ACT>(x1);
          const float recv = __shfl_xor_sync(0xffffffffu, odd ? x0 : x1, 1);
          const float fa = odd ? recv : x0, fb = odd ? x1 : recv;  // features kp, kp + 1 of environment row c + odd
          const uint32_t off = base + uint32_t(c) * 128u + ((j16 ^ ((uint32_t(c) + odd) & 7u)) << 4);
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(fa, fb);
          *reinterpret_cast<__nv_bfloat162*>(dst_hi + off) = h2;
          if (split)
            *reinterpret_cast<__nv_bfloat162*>(dst_lo + off) =
                __floats2bfloat162_rn(fa - __low2float(h2), fb - __high2float(h2));
        }
        publish_tile(mt);
      }
    };

    // ------------------------------------------------------------------------------------ prologue
    {
      const uint32_t x0_bytes = uint32_t(a.KC0) * kPChunk;
      for (uint32_t i = et * 16; i < x0_bytes; i += kPEpiThreads * 16) {
        *reinterpret_cast<uint4*>(s.x0_hi + i) = make_uint4(0, 0, 0, 0);
        if (split) *reinterpret_cast<uint4*>(s.x0_lo + i) = make_uint4(0, 0, 0, 0);
      }
      named_bar_sync(1, kPEpiThreads);
      for (int i = et; i < kPL * a.Dc; i += kPEpiThreads) {
        const int e = i / a.Dc, k = i % a.Dc;
        const int env = env0 + e;
        const float v = env < a.E ? a.state[size_t(env) * a.Dc_in + k] : 0.f;
        store_operand<kPL>(s.x0_hi, s.x0_lo, e, a.D + k, v, split);
      }
#pragma unroll 1
      for (int j = 0; j < kPX; ++j) {
        const int i = et + j * kPEpiThreads;
        float x = 0.f;
        if (i < nxl) {
          const int e = i / a.D, f = i - e * a.D;
          const int env = env0 + e;
          if (env < a.E) {
            if (a.eval_mode)
              x = a.chains_in[(size_t(env) * (a.ft + 1)) * a.D + f];
            else if (a.noise)
              x = a.noise[size_t(env) * a.D + f];
            else
              x = philox_normal(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + f, 0u);
            if (!a.eval_mode && a.chain && a.ft == a.S) a.chain[(size_t(env) * (a.ft + 1)) * a.D + f] = x;
          }
          store_operand<kPL>(s.x0_hi, s.x0_lo, e, f, x, split);
        }
        xreg[j] = x;
      }
      signal_x0();
    }

    // ------------------------------------------------------------------------------------ step loop
    for (int step = a.first_step; step < a.S; ++step) {
      const StepRow row = a.rows[step];
      const int net = (row.ft && !a.use_base) ? 1 : 0;
      const float* side = a.side[net];
      const float* tb = side + a.off_tb + size_t(row.t) * a.H;
      const int n_hidden = 1 + 2 * a.nb;
      // L = 0: h = W0 [x | obs] + TB[t];  L odd: y = W1 act(h) + b1;  L even: h += W2 act(y) + b2 (biases of h are added
      // here, never stored back into TMEM: TB[t] + the prefix sum of the b2 so far, see pack.cu)
      for (int L = 0; L < n_hidden; ++L) {
        const int b = L > 0 ? (L - 1) >> 1 : 0;
        const float* blk = side + a.off_blk + size_t(b) * a.blk_stride;
        uint32_t region = col_h;
        const float *ba = tb, *bb = nullptr;
        bool identity = false;
        if (L & 1) {
          region = col_y, ba = blk;
        } else if (L > 0) {
          ba = blk + a.H, bb = tb;
          identity = b + 1 >= a.nb;  // no activation between the last block and the output layer
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = (mt0 + (i < MTo ? i : 0)) * 128 + fl;
          pre_b[i] = ba[f] + (bb ? bb[f] : 0.f);  // fetched before the wait: L2 latency off the critical path
        }
        wait_layer();
        expect_peer_tiles();
        epi_hidden(region, ba, bb, identity);
      }

      // output layer + posterior on this CTA's own 32 environments
      const float bo = fl < a.D ? side[a.off_bout + fl] : 0.f;
      {
#pragma unroll 1
        for (int j = 0; j < kPX; ++j) {
          const int i = et + j * kPEpiThreads;
          if (i >= nxl) break;
          const int e = i / a.D, f = i - e * a.D;
          const int env = env0 + e;
          float z = 0.f;
          if (env < a.E && a.eval_mode) {
            z = a.chains_in[(size_t(env) * (a.ft + 1) + (step - a.first_step) + 1) * a.D + f];  // the stored next sample
          } else if (env < a.E) {
            if (a.noise)
              z = a.noise[(size_t(step + 1) * a.E + env) * a.D + f];
            else
              z = philox_normal(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + f, uint32_t(step + 1));
            z = fminf(fmaxf(z, -a.randn_clip), a.randn_clip);
          }
          zreg[j] = z;
        }
      }
      wait_layer();
      {
        // eps as [environment][feature] fp32 in the staging block (every push out of it has landed: the output layer's
        // MMAs, which are complete, read the blocks it fed)
        float* s_eps = reinterpret_cast<float*>(s.stg_hi);
        if (mine && q * 32 < a.D) {
          float v[kPCols];
          tmem_ld(tmem + lane_addr + col_y + col0, v);
          if (fl < a.D) {
#pragma unroll
            for (int c = 0; c < kPCols; ++c) s_eps[(int(row0) + c) * a.D + fl] = v[c] + bo;
          }
        }
        tc_fence_before();
        named_bar_sync(1, kPEpiThreads);
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
        for (int j = 0; j < kPX; ++j) {
          const int i = et + j * kPEpiThreads;
          if (i >= nxl) break;
          const int e = i / a.D, f = i - e * a.D;
          const int env = env0 + e;
          float eps = s_eps[i];
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
          store_operand<kPL>(s.x0_hi, s.x0_lo, e, f, xn, split);
        }
      }
      signal_x0();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves (or frees the pair's TMEM) while its peer can still signal it or copy into it
  tc_fence_after();
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

template <int ACT>
int launch_pair_t(const ChainArgs& a, size_t smem_bytes, cudaStream_t st) {
  auto kfn = chain_pair_kernel<ACT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(chain_pair_kernel)");
    configured = true;
  }
  const int tiles = (a.E + kPN - 1) / kPN;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(tiles * 2)), cfg.blockDim = dim3(kPThreads), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kfn, a);
  if (e != cudaSuccess) return cuda_fail(e, "chain_pair_kernel launch");
  return DPPO_OK;
}

}  // namespace

bool chain_pair_applicable(const MlpGeom& g, int NE, int C, bool forced) {
  static int env_pair = -1;  // DPPO_B200_PAIR=1: use this kernel wherever it applies (default: only when forced)
  if (env_pair < 0) {
    const char* e = getenv("DPPO_B200_PAIR");
    env_pair = e ? atoi(e) : 0;
  }
  if (!(env_pair || forced) || NE != kPN || C != 2 || g.ln || g.CH || g.MT < 2 || g.MT > 8 || (g.MT & 1)) return false;
  if (g.D > 128 || kPL * g.D > kPX * kPEpiThreads) return false;
  if (size_t(kPL) * g.D * 4 > size_t(g.MT / 2) * 2 * kPChunk * g.nsplit) return false;  // eps tile aliases the staging block
  return pair_fixed_bytes(g) + 4 * kPTile <= 232448;
}

int launch_chain_pair(ChainArgs a, const MlpGeom& g, cudaStream_t st) {
  const size_t fixed = pair_fixed_bytes(g);
  int nstage = int((232448 - fixed) / kPTile);
  if (nstage > kPMaxStages) nstage = kPMaxStages;
  static int env_stages = -1;
  if (env_stages < 0) {
    const char* e = getenv("DPPO_B200_STAGES");
    env_stages = e ? atoi(e) : 0;
  }
  if (env_stages >= 2 && env_stages < nstage) nstage = env_stages;
  a.nstage = nstage;
  a.C = 2;
  const size_t smem_bytes = fixed + size_t(nstage) * kPTile;
  return g.act == DPPO_ACT_RELU ? launch_pair_t<DPPO_ACT_RELU>(a, smem_bytes, st) : launch_pair_t<DPPO_ACT_MISH>(a, smem_bytes, st);
}

}  // namespace dppo
