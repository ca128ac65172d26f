// Persistent denoise-chain kernel for the DiffusionMLP denoiser (sm_100a: tcgen05.mma + TMEM + bulk TMA copies).
//
// Replaces, for one tile of NE environments per CTA and ALL S denoising steps in one launch:
//   VPGDiffusion.forward / p_mean_var          reference dppo/model/diffusion/diffusion_vpg.py:139-315
//   DiffusionMLP.forward + ResidualMLP         reference dppo/model/diffusion/mlp_diffusion.py:218-250, common/mlp.py:84-154
//   VPGDiffusion.get_logprobs (EVAL mode)      reference dppo/model/diffusion/diffusion_vpg.py:319-396
//
// Formulation ("swap-AB"): every Linear is D[f, e] = sum_k W[f, k] * X[e, k] with the WEIGHTS as the MMA A operand
// (M = 128 output features per tile, streamed from L2 through a ring of 16 KiB pre-swizzled tiles by cp.async.bulk)
// and the ACTIVATIONS of the NE environments as the B operand (N = NE, resident in shared memory as bf16 hi [+ lo]).
// Accumulators live in TMEM: region h (residual stream, MT*NE columns) and region y (block hidden / output layer).
// The residual add is free: the second Linear of a block accumulates straight onto h in TMEM.
//
// Feature-split clusters: the kernel is bound by how fast ONE SM can pull weight tiles out of L2 (~35 B/cycle/SM measured,
// profiles/earlier/r1a_microbench.txt), and a CTA that owns all features of its environments must pull the whole network every
// step.  A cluster of C CTAs therefore shares one tile of NE environments: CTA r computes the M-tiles
// [r*MT/C, (r+1)*MT/C) of every hidden layer (streaming only 1/C of the weights), writes its activations into its own
// copy of X and pushes that column block into the peers' copies with one bulk shared::cta -> shared::cluster copy per
// peer, operand half and M tile (the copy completes on the peer's x_full[tile] mbarrier).  The tiny output layer, the
// posterior step and the layer-0 operand are computed redundantly by every CTA, so nothing else crosses CTAs.
//
// Layer pipelining: X is handed over M tile by M tile (x_full[t]); layers that read X visit the K chunks
// own-tiles-first and chunk-major, and consecutive layers alternate between the two TMEM regions, so the MMAs of layer
// l+1 start while the epilogue of layer l is still writing its later tiles and the peers' blocks are in flight.
//
// Warp roles (320 threads): warp 0 = weight-tile producer, warp 1 = MMA issuer (+ TMEM allocator),
// warps 2..9 = epilogue (TMEM -> registers -> bias / LayerNorm / activation -> bf16 split -> shared memory, and the
// posterior-mean / noise-injection / chain-store step after the output layer).
//
// Precision: SPLIT3 issues x_hi*w_hi + x_lo*w_hi + x_hi*w_lo (three bf16 MMAs, fp32 accumulate), BF16 issues one.
#include <stdlib.h>

#include "chain_mlp.cuh"

namespace dppo {

constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr uint32_t kTile = 16384;
constexpr uint32_t kPeerBar = 8;  // x_full[] slot (after the 8 tile slots) shared by every peer tile when each CTA owns a single M tile
constexpr int kMaxStages = 12;

struct Smem {
  uint8_t *x_hi, *x_lo, *x0_hi, *x0_lo, *ring;
  uint64_t *full, *empty, *layer_done, *x_full, *x0_full, *can_send, *ln_bar, *tile_done, *early_ok;
  uint32_t* tmem_slot;
  float* ln_part;  // [kEpiWarps][2][32] per-warp partial sums
  float* ln_x;     // [8 ranks][64 envs][2] per-CTA partial sums of a cluster (LayerNorm over features split across CTAs)
};

template <int NE>
__device__ __forceinline__ Smem carve(uint8_t* base, const ChainArgs& a) {
  Smem s;
  const uint32_t xb = uint32_t(NE) * a.H * 2, x0b = uint32_t(a.KC0) * NE * 128;
  uint8_t* p = base;
  s.x_hi = p, p += xb;
  s.x_lo = p, p += (a.nsplit == 2 ? xb : 0);
  s.x0_hi = p, p += x0b;
  s.x0_lo = p, p += (a.nsplit == 2 ? x0b : 0);
  s.ring = p, p += size_t(a.nstage) * kTile;
  s.full = reinterpret_cast<uint64_t*>(p), p += 8 * kMaxStages;
  s.empty = reinterpret_cast<uint64_t*>(p), p += 8 * kMaxStages;
  s.layer_done = reinterpret_cast<uint64_t*>(p), p += 8;
  s.x_full = reinterpret_cast<uint64_t*>(p), p += 72;   // one barrier per M tile (= K-chunk pair) of X, MT <= 8, + kPeerBar
  s.x0_full = reinterpret_cast<uint64_t*>(p), p += 8;   // layer-0 / cond_mlp operands (written as a whole)
  s.can_send = reinterpret_cast<uint64_t*>(p), p += 16;  // two barriers, used alternately (see wait_layer)
  s.ln_bar = reinterpret_cast<uint64_t*>(p), p += 8;
  s.tile_done = reinterpret_cast<uint64_t*>(p), p += 8;  // early order: the first output tile of a block layer is complete
  s.early_ok = reinterpret_cast<uint64_t*>(p), p += 8;   // early order: the peers hold the blocks this CTA pushed last
  s.tmem_slot = reinterpret_cast<uint32_t*>(p), p += 16;
  s.ln_part = reinterpret_cast<float*>(p), p += a.ln ? kEpiWarps * 2 * 32 * 4 : 0;  // LayerNorm scratch only when used:
  s.ln_x = reinterpret_cast<float*>(p);                                             // 6 KiB decide about a ring stage
  return s;
}

static size_t smem_fixed_bytes(const MlpGeom& g, int NE) {
  const size_t xb = size_t(NE) * g.H * 2 * g.nsplit, x0b = size_t(g.KC0) * NE * 128 * g.nsplit;
  const size_t ln_bytes = g.ln ? kEpiWarps * 2 * 32 * 4 + 8 * 64 * 2 * 4 : 0;
  return xb + x0b + 16 * kMaxStages + 192 + ln_bytes + 1024 /* alignment slack */;
}

// ============================================================================================== the kernel
template <int NE, int ACT, bool LN>
__global__ void __launch_bounds__(kThreads, 1) chain_mlp_kernel(const ChainArgs a) {
  constexpr int CPT = NE / 2;  // accumulator columns (environments) per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Smem s = carve<NE>(smem, a);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = a.nsplit == 2;
  const int C = a.C;
  const uint32_t rank = C > 1 ? cluster_ctarank() : 0u;
  const int env0 = (blockIdx.x / C) * NE;
  const int MTo = a.MT / C;          // M-tiles of every hidden layer this CTA computes
  const int mt0 = int(rank) * MTo;   // first one

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nstage; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(s.layer_done, 1);
    for (int i = 0; i < 9; ++i) mbar_init(&s.x_full[i], 1);
    mbar_init(s.x0_full, 1);
    mbar_init(&s.can_send[0], C > 1 ? C - 1 : 1);
    mbar_init(&s.can_send[1], C > 1 ? C - 1 : 1);
    mbar_init(s.ln_bar, 1);  // one expect-tx arrival per use; the partial sums arrive as async stores (8 bytes each)
    mbar_init(s.tile_done, 1);
    mbar_init(s.early_ok, C > 1 ? C - 1 : 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (C > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them
  const uint32_t tmem = *s.tmem_slot;
  const uint32_t col_h = 0, col_y = uint32_t(MTo) * NE;

  if (warp == 0) {
    // ======================================================================================= weight-tile producer
    // (whole warp runs the loop so control flow stays uniform; one elected lane issues the copies)
    uint32_t stage = 0, phase = 0;
    int cur_net = -1;
    long long p_wait = 0;
    const long long p_t0 = clock64();
    // tiles of M-tiles [m_begin, m_end) of one Linear whose tile group starts at `base` (layout: m-tile major, k-chunk minor,
    // hi tile then lo tile)
    // `rot`: the K chunks are visited starting at chunk `rot` (wrapping): layers that read X start with the chunks this
    // CTA produced itself, which are ready first (see the MMA warp)
    // `bytes`: leading part of every 16 KiB tile that is fetched (a tile is 128 rows of 128 bytes, row-major: its first R
    // rows are its first R * 128 bytes).  The output layer has only D <= 128 real rows; the rows behind them keep whatever
    // the ring slot held before (finite or not - they only reach accumulator lanes >= D, which nobody reads).
    const uint32_t p_leader = elect_one() ? 1u : 0u;  // same issue discipline as the MMA warp: no divergent region per tile
    // `early`: the order of run_layer's early mode - phase 0 = the first two chunks for both output tiles, phase 1 / 2 =
    // the remaining chunks for the first / second output tile
    auto stream = [&](const uint8_t* base, int m_begin, int m_end, int KCl, int rot, uint32_t bytes, bool early) {
      // chunk-major: every output tile consumes K chunk kc before anybody touches the next chunk, so one arrived tile of
      // X feeds (m_end - m_begin) x 2 tile-MMAs before the next one is needed
      for (int ph = 0; ph < (early ? 3 : 1); ++ph)
      for (int j = (early && ph > 0) ? 2 : 0; j < ((early && ph == 0) ? 2 : KCl); ++j) {
        int kc = j + rot;
        if (kc >= KCl) kc -= KCl;
        for (int mt = (early && ph == 2) ? m_begin + 1 : m_begin; mt < ((early && ph == 1) ? m_begin + 1 : m_end); ++mt) {
          const uint8_t* src = base + size_t(mt) * KCl * a.nsplit * kTile;
          for (int h = 0; h < a.nsplit; ++h) {
#ifdef DPPO_CHAIN_PROF
            const long long tw = clock64();
#endif
            mbar_wait(&s.empty[stage], phase ^ 1);
#ifdef DPPO_CHAIN_PROF
            p_wait += clock64() - tw;
#endif
            bulk_g2s_expect_p(s.ring + size_t(stage) * kTile, src + size_t(kc * a.nsplit + h) * kTile, bytes,
                              &s.full[stage], p_leader);
            if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
          }
        }
      }
    };
    const int rot = mt0 * 2;
    const uint32_t out_bytes = uint32_t((a.D + 7) / 8) * 8u * 128u;  // D <= 128
    const size_t lin0 = size_t(a.MT) * a.KC0 * a.nsplit * kTile, linh = size_t(a.MT) * a.KCH * a.nsplit * kTile;
    for (int step = a.first_step; step < a.S; ++step) {
      const int net = (a.rows[step].ft && !a.use_base) ? 1 : 0;
      if (a.CH && net != cur_net) {
        stream(a.tiles[net], 0, a.MTc, a.KCc, 0, kTile, false);
        stream(a.tiles[net] + size_t(a.MTc) * a.KCc * a.nsplit * kTile, 0, 1, a.CH / 64, 0, kTile, false);
      }
      cur_net = net;
      const uint8_t* base = a.tiles[net] + a.off_step_tiles;
      stream(base, mt0, mt0 + MTo, a.KC0, 0, kTile, false);
      base += lin0;
      for (int b = 0; b < 2 * a.nb; ++b, base += linh) stream(base, mt0, mt0 + MTo, a.KCH, rot, kTile, a.early != 0);
      stream(base, 0, 1, a.KCH, rot, out_bytes, false);
    }
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 0] = p_wait, a.prof[blockIdx.x * 16 + 1] = clock64() - p_t0;
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    // Whole warp runs the loop (uniform control flow => descriptor arithmetic stays in uniform registers); one elected
    // lane issues tcgen05.mma / tcgen05.commit.
    const uint32_t idesc = umma_idesc_bf16(128, NE);
    // Operand hand-off.  Layers that read X (block layers, output layer) wait per M tile of the previous layer's output
    // (x_full[t] = chunk pair t of X is complete: written by this CTA's epilogue or pushed by its owner) and visit the K
    // chunks starting with this CTA's own tiles, so the next layer's MMAs begin while the previous layer's epilogue is
    // still writing its later tiles and the peers' blocks are still in flight (the accumulator regions alternate between
    // consecutive layers, so those MMAs never touch what the epilogue is reading).  Layer 0 / cond_mlp read operands
    // that are written as a whole: one barrier (x0_full).
    uint32_t stage = 0, phase = 0, x0_phase = 0, xf_phase = 0;
    const int rot = mt0 * 2;
    long long m_wait_x = 0, m_wait_full = 0;
    [[maybe_unused]] long long m_wx[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // DPPO_CHAIN_PROF: x waits by (layer kind, tile class)
    [[maybe_unused]] int kind = 0;                                                // 0 layer 0 / cond, 1 l1, 2 l2, 3 output layer
    const long long m_t0 = clock64();
    const uint32_t ring_lo = umma_desc_lo(smem_u32(s.ring));
    // The issue loop is ONE warp's dependent instruction stream on a scheduler it shares with two epilogue warps, and it
    // is on the kernel's critical path (measured: a handful of extra instructions per tile cost 7 % of the launch).  So:
    // no branch around the MMAs (predicated on the lane elected once, here), and the wait counters of the role profile
    // are compiled in only with -DDPPO_CHAIN_PROF (scripts/chain_prof.py needs a build with DPPO_B200_CHAIN_PROF=1).
    const uint32_t leader = elect_one() ? 1u : 0u;
#ifdef DPPO_CHAIN_PROF
#define DPPO_MMA_T0() tw = clock64()
#define DPPO_MMA_T1(acc_) acc_ += clock64() - tw
#define DPPO_MMA_TX(cls_) m_wx[kind * 3 + (cls_)] += clock64() - tw
#else
#define DPPO_MMA_TX(cls_)
#define DPPO_MMA_T0()
#define DPPO_MMA_T1(acc_)
#endif
    // `early` (block layers of a CTA that owns two M tiles): phase 0 issues the two K chunks of this CTA's own first tile
    // of X for both output tiles, phase 1 the remaining chunks for output tile 0 (commit -> tile_done), phase 2 the
    // remaining chunks for output tile 1 (commit -> layer_done).  After tile_done nobody in this CTA reads the first own
    // tile of X any more, so the epilogue of output tile 0 (which overwrites exactly that tile) runs under phase 2.
    auto run_layer = [&](const uint8_t* b_hi, const uint8_t* b_lo, int MTl, int KCl, uint32_t d_col, bool acc, bool tiled,
                         bool early) {
      [[maybe_unused]] long long tw = 0;
      DPPO_MMA_T0();
      if (!tiled) {
        mbar_wait(s.x0_full, x0_phase);
        x0_phase ^= 1;
        tc_fence_after();
        DPPO_MMA_TX(0);
      }
      DPPO_MMA_T1(m_wait_x);
      uint32_t waited = 0;
      const uint32_t bh = umma_desc_lo(smem_u32(b_hi)), bl = umma_desc_lo(smem_u32(b_lo));
      for (int ph = 0; ph < (early ? 3 : 1); ++ph) {
      for (int j = (early && ph > 0) ? 2 : 0; j < ((early && ph == 0) ? 2 : KCl); ++j) {
        int kc = j;
        if (tiled) {
          kc = j + rot;
          if (kc >= KCl) kc -= KCl;
          uint32_t t = uint32_t(kc) >> 1;
          if (MTo == 1 && int(t) != mt0) t = kPeerBar;  // one tile per CTA: all the peers' tiles share one barrier
          if (!((waited >> t) & 1u)) {
            DPPO_MMA_T0();
            mbar_wait(&s.x_full[t], (xf_phase >> t) & 1u);
            DPPO_MMA_T1(m_wait_x);
            DPPO_MMA_TX(int(t) == mt0 ? 0 : ((int(t) > mt0 && int(t) < mt0 + MTo) ? 1 : 2));
            tc_fence_after();
            waited |= 1u << t;
          }
        }
        const uint32_t boff = uint32_t(kc) * (NE * 128 / 16);
        const uint32_t first_acc = (acc || j > 0) ? 1u : 0u;
        for (int mt = (early && ph == 2) ? 1 : 0; mt < ((early && ph == 1) ? 1 : MTl); ++mt) {
          const uint32_t d = tmem + d_col + uint32_t(mt) * NE;
          DPPO_MMA_T0();
          mbar_wait(&s.full[stage], phase);
          DPPO_MMA_T1(m_wait_full);
          tc_fence_after();
          {
            const uint32_t wa = ring_lo + stage * (kTile / 16);
            const uint32_t rel = smem_u32(&s.empty[stage]);
            umma_bf16_lo_x4_p(d, wa, bh + boff, idesc, first_acc, leader, split ? 0u : rel);
            if (split) umma_bf16_lo_x4_p(d, wa, bl + boff, idesc, 1u, leader, rel);
          }
          if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
          if (split) {
            DPPO_MMA_T0();
            mbar_wait(&s.full[stage], phase);
            DPPO_MMA_T1(m_wait_full);
            tc_fence_after();
            umma_bf16_lo_x4_p(d, ring_lo + stage * (kTile / 16), bh + boff, idesc, 1u, leader, smem_u32(&s.empty[stage]));
            if (++stage == uint32_t(a.nstage)) stage = 0, phase ^= 1;
          }
        }
      }
      if (early && ph == 1) umma_commit_p(s.tile_done, leader);
      }
      xf_phase ^= waited;
      umma_commit_p(s.layer_done, leader);
    };
    int cur_net = -1;
    for (int step = a.first_step; step < a.S; ++step) {
      const int net = (a.rows[step].ft && !a.use_base) ? 1 : 0;
      if (a.CH && net != cur_net) {
        run_layer(s.x_hi, s.x_lo, a.MTc, a.KCc, col_y, false, false, false);
        run_layer(s.x_hi, s.x_lo, 1, a.CH / 64, col_y, false, false, false);
      }
      cur_net = net;
      kind = 0;
      run_layer(s.x0_hi, s.x0_lo, MTo, a.KC0, col_h, false, false, false);
      for (int b = 0; b < a.nb; ++b) {
        kind = 1;
        run_layer(s.x_hi, s.x_lo, MTo, a.KCH, col_y, false, true, a.early != 0);
        kind = 2;
        run_layer(s.x_hi, s.x_lo, MTo, a.KCH, col_h, true, true, a.early != 0);
      }
      kind = 3;
      run_layer(s.x_hi, s.x_lo, 1, a.KCH, col_y, false, true, false);
    }
    if (a.prof && lane == 0) {
      a.prof[blockIdx.x * 16 + 2] = m_wait_x, a.prof[blockIdx.x * 16 + 3] = m_wait_full;
      a.prof[blockIdx.x * 16 + 4] = clock64() - m_t0;
#ifdef DPPO_CHAIN_PROF
      for (int i = 0; i < 12; ++i) a.prof[(4096 + blockIdx.x) * 16 + i] = m_wx[i];
#endif
    }
  } else {
    // ======================================================================================= epilogue warps
    const int et = threadIdx.x - 64;     // 0..255
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which half of the NE columns
    const int fl = q * 32 + lane;        // feature (TMEM lane) within an m-tile / action element index
    const int col0 = half * CPT;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    uint32_t ld_phase = 0;
    float xreg[CPT];
    float zreg[CPT];  // the step's injected noise, drawn while the output-layer MMAs run (see the step loop)
    const int nxe = NE * a.D;  // sample elements of this tile (<= 256 * CPT because D <= 128)

    long long e_wait = 0, e_hand = 0, e_early = 0, e_ack = 0;
    const long long e_t0 = clock64();
    // Handshakes alternate between TWO barriers.  Consecutive handshakes are not always separated by an all-to-all data
    // dependency: the output layer ends with a handshake but no push, and layer 0 of the next step runs on the local x0
    // operand, so a fast peer can send its NEXT handshake arrival while a slow peer has not sent the current one.  With a
    // single barrier that early arrival completes the current phase prematurely, blocks get pushed into a copy of X
    // that is still being read and the transaction counts of x_full drift until a wait never completes (seen once per
    // ~3000 launches).  A peer can be at most one handshake ahead (the next one needs everybody's push), so two suffice.
    uint32_t cs_phase[2] = {0u, 0u}, hs = 0;
    // `handshake`: this epilogue overwrites parts of X (its column block, or the s_eps alias), so (a) the peers must have
    // consumed the block pushed earlier and (b) they must be done reading their copy before a new block is pushed
    auto wait_layer = [&](bool exchange) {
      const long long tw = clock64();
      mbar_wait(s.layer_done, ld_phase);
      e_wait += clock64() - tw;
      ld_phase ^= 1;
      tc_fence_after();
      if (exchange && C > 1) {
        // this CTA's MMAs no longer read its copy of X: the peers may overwrite their column blocks in it ...
        uint64_t* bar = &s.can_send[hs & 1u];
        if (et == 0)
          for (uint32_t p = 0; p < uint32_t(C); ++p)
            if (p != rank) mbar_arrive_remote_nodata(bar, p);
        // ... and once every peer says the same, (a) the block this CTA pushed after the previous layer has been
        // consumed, so its source may be overwritten, and (b) the new block may be pushed into the peers' copies
        const long long th = clock64();
        mbar_wait_cluster(bar, hs & 1u ? cs_phase[1] : cs_phase[0]);
        e_hand += clock64() - th;
        if (hs & 1u) cs_phase[1] ^= 1u; else cs_phase[0] ^= 1u;
        ++hs;
      }
    };
    // before an exchange layer's epilogue: this CTA's barriers of the PEERS' tiles expect the bytes their owners will push
    // (the bytes may land first: the transaction count is signed)
    auto expect_peer_tiles = [&]() {
      if (C > 1 && et == 0) {
        if (MTo == 1) {  // (C - 1) single-tile blocks land on the shared barrier: one wait instead of C - 1
          mbar_arrive_expect_tx(&s.x_full[kPeerBar], uint32_t(C - 1) * 2u * NE * 128u * uint32_t(a.nsplit));
        } else {
          for (int t = 0; t < a.MT; ++t)
            if (t < mt0 || t >= mt0 + MTo) mbar_arrive_expect_tx(&s.x_full[t], 2u * NE * 128u * uint32_t(a.nsplit));
        }
      }
    };
    // hand the operand written by this epilogue over to the MMA warp (and, for an exchange layer, to the peers)
    const uint32_t blk_off = uint32_t(mt0) * 2u * NE * 128u;  // this CTA's column block of X starts here
    // The block is published M tile by M tile: the epilogue releases tile mt as soon as its columns are written (its copy
    // into the peers overlaps the arithmetic of the later tiles), signal_x releases the last one.
    const uint32_t tile_bytes = 2u * NE * 128u;  // one M tile = two 64-feature chunks of X
    // tile `mt` of this CTA's block is complete in shared memory: copy it into the peers (completing on THEIR barrier of
    // that tile) and release it to this CTA's MMA warp
    auto push_tile = [&](int mt) {
      const uint32_t off = blk_off + uint32_t(mt) * tile_bytes;
      uint64_t* bar = &s.x_full[mt0 + mt];
      uint64_t* peer_bar = MTo == 1 ? &s.x_full[kPeerBar] : bar;  // where this tile is accounted at the receivers
      for (uint32_t p = 0; p < uint32_t(C); ++p) {
        if (p == rank) continue;
        bulk_s2peer(s.x_hi + off, s.x_hi + off, tile_bytes, peer_bar, p);
        if (split) bulk_s2peer(s.x_lo + off, s.x_lo + off, tile_bytes, peer_bar, p);
      }
    };
    auto publish_tile = [&](int mt) {
      push_tile(mt);
      mbar_arrive(&s.x_full[mt0 + mt]);
    };
    // `exchange`: the epilogue wrote (the last tile of) this CTA's block of X; otherwise a whole-operand hand-off
    auto signal_x = [&](bool exchange) {
      tmem_wait_st();
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        if (exchange) {
          publish_tile(MTo - 1);
        } else {
          mbar_arrive(s.x0_full);
        }
      }
    };
    // raw observation -> X (input of the cond_mlp), zero padded to the 64-wide chunks it occupies
    auto stage_cond_input = [&]() {
      const int kw = a.KCc * 64;
      for (int i = et; i < NE * kw; i += kEpiThreads) {
        const int e = i / kw, k = i % kw;
        const int env = env0 + e;
        const float v = (env < a.E && k < a.Dc_in) ? a.state[size_t(env) * a.Dc_in + k] : 0.f;
        store_operand<NE>(s.x_hi, s.x_lo, e, k, v, split);
      }
    };
    // generic hidden-layer epilogue: v = acc + bias_a (+ bias_b); [LayerNorm]; activation; -> X
    // `mt_first`: global index of the layer's first M-tile held in `region` (feature = (mt_first + mt) * 128 + lane).
    // The accumulator is [feature lane][env column]; X wants [env row][feature] with 2-byte elements.  Neighbouring
    // lanes swap one value per column pair (even lane keeps column c, odd lane column c + 1), so that every thread owns
    // TWO consecutive features of one env row: one packed bf16x2 convert and one 4-byte store per operand half.
    const uint32_t odd = lane & 1;
    uint32_t ln_phase = 0, td_phase = 0, eo_phase = 0;
    // Per-feature constants (bias, LayerNorm gain / shift) of the first four M tiles are fetched BEFORE the wait for the
    // layer's MMAs: they come from the L2-resident side table (the L1 is a few KB next to 220 KB of shared memory), and
    // a load issued after tcgen05.ld would put ~700 cycles of L2 latency on the critical path of every M tile.
    float pre_b[4], pre_g[4], pre_be[4];
    auto prefetch_side = [&](int mt_first, int MTl, const float* bias_a, const float* bias_b, const float* ln_g,
                             const float* ln_b) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int f = (mt_first + (i < MTl ? i : 0)) * 128 + fl;
        pre_b[i] = bias_a[f] + (bias_b ? bias_b[f] : 0.f);
        pre_g[i] = (LN && ln_g != nullptr) ? ln_g[f] : 1.f;
        pre_be[i] = (LN && ln_g != nullptr) ? ln_b[f] : 0.f;
      }
    };
    // `whole_row`: this CTA holds every feature of the layer (cond_mlp), otherwise they are split over the cluster
    // [mt_lo, mt_hi): the M tiles of the layer handled by this call (the early order handles them in two calls)
    auto epi_hidden = [&](uint32_t region, int mt_first, int MTl, const float* bias_a, const float* bias_b, bool identity,
                          const float* ln_g, const float* ln_b, bool whole_row, bool push, int mt_lo, int mt_hi) {
      float mean[LN ? CPT : 1], rstd[LN ? CPT : 1];
      if (LN && ln_g != nullptr) {
        float s1[CPT], s2[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) s1[c] = 0.f, s2[c] = 0.f;
        for (int mt = 0; mt < MTl; ++mt) {
          float v[CPT];
          tmem_ld(tmem + lane_addr + region + uint32_t(mt) * NE + col0, v);
          const int f = (mt_first + mt) * 128 + fl;
          const float b = mt < 4 ? pre_b[mt & 3] : bias_a[f] + (bias_b ? bias_b[f] : 0.f);
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const float x = v[c] + b;
            s1[c] += x, s2[c] += x * x;
          }
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
            s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o);
          }
        }
        float* mine = s.ln_part + (warp - 2) * 64;
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) mine[c] = s1[c], mine[32 + c] = s2[c];
        }
        named_bar_sync(1, kEpiThreads);
        if (whole_row || C == 1) {
          const float inv_n = 1.f / float(MTl * 128);
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float* part = s.ln_part + (half * 4 + w) * 64;
              t1 += part[c], t2 += part[32 + c];
            }
            const float m = t1 * inv_n;
            const float var = fmaxf(t2 * inv_n - m * m, 0.f);
            mean[LN ? c : 0] = m, rstd[LN ? c : 0] = rsqrtf(var + 1e-6f);
          }
        } else {
          // the row's features are split over the C CTAs of the cluster: thread e < NE publishes this CTA's partial sums of
          // env e to every CTA as ONE async 8-byte store that completes on that CTA's ln_bar (no release fence: the first
          // version paid MEMBAR.ALL.GPU per arrive); everybody then adds C partials
          if (et == 0) mbar_arrive_expect_tx(s.ln_bar, uint32_t(C) * NE * 8u);
          if (et < NE) {
            const int h_ = et / CPT, c_ = et % CPT;
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              const float* part = s.ln_part + (h_ * 4 + w) * 64;
              t1 += part[c_], t2 += part[32 + c_];
            }
            float* slot = s.ln_x + (rank * 64 + et) * 2;
            for (uint32_t p = 0; p < uint32_t(C); ++p) st_async_v2(slot, p, t1, t2, s.ln_bar);
          }
          mbar_wait(s.ln_bar, ln_phase);
          ln_phase ^= 1;
          const float inv_n = 1.f / float(MTl * 128 * C);
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            float t1 = 0.f, t2 = 0.f;
            for (int r = 0; r < C; ++r) {
              const float* slot = s.ln_x + (r * 64 + col0 + c) * 2;
              t1 += slot[0], t2 += slot[1];
            }
            const float m = t1 * inv_n;
            const float var = fmaxf(t2 * inv_n - m * m, 0.f);
            mean[LN ? c : 0] = m, rstd[LN ? c : 0] = rsqrtf(var + 1e-6f);
          }
        }
      }
      for (int mt = mt_lo; mt < mt_hi; ++mt) {
        float v[CPT];
        tmem_ld(tmem + lane_addr + region + uint32_t(mt) * NE + col0, v);
        const int f = (mt_first + mt) * 128 + fl;
        const float b = mt < 4 ? pre_b[mt & 3] : bias_a[f] + (bias_b ? bias_b[f] : 0.f);
        float g = 1.f, be = 0.f;
        if (LN && ln_g != nullptr) g = mt < 4 ? pre_g[mt & 3] : ln_g[f], be = mt < 4 ? pre_be[mt & 3] : ln_b[f];
        const uint32_t kp = uint32_t(f) & ~1u;  // first feature of the pair this thread stores
        const uint32_t j16 = (kp & 63u) >> 3;
        const uint32_t base = (kp >> 6) * (NE * 128u) + ((kp & 7u) << 1) + (uint32_t(col0) + odd) * 128u;
#pragma unroll
        for (int c = 0; c < CPT; c += 2) {
          float x0 = v[c] + b, x1 = v[c + 1] + b;
          if (LN && ln_g != nullptr) {
            x0 = (x0 - mean[LN ? c : 0]) * rstd[LN ? c : 0] * g + be;
            x1 = (x1 - mean[LN ? c + 1 : 0]) * rstd[LN ? c + 1 : 0] * g + be;
          }
          if (!identity) x0 = activate<ACT>(x0), x1 = activate<ACT>(x1);
          const float recv = __shfl_xor_sync(0xffffffffu, odd ? x0 : x1, 1);
          const float fa = odd ? recv : x0, fb = odd ? x1 : recv;  // features kp, kp + 1 of env row col0 + c + odd
          const uint32_t off = base + uint32_t(c) * 128u + ((j16 ^ ((uint32_t(c) + odd) & 7u)) << 4);
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(fa, fb);
          *reinterpret_cast<__nv_bfloat162*>(s.x_hi + off) = h2;
          if (split)
            *reinterpret_cast<__nv_bfloat162*>(s.x_lo + off) =
                __floats2bfloat162_rn(fa - __low2float(h2), fb - __high2float(h2));
        }
        if (push && mt + 1 < MTl) {
          // tile mt of this CTA's block is complete in shared memory: hand it to the MMA warp and the peers now
          tc_fence_before();
          fence_proxy_async_smem();
          named_bar_sync(1, kEpiThreads);
          if (et == 0) publish_tile(mt);
        }
      }
    };

    // ------------------------------------------------------------------------------------ prologue
    {
      // zero the layer-0 operand (padding columns must read as 0), then place the conditioning columns
      const uint32_t x0_bytes = uint32_t(a.KC0) * NE * 128;
      for (uint32_t i = et * 16; i < x0_bytes; i += kEpiThreads * 16) {
        *reinterpret_cast<uint4*>(s.x0_hi + i) = make_uint4(0, 0, 0, 0);
        if (split) *reinterpret_cast<uint4*>(s.x0_lo + i) = make_uint4(0, 0, 0, 0);
      }
      named_bar_sync(1, kEpiThreads);
      if (a.CH == 0) {
        for (int i = et; i < NE * a.Dc; i += kEpiThreads) {
          const int e = i / a.Dc, k = i % a.Dc;
          const int env = env0 + e;
          const float v = env < a.E ? a.state[size_t(env) * a.Dc_in + k] : 0.f;
          store_operand<NE>(s.x0_hi, s.x0_lo, e, a.D + k, v, split);
        }
      } else {
        stage_cond_input();
      }
      // x_T (sampling) or the first stored chain entry (evaluation).  Flat element mapping: thread `et` owns elements
      // i = et + j*256 of the tile's NE x D sample block for the whole chain (x stays in registers between steps), so
      // every global access below is contiguous across the 256 epilogue threads.  The loops over j are deliberately NOT
      // unrolled (xreg sits in L1-resident local memory): a handful of elements per thread, but a large body.
#pragma unroll 1
      for (int j = 0; j < CPT; ++j) {
        const int i = et + j * kEpiThreads;
        float x = 0.f;
        if (i < nxe) {
          const int e = i / a.D, f = i - e * a.D;
          const int env = env0 + e;
          if (env < a.E) {
            if (a.eval_mode)
              x = a.chains_in[(size_t(env) * (a.ft + 1)) * a.D + f];
            else if (a.noise)
              x = a.noise[size_t(env) * a.D + f];
            else
              x = philox_normal(a.seed, a.offset, uint64_t(a.env_offset + env) * a.D + f, 0u);
            if (!a.eval_mode && a.chain && a.ft == a.S && rank == 0) a.chain[(size_t(env) * (a.ft + 1)) * a.D + f] = x;
          }
          store_operand<NE>(s.x0_hi, s.x0_lo, e, f, x, split);
        }
        xreg[j] = x;
      }
      signal_x(false);
    }

    // ------------------------------------------------------------------------------------ step loop
    int cur_net = -1;
    for (int step = a.first_step; step < a.S; ++step) {
      const StepRow row = a.rows[step];
      const int net = (row.ft && !a.use_base) ? 1 : 0;
      const float* side = a.side[net];
      // Hidden layers through ONE instance of the epilogue code (instruction-cache footprint):
      //   L = -2  cond_mlp hidden   (only when the network changes; computed redundantly by every CTA of a cluster)
      //   L = -1  cond_mlp output   -> conditioning columns of the layer-0 operand
      //   L =  0  layer 0           h  = W0 [x | cond] + TB[t]             (TB folds bias + time embedding)
      //   L odd   l1 of block b     y  = W1 act(norm1(h)) + b1
      //   L even  l2 of block b     h += W2 act(norm2(y)) + b2             (h's biases are never stored back into TMEM:
      //                                                                      its epilogue adds TB[t] + prefix sum of b2)
      const float* tb = side + a.off_tb + size_t(row.t) * a.H;
      const int n_hidden = 1 + 2 * a.nb;
      for (int L = (a.CH && net != cur_net) ? -2 : 0; L < n_hidden; ++L) {
        if (L == -1) {
          wait_layer(false);
          float v[CPT];
          tmem_ld(tmem + lane_addr + col_y + col0, v);
          if (fl < a.CO) {
            const float b = side[a.off_bc1 + fl];
#pragma unroll
            for (int c = 0; c < CPT; ++c) store_operand<NE>(s.x0_hi, s.x0_lo, col0 + c, a.D + fl, v[c] + b, split);
          }
          signal_x(false);
          continue;
        }
        const int b = L > 0 ? (L - 1) >> 1 : 0;
        const float* blk = side + a.off_blk + size_t(b) * a.blk_stride;
        uint32_t region = col_h;
        int mt_first = mt0, MTl = MTo;
        const float *ba = tb, *bb = nullptr, *lg = nullptr, *lb = nullptr;
        bool identity = false;
        if (L == -2) {
          region = col_y, mt_first = 0, MTl = a.MTc, ba = side + a.off_bc0;
        } else if (L == 0) {
          if (a.ln) lg = blk + 2 * a.H, lb = blk + 3 * a.H;
        } else if (L & 1) {
          region = col_y, ba = blk;
          if (a.ln) lg = blk + 4 * a.H, lb = blk + 5 * a.H;
        } else {
          ba = blk + a.H, bb = tb;  // slot 1 of a block holds the PREFIX SUM of the l2 biases (pack.cu)
          if (b + 1 < a.nb) {
            if (a.ln) lg = blk + a.blk_stride + 2 * a.H, lb = blk + a.blk_stride + 3 * a.H;
          } else {
            identity = true;  // no activation between the last block and the output layer
          }
        }
        prefetch_side(mt_first, MTl, ba, bb, lg, lb);
        if (!LN && a.early && L >= 1) {
          // early order (see run_layer): output tile 0 is complete while the MMAs of output tile 1 still run.  Its
          // epilogue overwrites this CTA's first own tile of X, which (a) the local MMAs no longer read (tile_done) and
          // (b) was the source of the push after the previous layer: every peer acknowledges, at ITS tile_done, that the
          // blocks pushed to it have landed (its phase 1 has consumed them).
          const long long te0 = clock64();
          mbar_wait(s.tile_done, td_phase);
          e_early += clock64() - te0;
          td_phase ^= 1;
          tc_fence_after();
          if (C > 1) {
            if (et == 0)
              for (uint32_t p = 0; p < uint32_t(C); ++p)
                if (p != rank) mbar_arrive_remote_nodata(s.early_ok, p);
            const long long te1 = clock64();
            mbar_wait_cluster(s.early_ok, eo_phase);
            e_ack += clock64() - te1;
            eo_phase ^= 1;
          }
          epi_hidden(region, mt_first, MTl, ba, bb, identity, lg, lb, false, false, 0, 1);
          tc_fence_before();
          fence_proxy_async_smem();
          named_bar_sync(1, kEpiThreads);
          if (et == 0) mbar_arrive(&s.x_full[mt0]);  // the next layer's first chunks may be issued right behind this layer
          wait_layer(true);
          expect_peer_tiles();
          if (et == 0) push_tile(0);  // the peers are done reading their copies: now the block may land in them
          epi_hidden(region, mt_first, MTl, ba, bb, identity, lg, lb, false, false, 1, 2);
          signal_x(true);
          continue;
        }
        wait_layer(L >= 0);
        if (L >= 0) expect_peer_tiles();
        epi_hidden(region, mt_first, MTl, ba, bb, identity, lg, lb, L < 0, L >= 0, 0, MTl);
        signal_x(L >= 0);
      }
      cur_net = net;

      // output layer + posterior.  The accumulator holds eps as [feature lane][env column]; the warps whose lane
      // quarter holds valid features move it (+ bias) to a [env][feature] fp32 tile in shared memory (aliasing X, which
      // the completed output-layer MMAs no longer read), then ALL epilogue threads run the posterior on the flat mapping.
      // Every CTA of a cluster does this redundantly (same inputs, same noise); rank 0 alone writes to global memory.
      const float bo = fl < a.D ? side[a.off_bout + fl] : 0.f;  // fetched while the output-layer MMAs run
      // The posterior is a serial stretch of the chain (output-layer MMAs -> posterior -> layer 0), and most of its
      // per-element latency is the Philox + Box-Muller draw, which depends on nothing the MMAs produce: draw now, while
      // this warp would otherwise sit in wait_layer.
      {
#pragma unroll 1
        for (int j = 0; j < CPT; ++j) {
          const int i = et + j * kEpiThreads;
          if (i >= nxe) break;
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
      wait_layer(true);
      {
        float* s_eps = reinterpret_cast<float*>(s.x_hi);
        if (q * 32 < a.D) {
          float v[CPT];
          tmem_ld(tmem + lane_addr + col_y + col0, v);
          if (fl < a.D) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) s_eps[(col0 + c) * a.D + fl] = v[c] + bo;
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
          {
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
                if (rank == 0) a.logp[(size_t(env) * a.ft + d_eval) * a.D + f] = -(diff * diff) * inv_2var - log_std - 0.91893853320467274f;
              } else {
                xn = mu + stdv * zreg[j];
                if (last && a.final_clip >= 0.f) xn = fminf(fmaxf(xn, -a.final_clip), a.final_clip);
                if (rank == 0) {
                  if (a.chain && row.slot >= 0) a.chain[(size_t(env) * (a.ft + 1) + row.slot) * a.D + f] = xn;
                  if (last) {
                    a.traj[size_t(env) * a.D + f] = xn;
                    if (!(fabsf(xn) <= 3.0e38f)) atomicOr(a.nonfinite, 1);
                  }
                }
              }
            }
            xreg[j] = xn;
            store_operand<NE>(s.x0_hi, s.x0_lo, e, f, xn, split);
          }
        }
        // the next step switches network and has a cond_mlp: its input must be staged before the hand-off
        if (a.CH && step + 1 < a.S) {
          const int nnet = (a.rows[step + 1].ft && !a.use_base) ? 1 : 0;
          if (nnet != net) {
            named_bar_sync(1, kEpiThreads);  // s_eps aliases X: every thread must be done reading it
            stage_cond_input();
          }
        }
      }
      signal_x(false);
    }
    if (a.prof && et == 0) a.prof[blockIdx.x * 16 + 5] = e_wait, a.prof[blockIdx.x * 16 + 6] = clock64() - e_t0, a.prof[blockIdx.x * 16 + 7] = e_hand;
    if (a.prof && lane == 0) a.prof[blockIdx.x * 16 + 8 + (warp - 2)] = (clock64() - e_t0) - e_wait - e_hand - e_early - e_ack;
#ifdef DPPO_CHAIN_PROF
    if (a.prof && et == 0) a.prof[(4096 + blockIdx.x) * 16 + 12] = e_early, a.prof[(4096 + blockIdx.x) * 16 + 13] = e_ack;
#endif
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (C > 1) cluster_sync_all();  // no CTA leaves while a peer can still signal it or copy into it
}

// ============================================================================================== host launch
// Launch shape: NE environments per tile (= N of every MMA) and C CTAs per cluster sharing one tile (feature split).
// Cost model per CTA and denoising step, from the micro-benchmarks in profiles/ (cycles):
//   ingest   weight tiles this CTA streams x 16 KiB / 34.7 B/cycle          (L2 -> shared memory, per-SM limit)
//   mma      MMAs x max(NE/2, 32 + NE/4)                                    (tensor floor vs. shared-memory operand read)
//   epi      activation elements per epilogue thread x ~30 cycles
//   exch     bytes pushed to the peers / ~25 B/cycle (3 peers) or ~6 B/cycle (7 peers)
// and the launch needs ceil(tiles / co-resident clusters) waves.  DPPO_B200_TILE_ENVS / DPPO_B200_CLUSTER override.
struct LaunchShape {
  int NE, C;
};
static LaunchShape pick_shape(const dppo_ctx* ctx, int E) {
  const MlpGeom& g = ctx->g;
  const int cap = g.H <= 512 ? 64 : 32;  // X operand (NE x H bf16 hi + lo) must leave room for the weight ring
  static int env_ne = -1, env_c = -1;
  if (env_ne < 0) {
    const char* e = getenv("DPPO_B200_TILE_ENVS");
    env_ne = e ? atoi(e) : 0;
    e = getenv("DPPO_B200_CLUSTER");
    env_c = e ? atoi(e) : 0;
  }
  const int forced_ne = ctx->force_ne ? ctx->force_ne : env_ne, forced_c = ctx->force_c > 0 ? ctx->force_c : env_c;
  LaunchShape best{cap, 1};
  double best_t = 1e30;
  for (int ci = 0; ci < 4; ++ci) {
    const int C = 1 << ci;
    if (g.MT % C) continue;
    if (forced_c > 0 && C != forced_c) continue;
    for (int NE = 16, ni = 0; NE <= cap; NE *= 2, ++ni) {
      if (forced_ne > 0 && NE != forced_ne) continue;
      const int tiles = (E + NE - 1) / NE;
      const int max_clusters = ctx->chain_clusters[ni][ci] > 0 ? ctx->chain_clusters[ni][ci] : 1;  // co-resident clusters (occupancy query)
      const int waves = (tiles + max_clusters - 1) / max_clusters;
      const double pairs = double(g.MT / C) * (g.KC0 + 2.0 * g.nb * g.KCH) + g.KCH;  // (hi, lo) tile pairs per step
      const double ingest = pairs * g.nsplit * 16384.0 / 34.7;
      const double per_mma = NE / 2.0 > 32.0 + NE / 4.0 ? NE / 2.0 : 32.0 + NE / 4.0;
      const double mma = pairs * 4.0 * (g.nsplit == 2 ? 3.0 : 1.0) * per_mma;
      const double layers = 1.0 + 2.0 * g.nb;
      const double epi = layers * (double(NE) * g.H / C / 256.0) * 30.0 + 2000.0;
      // pushes into up to 3 peers move ~25 B/cycle, into 7 peers ~6 B/cycle (fitted to furniture at 125 / 250 envs: NE = 32,
      // C = 8 takes 267 k cycles per step where 25 B/cycle predicted 157 k, and NE = 16, C = 4 is the faster shape)
      const double push_bw = C > 4 ? 6.0 : 25.0;
      const double exch = C > 1 ? layers * (C - 1) * (2.0 * g.MT / C * NE * 128.0 * g.nsplit) / push_bw + layers * 1500.0 : 0.0;
      const double t = waves * ((ingest > mma ? ingest : mma) + epi + exch);
      if (t < best_t) best_t = t, best = LaunchShape{NE, C};
    }
  }
  return best;
}

template <int NE, int ACT, bool LN>
static int launch(const ChainArgs& a, size_t smem_bytes, cudaStream_t st) {
  auto kfn = chain_mlp_kernel<NE, ACT, LN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(chain_mlp_kernel)");
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
  if (e != cudaSuccess) return cuda_fail(e, "chain_mlp_kernel launch");
  return DPPO_OK;
}

// co-resident clusters of C CTAs of the chain kernel on this device, from the occupancy calculator (the GPC layout strands
// SMs for the larger cluster sizes, so sm_count / C overestimates)
template <int NE, int ACT, bool LN>
static int query_clusters_t(int C, size_t smem_bytes) {
  auto kfn = chain_mlp_kernel<NE, ACT, LN>;
  if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(C) * kDefaultSmCount), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = unsigned(C), attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

static void fill_chain_clusters(dppo_ctx* ctx) {
  const MlpGeom& g = ctx->g;
  const int fallback[4] = {ctx->sm_count, ctx->sm_count / 2, (ctx->sm_count - 16) / 4, 16};
  for (int ni = 0; ni < 3; ++ni) {
    const int NE = 16 << ni;
    const size_t fixed = smem_fixed_bytes(g, NE);
    for (int ci = 0; ci < 4; ++ci) {
      const int C = 1 << ci;
      int n = 0;
      if (fixed + 2 * kTile <= 232448) {
        size_t nst = (232448 - fixed) / kTile;
        if (nst > size_t(kMaxStages)) nst = kMaxStages;
        const size_t smem = fixed + nst * kTile;
#define DPPO_Q(NE_, ACT_) (g.ln ? query_clusters_t<NE_, ACT_, true>(C, smem) : query_clusters_t<NE_, ACT_, false>(C, smem))
#define DPPO_QA(NE_) (g.act == DPPO_ACT_RELU ? DPPO_Q(NE_, DPPO_ACT_RELU) : DPPO_Q(NE_, DPPO_ACT_MISH))
        n = NE == 64 ? DPPO_QA(64) : (NE == 32 ? DPPO_QA(32) : DPPO_QA(16));
#undef DPPO_QA
#undef DPPO_Q
      }
      ctx->chain_clusters[ni][ci] = n > 0 ? n : fallback[ci];
    }
  }
  ctx->chain_clusters_known = true;
}

int small_chain_capacity(const dppo_ctx* ctx);
int sample_chain_small_impl(dppo_ctx* ctx, const float* state, int E, const float* noise, uint64_t seed, uint64_t offset,
                            int64_t env_offset, int deterministic, int use_base, float min_std, float* traj, float* chain,
                            cudaStream_t st);

int sample_chain_impl(dppo_ctx* ctx, const float* state, int E, const float* noise, uint64_t seed, uint64_t offset,
                      int64_t env_offset, int deterministic, int use_base, float min_std, float* traj, float* chain,
                      const float* chains_in, float* logp, cudaStream_t st) {
  const MlpGeom& g = ctx->g;
  // A handful of environments is pure latency: the weights-stationary cluster kernel (chain_small.cu) takes the call when
  // its geometry applies.  dppo_debug_set_shape(ctx, 0, -1) forces it, any explicit tile / cluster shape bypasses it,
  // DPPO_B200_SMALL=0 disables it.
  {
    static int env_small = -1;
    if (env_small < 0) {
      const char* e = getenv("DPPO_B200_SMALL");
      env_small = e ? atoi(e) : 1;
    }
    const bool forced = ctx->force_c == -1;
    if (!chains_in && (forced || (env_small && ctx->force_ne == 0 && ctx->force_c == 0))) {
      bool ok = E <= small_chain_capacity(ctx);
      for (int w = 0; w < 2 && ok; ++w)
        for (const float* q : ctx->nets[w].raw) ok = ok && (reinterpret_cast<uintptr_t>(q) & 15) == 0;
      if (ok) return sample_chain_small_impl(ctx, state, E, noise, seed, offset, env_offset, deterministic, use_base, min_std, traj, chain, st);
      if (forced) return set_error("small chain kernel: geometry, alignment or %d environments outside its range", E), DPPO_ERR_UNSUPPORTED;
    }
  }
  ChainArgs a{};
  a.D = g.D, a.Dc_in = g.Dc_in, a.Dc = g.Dc, a.H = g.H, a.nb = g.nb, a.act = g.act, a.ln = g.ln, a.CH = g.CH, a.CO = g.CO;
  a.MT = g.MT, a.KCH = g.KCH, a.KC0 = g.KC0, a.KCc = g.KCc, a.MTc = g.MTc, a.nsplit = g.nsplit;
  a.off_tb = uint32_t(g.off_tb), a.off_blk = uint32_t(g.off_blk), a.blk_stride = uint32_t(g.blk_stride);
  a.off_bout = uint32_t(g.off_bout), a.off_bc0 = uint32_t(g.off_bc0), a.off_bc1 = uint32_t(g.off_bc1);
  for (int w = 0; w < 2; ++w) a.tiles[w] = ctx->nets[w].tiles, a.side[w] = ctx->nets[w].side;
  a.n_cond_tiles = uint32_t(g.n_cond_tiles), a.n_step_tiles = uint32_t(g.n_step_tiles);
  a.off_step_tiles = g.off_step_tiles;
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

  if (!ctx->chain_clusters_known) fill_chain_clusters(ctx);
  LaunchShape shape = pick_shape(ctx, E);
  const bool force_pair = ctx->force_c == -2;  // dppo_debug_set_shape(ctx, 0, -2): the cta_group::2 pair kernel
  if (force_pair) shape = LaunchShape{64, 2};
  const int NE = shape.NE;
  a.C = shape.C;
  {
    static int env_early = -1;  // DPPO_B200_EARLY=0: block layers in the plain chunk-major order (A/B measurements)
    if (env_early < 0) {
      const char* e = getenv("DPPO_B200_EARLY");
      env_early = e ? atoi(e) : 1;
    }
    a.early = (env_early && !g.ln && g.MT / a.C == 2 && g.KCH >= 4) ? 1 : 0;
  }
  // 64 environments on a pair of CTAs without LayerNorm / cond_mlp: the cta_group::2 kernel (chain_pair.cu)
  if (chain_pair_applicable(g, NE, a.C, force_pair)) return launch_chain_pair(a, g, st);
  if (force_pair) return set_error("pair chain kernel: needs no LayerNorm / cond_mlp and an even number of M tiles"), DPPO_ERR_UNSUPPORTED;
  const size_t fixed = smem_fixed_bytes(g, NE);
  const size_t budget = 232448;
  if (fixed + 2 * kTile > budget) return set_error("chain kernel: geometry needs %zu B of shared memory", fixed), DPPO_ERR_INVALID;
  int nstage = int((budget - fixed) / kTile);
  if (nstage > kMaxStages) nstage = kMaxStages;
  static int env_stages = -1;  // bring-up: cap the ring depth (sensitivity measurements)
  if (env_stages < 0) {
    const char* e = getenv("DPPO_B200_STAGES");
    env_stages = e ? atoi(e) : 0;
  }
  if (env_stages >= 2 && env_stages < nstage) nstage = env_stages;
  a.nstage = nstage;
  const size_t smem_bytes = fixed + size_t(nstage) * kTile;

#define DPPO_LAUNCH(NE_, ACT_)                                             \
  (g.ln ? launch<NE_, ACT_, true>(a, smem_bytes, st) : launch<NE_, ACT_, false>(a, smem_bytes, st))
#define DPPO_LAUNCH_ACT(NE_) (g.act == DPPO_ACT_RELU ? DPPO_LAUNCH(NE_, DPPO_ACT_RELU) : DPPO_LAUNCH(NE_, DPPO_ACT_MISH))
  if (NE == 64) return DPPO_LAUNCH_ACT(64);
  if (NE == 32) return DPPO_LAUNCH_ACT(32);
  return DPPO_LAUNCH_ACT(16);
#undef DPPO_LAUNCH_ACT
#undef DPPO_LAUNCH
}

}  // namespace dppo

// bring-up: the launch shape the cost model picks for E environments and the occupancy table it uses (12 ints)
extern "C" int dppo_debug_get_shape(dppo_ctx* ctx, int E, int* tile_envs, int* cluster, int* table) {
  using namespace dppo;
  if (!ctx || ctx->kind != 0) return DPPO_ERR_INVALID;
  if (!ctx->chain_clusters_known) fill_chain_clusters(ctx);
  const LaunchShape s = pick_shape(ctx, E);
  if (tile_envs) *tile_envs = s.NE;
  if (cluster) *cluster = s.C;
  if (table)
    for (int i = 0; i < 12; ++i) table[i] = ctx->chain_clusters[i / 4][i % 4];
  return DPPO_OK;
}
