// Thin inline-PTX wrappers for sm_100a: mbarrier, bulk (TMA-engine) copies, tcgen05 MMA / TMEM.
// Hand-written for this project; encodings follow the PTX ISA (descriptor bit layout cross-checked against the
// CUTLASS headers vendored in the image: cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dppo {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  // the suspend-time hint lets the hardware park the warp instead of burning issue slots of its scheduler
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFF) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {
        printf("dppo_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

// Polling wait (mbarrier.test_wait, no suspend): for waits where the whole CTA has nothing else to issue and the wake-up
// latency of the suspending try_wait (NANOSLEEP.SYNCS) is on the critical path.  Bounded like mbar_wait.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  uint64_t t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if ((++spins & 0xFFFF) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {
        printf("dppo_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

// one lane of the (converged) warp gets true; keeps surrounding address arithmetic in the uniform datapath
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      :
      : "memory");
  return pred != 0;
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ bulk copy global -> shared (UBLKCP, TMA engine)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// expect-tx arrive + bulk copy, predicated on `leader` (issue loops without a divergent region, see chain_mlp.cu)
__device__ __forceinline__ void bulk_g2s_expect_p(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                  uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "r"(leader)
      : "memory");
}

// ------------------------------------------------------------------ thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the location with the same CTA-relative offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(bar), rank)) : "memory");
}
// The same arrive without the cluster-scope release: ONE instruction (SYNCS.ARRIVE), where the form above is
// MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR + SYNCS.ARRIVE (~1000 cycles when bulk copies are in flight,
// profiles/r2/r2i_chain_pair_kernel_and_early_order.txt).  For arrivals that publish no data written by this thread
// ("I am done reading", "your block has landed"); anything that follows remote st.shared::cluster stores needs the
// release form.
__device__ __forceinline__ void mbar_arrive_remote_nodata(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(bar), rank)) : "memory");
}
// store one float into CTA `rank`'s shared memory at the CTA-relative address of `local`
__device__ __forceinline__ void st_remote_f32(float* local, uint32_t rank, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(mapa_u32(smem_u32(local), rank)), "f"(v) : "memory");
}
// The same store as an ASYNC store that completes 4 bytes on the mbarrier `bar` of the destination CTA: the receiver
// waits for the expected byte count on its own barrier, no cluster-scope release fence on the sending side
// (barrier.cluster.arrive.release / mbarrier.arrive.release.cluster are MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in SASS).
__device__ __forceinline__ void st_async_f32(float* local, uint32_t rank, float v, uint64_t* bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(
                   mapa_u32(smem_u32(local), rank)),
               "f"(v), "r"(mapa_u32(smem_u32(bar), rank))
               : "memory");
}
// two floats (8-byte aligned), 8 bytes completed on the destination CTA's barrier
__device__ __forceinline__ void st_async_v2(float* local, uint32_t rank, float x, float y, uint64_t* bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(
                   mapa_u32(smem_u32(local), rank)),
               "f"(x), "f"(y), "r"(mapa_u32(smem_u32(bar), rank))
               : "memory");
}
// one arrival + `bytes` expected transaction bytes on the barrier of CTA `rank` (the sender announces how much it sends)
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint64_t* bar, uint32_t rank, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(mapa_u32(smem_u32(bar), rank)), "r"(bytes)
               : "memory");
}
// 16-byte variant (the address must be 16-byte aligned)
__device__ __forceinline__ void st_remote_v4(float* local, uint32_t rank, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mapa_u32(smem_u32(local), rank)), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// acquire at cluster scope: pairs with mbar_arrive_remote of a peer CTA
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0xFF) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {
        printf("dppo_b200: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}
// bulk copy from this CTA's shared memory into CTA `rank`'s shared memory (same CTA-relative destination offset as
// `dst_local`), completing `bytes` on that CTA's mbarrier `bar`
__device__ __forceinline__ void bulk_s2peer(void* dst_local, const void* src_local, uint32_t bytes, uint64_t* bar,
                                            uint32_t rank) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   mapa_u32(smem_u32(dst_local), rank)),
               "r"(smem_u32(src_local)), "r"(bytes), "r"(mapa_u32(smem_u32(bar), rank))
               : "memory");
}
// bulk copy global -> shared memory of every CTA in `mask` (same CTA-relative offsets), completing on each one's barrier
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// one full warp; writes the TMEM base address to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (ignored for swizzled K-major)
//   bits [32,46) stride byte offset >> 4     bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// A tile is rows x 64 bf16: row r at byte r*128, 16-byte chunk c of a row stored at chunk (c ^ (r & 7)); groups of
// 8 rows are SBO = 1024 bytes apart.  The tile base must be 1024-byte aligned.  Stepping K by 16 elements inside
// the 64-wide swizzle atom is +32 bytes on the start address.
constexpr uint64_t kDescSw128KMajor =
    (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint64_t high_bits = kDescSw128KMajor) {
  return high_bits | uint64_t((smem_addr >> 4) & 0x3FFF);
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K)  [16] B major (0 = K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// same, with the constant high word split off so that stepping K is a single 32-bit add on the low word
constexpr uint32_t kDescHi32 = uint32_t(kDescSw128KMajor >> 32);
constexpr uint32_t kDescLoFlags = uint32_t(kDescSw128KMajor & 0xFFFFFFFFull);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return kDescLoFlags | ((smem_addr >> 4) & 0x3FFF); }
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool accumulate) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  if (accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc) : "memory");
  }
}

// Accumulating MMA with an explicit A-collector policy: FILL keeps the A operand (128 x 16) in the tensor core's collector
// buffer, LASTUSE consumes it from there - the next MMA with the SAME A descriptor skips its shared-memory read of A.
__device__ __forceinline__ void umma_bf16_lo_fill(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool accumulate) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  if (accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc) : "memory");
  }
}
__device__ __forceinline__ void umma_bf16_lo_lastuse(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc) : "memory");
}

// Predicated forms for an issue loop whose control flow stays warp-uniform: every lane executes the statement, only the
// lane with `leader != 0` (elected once, in front of the loop) issues.  No branch around the MMAs, so the compiler
// keeps one convergent instruction stream (a single vote -> uniform predicate) instead of a divergent region per tile.
__device__ __forceinline__ void umma_bf16_lo_p(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                               uint32_t accumulate, uint32_t leader) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_p(uint64_t* bar, uint32_t leader) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}

// The four K = 16 steps of one 64-wide swizzle atom (descriptor start address + 32 bytes each) in ONE statement: one
// predicate set-up for four MMAs, and, with `commit_bar != 0`, the ring release behind them.
__device__ __forceinline__ void umma_bf16_lo_x4_p(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                  uint32_t accumulate_first, uint32_t leader, uint32_t commit_bar) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p, q, t, c;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\tsetp.eq.u32 t, 1, 1;\n\t"
      "setp.ne.and.b32 c, %6, 0, q;\n\t"
      "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
      "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t"
      "@c tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate_first), "r"(leader), "r"(commit_bar)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i), columns c..c+31.
// The issuing warp may only touch the lane quarter (warp_id % 4).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}


__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
        "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
        "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
        "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      :
      : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
        "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[32]) { tmem_st32(taddr, v); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[16]) { tmem_st16(taddr, v); }
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[8]) { tmem_ld8(taddr, v); }
__device__ __forceinline__ void tmem_st(uint32_t taddr, const float (&v)[8]) { tmem_st8(taddr, v); }

// named barrier among a subset of the CTA's warps
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ numerics helpers
// split an fp32 value into bf16 hi + bf16 lo (hi + lo carries 16 mantissa bits)
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// byte offset of element (row, k) inside a K-major SWIZZLE_128B operand made of 64-wide K chunks of `rows` rows each
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k, uint32_t rows) {
  const uint32_t chunk = k >> 6, kk = k & 63;
  return chunk * rows * 128u + row * 128u + ((((kk >> 3) ^ (row & 7u)) << 4) | ((kk & 7u) << 1));
}

// mish(x) = x * tanh(softplus(x)) = x * n / (n + 2), n = e^x (e^x + 2); torch switches softplus to identity above 20.
// ex2 / rcp in their .ftz forms: without .ftz the compiler wraps each in a denormal range fix-up (7 extra instructions
// per element in an epilogue that is instruction-issue bound); n + 2 >= 2 is never denormal, and an e^x flushed to
// zero only turns a result of magnitude < 1e-36 into 0.
__device__ __forceinline__ float mish_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(x, 20.0f) * 1.4426950408889634f));
  const float n = e * (e + 2.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.0f));
  return x * n * r;
}

}  // namespace dppo
