// Micro-benchmarks behind two bring-up entry points (not on the product path, not in the public header):
//   dppo_debug_mma_rate     cycles per tcgen05.mma (M=128, K=16, bf16, both operands in shared memory) as a function of
//                           N, issued the way the chain kernel issues them (uniform warp loop, one elected lane)
//   dppo_debug_stream_rate  bytes/cycle/SM a grid of CTAs gets when every CTA streams the SAME region of global memory
//                           (L2 resident) into a shared-memory ring with 16 KiB bulk copies - optionally as thread-block
//                           clusters in which each CTA fetches 1/C of every tile and multicasts it to its peers
// Both tell the chain kernel which resource bounds it (DESIGN.md "Chain kernel roofline").
#include "common.cuh"

namespace dppo {

// `load` adds what runs next to the MMA stream in the chain kernel (bit mask):
//   1  warp 1 keeps 4 x 16 KiB bulk copies global -> shared in flight (weight ingest; `region` = 64 KiB of global memory)
//   2  warps 2, 3 read the accumulator region back with tcgen05.ld in a loop (epilogue TMEM reads)
//   4  warps 2, 3 store 4-byte words to shared memory in a loop (epilogue operand stores)
// out[0] = cycles of the MMA stream, out[1] = bytes ingested meanwhile, out[2] = tcgen05.ld / store iterations
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int n_mma, int n_b, int reuse, int load,
                                                          const uint8_t* __restrict__ region, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, ring_bar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (192 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ring_bar[i], 1);
    fence_mbar_init();
    done = 0;
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // A: 6 tiles of 128 x 64 (16 KiB each) at [0, 96 KiB); B: n_b operands of N x 64 at 96 KiB (<= 32 KiB);
  // ingest ring: 4 x 16 KiB at 128 KiB; store scratch at 192 KiB - 8 KiB .. (inside the ring's last stage when load & 4 only)
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a0 = umma_desc_lo(smem_u32(smem)), b0 = umma_desc_lo(smem_u32(smem + 96 * 1024));
    const uint32_t n_acc = 512 / N >= 4 ? 4 : 512 / N;
    uint32_t at = 0, acc = 0, bt = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      if (elect_one()) {
        const uint32_t d = tmem + acc * N;
        const uint32_t aa = a0 + at * (16384 / 16), bb = b0 + bt * (uint32_t(N) * 128 / 16);
        if (reuse) {
          // pairs that share their A operand (the chain kernel's w_hi * x_hi, w_hi * x_lo): second one reads A from the collector
          const uint32_t b2 = b0 + ((bt + 1) % uint32_t(n_b)) * (uint32_t(N) * 128 / 16);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_bf16_lo_fill(d, aa + 2 * k, bb + 2 * k, idesc, true);
            umma_bf16_lo_lastuse(d, aa + 2 * k, b2 + 2 * k, idesc);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_lo(d, aa + 2 * k, bb + 2 * k, idesc, true);
        }
      }
      __syncwarp();
      if (++at == 6) at = 0;
      if (++acc == n_acc) acc = 0;
      if (++bt == uint32_t(n_b)) bt = 0;
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) out[0] = clock64() - t0;
    done = 1;
  } else if (warp == 1) {
    if ((load & 1) && (threadIdx.x & 31) == 0) {
      unsigned long long bytes = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      for (int st = 0; st < 4; ++st) {
        mbar_arrive_expect_tx(&ring_bar[st], 16384);
        bulk_g2s(smem + 128 * 1024 + st * 16384, region + st * 16384, 16384, &ring_bar[st]);
      }
      int st = 0;
      while (!done) {
        mbar_wait(&ring_bar[st], ph[st]);
        ph[st] ^= 1;
        bytes += 16384;
        mbar_arrive_expect_tx(&ring_bar[st], 16384);
        bulk_g2s(smem + 128 * 1024 + st * 16384, region + st * 16384, 16384, &ring_bar[st]);
        st = (st + 1) & 3;
      }
      for (int i = 0; i < 4; ++i, st = (st + 1) & 3) mbar_wait(&ring_bar[st], ph[st]);  // drain before exit
      out[1] = bytes;
    }
  } else {
    unsigned long long it = 0;
    if (load & 2) {
      float sink = 0.f;
      while (!__shfl_sync(0xffffffffu, int(done), 0)) {  // warp-uniform exit: tcgen05.ld is .sync.aligned
        float v[32];
        tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + uint32_t((it & 7) * 32), v);
        sink += v[0] + v[31];
        ++it;
      }
      if (sink == 12345.f) out[3] = 1;
    }
    if (load & 4) {
      uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + 120 * 1024);  // 8 KiB behind the B operands, nobody reads it
      while (!__shfl_sync(0xffffffffu, int(done), 0)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) scratch[((it * 8 + j) & 31) * 64 + (warp - 2) * 32 + (threadIdx.x & 31)] = uint32_t(it);
        ++it;
      }
    }
    if ((threadIdx.x & 31) == 0 && warp == 2) out[2] = it;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------- streaming
constexpr int kSbStages = 8;
constexpr uint32_t kSbTile = 16384;

__global__ void __launch_bounds__(64, 1) stream_rate_kernel(const uint8_t* __restrict__ region, int region_tiles,
                                                            int n_tiles, int csz, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kSbStages], empty[kSbStages];
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = csz > 1 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSbStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], csz);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (csz > 1) cluster_sync_all();
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t stage = 0, phase = 0;
    const uint32_t slice = kSbTile / csz;
    for (int i = 0; i < n_tiles; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[stage], kSbTile);
        const uint8_t* src = region + size_t(i % region_tiles) * kSbTile;
        if (csz == 1)
          bulk_g2s(ring + size_t(stage) * kSbTile, src, kSbTile, &full[stage]);
        else
          bulk_g2s_multicast(ring + size_t(stage) * kSbTile + rank * slice, src + rank * slice, slice, &full[stage],
                             uint16_t((1u << csz) - 1));
      }
      __syncwarp();
      if (++stage == kSbStages) stage = 0, phase ^= 1;
    }
  } else {
    uint32_t stage = 0, phase = 0;
    for (int i = 0; i < n_tiles; ++i) {
      mbar_wait(&full[stage], phase);
      if (elect_one()) {
        if (csz == 1) {
          mbar_arrive(&empty[stage]);
        } else {
          for (int r = 0; r < csz; ++r) mbar_arrive_remote(&empty[stage], uint32_t(r));
        }
      }
      __syncwarp();
      if (++stage == kSbStages) stage = 0, phase ^= 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  if (csz > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it
}


// ---------------------------------------------------------------------------------------------------- CTA pair
// tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, each CTA holds 128 rows of A, N/2 rows of B and a 128-lane x N
// accumulator).  First one K = 64 pass over known small-integer operands whose accumulator both CTAs write to `d`
// ((2, 128, N) fp32) - the layout check - then the issue rate of n_mma instructions like mma_rate_kernel measures it.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint64_t a = (uint64_t(kDescHi32) << 32) | a_lo, b = (uint64_t(kDescHi32) << 32) | b_lo;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}

__host__ __device__ inline int pair_a_val(int cta, int r, int k) { return (r + 3 * k + 7 * cta) % 5 - 2; }
__host__ __device__ inline int pair_b_val(int n, int k) { return (2 * n + k) % 7 - 3; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    mma_pair_rate_kernel(int N, int n_mma, float* d, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = int(cluster_ctarank());
  const int Nh = N / 2;
  // A: 8 tiles of 128 x 64 at [0, 128 KiB) (tile 0 holds the known values, the rest zeros); B: Nh x 64 at 128 KiB
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(smem + sw128_offset(r, k, 128)) = __float2bfloat16_rn(float(pair_a_val(cta, r, k)));
  }
  for (int i = threadIdx.x; i < Nh * 64; i += blockDim.x) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(smem + 128 * 1024 + sw128_offset(r, k, Nh)) =
        __float2bfloat16_rn(float(pair_b_val(cta * Nh + r, k)));
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(&slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' operands and barriers are in place before the leader issues
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = umma_idesc_bf16(256, N);
  const uint32_t a0 = umma_desc_lo(smem_u32(smem)), b0 = umma_desc_lo(smem_u32(smem + 128 * 1024));
  if (warp == 0 && cta == 0) {
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, a0 + 2 * k, b0 + 2 * k, idesc, k > 0);
      umma_commit_pair(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + uint32_t(c0), v);
    for (int j = 0; j < 32; ++j) d[(size_t(cta) * 128 + warp * 32 + lane) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 0 && cta == 0) {
    uint32_t at = 0;
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      if (elect_one()) {
        const uint32_t aa = a0 + at * (16384 / 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, aa + 2 * k, b0 + 2 * k, idesc, 1);
      }
      __syncwarp();
      if (++at == 8) at = 0;
    }
    if (elect_one()) umma_commit_pair(&bar);
    __syncwarp();
    mbar_wait(&bar, 1);
    if (threadIdx.x == 0) out[0] = clock64() - t0;
  } else {
    mbar_wait(&bar, 1);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 256);
}

}  // namespace dppo

extern "C" int dppo_debug_mma_rate(int N, int n_mma, int n_b, int reuse, int load, const void* region,
                                   unsigned long long* out, void* stream) {
  using namespace dppo;
  if (N < 16 || N > 256 || N % 16 || n_b < 1 || n_b * N > 256) return -1;
  if ((load & 1) && region == nullptr) return -1;
  const int smem = 193 * 1024 + 1024;
  if (cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
  mma_rate_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(N, n_mma, n_b, reuse, load, static_cast<const uint8_t*>(region), out);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int dppo_debug_stream_rate(const void* region, int region_tiles, int n_tiles, int grid, int cluster,
                                      unsigned long long* out, void* stream) {
  using namespace dppo;
  if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) return -1;
  if (grid % cluster) return -1;
  const int smem = kSbStages * kSbTile + 1024;
  if (cudaFuncSetAttribute(stream_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(64), cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, stream_rate_kernel, static_cast<const uint8_t*>(region), region_tiles, n_tiles,
                                     cluster, out);
  return e == cudaSuccess ? 0 : -3;
}

// d: (2, 128, N) fp32 device buffer, expect: same shape on the host filled with the exact result (or null)
extern "C" int dppo_debug_mma_pair_rate(int N, int n_mma, float* d, unsigned long long* out, float* expect, void* stream) {
  using namespace dppo;
  if (N < 32 || N > 256 || N % 32) return -1;
  if (expect) {
    for (int cta = 0; cta < 2; ++cta)
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          int acc = 0;
          for (int k = 0; k < 64; ++k) acc += pair_a_val(cta, r, k) * pair_b_val(n, k);
          expect[(size_t(cta) * 128 + r) * N + n] = float(acc);
        }
  }
  const int smem = 161 * 1024 + 1024;
  if (cudaFuncSetAttribute(mma_pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
  mma_pair_rate_kernel<<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(N, n_mma, d, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// ------------------------------------------------------------------------------------------------ TMEM read rate
// 8 warps read a 128-lane x `cols`-column fp32 accumulator region back with tcgen05.ld in different shapes / batching
// (what an epilogue does), optionally next to a stream of N = 256 MMAs into the other half of TMEM.
//   variant 0: x16 + wait per load     1: x32 + wait per load     2: two x32 loads, one wait     3: four x32 loads, one wait
//   variant 4: x16 x 4 loads, one wait
// out[0] = cycles (max over warps is taken on the host: out[w]), out[8] = MMA cycles
namespace dppo {
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(320, 1) tmem_read_rate_kernel(int variant, int iters, int with_mma, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    done = 0;
  }
  if (warp == 1) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    if (with_mma) {
      const uint32_t idesc = umma_idesc_bf16(128, 256);
      const uint32_t a0 = umma_desc_lo(smem_u32(smem)), b0 = umma_desc_lo(smem_u32(smem + 32 * 1024));
      const long long t0 = clock64();
      long long n = 0;
      while (!done) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_lo(tmem + 256, a0 + 2 * k, b0 + 2 * k, idesc, true);
          umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, uint32_t(n) & 1u);
        ++n;
      }
      if (lane == 0) out[8] = clock64() - t0, out[9] = n * 4;
    }
  } else if (warp >= 2) {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t base = tmem + (uint32_t(q * 32) << 16) + uint32_t(half * 128);
    uint32_t acc = 0;
    named_bar_sync(1, 256);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (variant == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t r[16];
          tmem_ld16_nowait(base + c * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += r[i];
        }
      } else if (variant == 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32_nowait(base + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += r[i];
        }
      } else if (variant == 2) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r0[32], r1[32];
          tmem_ld32_nowait(base + c * 64, r0);
          tmem_ld32_nowait(base + c * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += r0[i] + r1[i];
        }
      } else if (variant == 3) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld32_nowait(base, r0);
        tmem_ld32_nowait(base + 32, r1);
        tmem_ld32_nowait(base + 64, r2);
        tmem_ld32_nowait(base + 96, r3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += r0[i] + r1[i] + r2[i] + r3[i];
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r0[16], r1[16], r2[16], r3[16];
          tmem_ld16_nowait(base + c * 64, r0);
          tmem_ld16_nowait(base + c * 64 + 16, r1);
          tmem_ld16_nowait(base + c * 64 + 32, r2);
          tmem_ld16_nowait(base + c * 64 + 48, r3);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += r0[i] + r1[i] + r2[i] + r3[i];
        }
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp - 2] = t1 - t0;
    if (acc == 0x12345678u) out[15] = acc;
    named_bar_sync(1, 256);
    if (threadIdx.x == 64) done = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}
}  // namespace dppo

// every epilogue warp reads 128 columns of its lane quarter per iteration (the 8 warps together: 128 lanes x 256 columns)
extern "C" int dppo_debug_tmem_read_rate(int variant, int iters, int with_mma, unsigned long long* out, void* stream) {
  using namespace dppo;
  const int smem = 97 * 1024;
  if (cudaFuncSetAttribute(tmem_read_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
  tmem_read_rate_kernel<<<1, 320, smem, static_cast<cudaStream_t>(stream)>>>(variant, iters, with_mma, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
