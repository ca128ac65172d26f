// Minimal single-CTA tcgen05 GEMM used to validate the descriptor encodings of common.cuh on real hardware:
//   C[128 x N] (fp32) = A[128 x K] (bf16, K-major) * B[N x K]^T (bf16, K-major),  K multiple of 64, N in {32, 64}.
// A arrives pre-swizzled from global memory through one bulk (TMA-engine) copy per 64-wide K chunk — exactly how the
// chain kernel streams weight tiles — and B is written to shared memory by ordinary stores the way the epilogue
// writes activations.  The descriptor high bits and the instruction descriptor are runtime arguments so that a test
// can probe encodings without a rebuild.
#include "common.cuh"

namespace dppo {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint8_t* __restrict__ a_tiles,  // K/64 tiles of 16 KiB, already in SW128 image
                     const float* __restrict__ b,          // N x K fp32 row-major (converted to bf16 here)
                     float* __restrict__ c,                // 128 x N fp32 row-major
                     int N, int K, uint64_t desc_hi, uint32_t idesc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = K / 64;
  uint8_t* sA = smem;                 // KC * 16 KiB
  uint8_t* sB = smem + KC * 16384;    // KC * N * 128 B
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  // B: generic-proxy stores into the swizzled K-major layout
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw128_offset(n, k, N)) = __float2bfloat16_rn(b[i]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, KC * 16384);
    for (int kc = 0; kc < KC; ++kc) bulk_g2s(sA + kc * 16384, a_tiles + size_t(kc) * 16384, 16384, &bar_load);
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    for (int kc = 0; kc < KC; ++kc) {
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint64_t da = umma_desc(smem_u32(sA + kc * 16384) + k16 * 32, desc_hi);
        const uint64_t db = umma_desc(smem_u32(sB + kc * N * 128) + k16 * 32, desc_hi);
        umma_bf16(tmem, da, db, idesc, (kc | k16) != 0);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) c[row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// host-side packer for the test: A (128 x K fp32) -> SW128 bf16 tiles
__global__ void pack_a_selftest_kernel(const float* __restrict__ a, uint8_t* __restrict__ tiles, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * K) return;
  const int r = i / K, k = i % K;
  const int kc = k >> 6;
  *reinterpret_cast<__nv_bfloat16*>(tiles + size_t(kc) * 16384 + sw128_offset(r, k & 63, 128)) =
      __float2bfloat16_rn(a[i]);
}

}  // namespace dppo

extern "C" int dppo_selftest_umma(const float* a, const float* b, float* c, void* scratch, int N, int K,
                                  uint64_t desc_hi, uint32_t idesc, void* stream_) {
  using namespace dppo;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (K % 64 != 0 || (N != 32 && N != 64) || K > 256) return -1;
  if (desc_hi == 0) desc_hi = kDescSw128KMajor;
  if (idesc == 0) idesc = umma_idesc_bf16(128, N);
  uint8_t* tiles = static_cast<uint8_t*>(scratch);  // >= K/64 * 16 KiB
  pack_a_selftest_kernel<<<(128 * K + 255) / 256, 256, 0, stream>>>(a, tiles, K);
  const int KC = K / 64;
  const size_t smem = size_t(KC) * 16384 + size_t(KC) * N * 128 + 1024;
  if (cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return -2;
  umma_selftest_kernel<<<1, 128, smem, stream>>>(tiles, b, c, N, K, desc_hi, idesc);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
