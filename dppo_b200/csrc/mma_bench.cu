// Micro-benchmark (bring-up aid, not on the product path): cycles per tcgen05.mma for M=128, N in {32..256}, K=16 bf16,
// with the accumulator rotating over n_acc TMEM tiles and the A operand rotating over n_a shared-memory tiles.
#include "common.cuh"

namespace dppo {

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int N, int n_acc, int n_a, int n_mma, int per_commit,
                                                           unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
    uint32_t phase = 0;
    const long long t0 = clock64();
    int since = 0;
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tmem + uint32_t(i % n_acc) * N;
      const uint32_t aa = a0 + uint32_t((i / 4) % n_a) * 16384 + (i % 4) * 32;
      umma_bf16(d, umma_desc(aa), umma_desc(b0 + (i % 4) * 32), idesc, 1u);
      if (++since == per_commit) {
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
        since = 0;
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, phase);
    out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace dppo

extern "C" int dppo_debug_mma_bench(int N, int n_acc, int n_a, int n_mma, int per_commit, unsigned long long* out,
                                    void* stream) {
  using namespace dppo;
  const int smem = 161 * 1024 + 1024;
  if (cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -2;
  mma_bench_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(N, n_acc, n_a, n_mma, per_commit, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
