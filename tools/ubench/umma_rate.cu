// UMMA issue-rate probe: cycles per tcgen05.mma (M = 128, K = 16, fp16) by operand layout and N, one CTA per SM,
// operands resident in shared memory (contents irrelevant), 2048 back-to-back MMAs into one accumulator.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../openvis_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include "ptx.cuh"
using namespace ovis;
__device__ __forceinline__ uint64_t dsc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// mode 0: A K-major SW128, B K-major SW128      (S^T-like)
// mode 1: A MN-major SW128 (2 atoms), B MN-major SW128 (1 atom)   (tc3 PV-like, N <= 64)
// mode 2: A K-major SW128, B MN-major SW128     (tc2 PV-like)
// mode 3: A MN-major SW128 (2 atoms), B K-major SW128
__global__ void __launch_bounds__(128) rate(long long* out, int mode, int N, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536 + 32768);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  fence_async_proxy();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(holder, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *holder;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(sm), b = smem_u32(sm) + 65536;
    const bool a_mn = mode == 1 || mode == 3, b_mn = mode == 1 || mode == 2;
    const uint32_t id = idesc(128, N, a_mn, b_mn);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t ad = a_mn ? dsc(a + kk * 2048, 16384, 1024, 2) : dsc(a + (kk & 3) * 32 + (kk >> 2) * 16384, 16, 1024, 2);
        const uint64_t bd = b_mn ? dsc(b + kk * 2048, 16384, 1024, 2) : dsc(b + (kk & 3) * 32, 16, 1024, 2);
        umma_f16(tmem, ad, bd, id, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[4] = {"A K-major, B K-major", "A MN-major (2 atoms), B MN-major", "A K-major, B MN-major", "A MN-major (2 atoms), B K-major"};
  for (int mode = 0; mode < 4; ++mode)
    for (int N : {16, 32, 48, 64, 112, 128, 256}) {
      if ((mode == 1 || mode == 2) && N > 128) continue;
      const int iters = 256;
      rate<<<148, 128, 100 * 1024>>>(d, mode, N, iters);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("%-34s N=%3d: %6.1f cycles per UMMA 128xNx16  (%5.1f %% of the dense fp16 peak of 4096 MAC/clk/SM)\n", names[mode], N,
             (double)c / (iters * 8), 100.0 * 128 * N * 16 / ((double)c / (iters * 8)) / 4096);
    }
  return 0;
}
