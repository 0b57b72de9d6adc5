// Micro-benchmark: mbarrier hop latency on B200: (a) plain arrive ping-pong between two warps, (b) the same with
// 128 waiters per barrier, (c) tcgen05.commit as the arrival (no MMAs outstanding).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../openvis_b200/csrc/ptx.cuh"
using namespace ovis;
__global__ void pingpong(long long* out, int iters, int mode) {
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwaiters = (mode == 1) ? (blockDim.x / 32 - 1) : 1;   // warps that wait on bar[0]
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], mode == 1 ? nwaiters : 1); fence_mbar_init(); }
  if (warp == 0 && mode == 2) tmem_alloc(&tmem_holder, 32);
  __syncthreads();
  long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        if (mode == 2) umma_commit(&bar[0]); else mbar_arrive(&bar[0]);
        mbar_wait(&bar[1], i & 1);
      }
    }
  } else if (warp <= nwaiters) {
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&bar[0], i & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[1]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (warp == 0 && mode == 2) tmem_dealloc(tmem_holder, 32);
  if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int iters = 10000;
  const char* names[3] = {"arrive, 1 waiting warp (lane0 + 31 lanes)", "arrive, 4 waiting warps", "tcgen05.commit, 1 waiting warp"};
  for (int mode = 0; mode < 3; ++mode) {
    int threads = mode == 1 ? 160 : 64;
    pingpong<<<1, threads>>>(d, iters, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-45s round trip %.0f cycles (%s)\n", names[mode], (double)h / iters, cudaGetErrorString(e));
  }
  return 0;
}
