// Probe of the UMMA shared-memory descriptor variants the transposed cross-attention kernel (xattn_tc3) relies on:
//   T1  K-major SWIZZLE_32B operands (rows of 16 halfs = 32 B, 8-row atoms of 256 B, SBO = 256)       A [128 x 16], B [N x 16]
//   T2  K-major SWIZZLE_NONE operands (8x8 core matrices of 128 B, LBO = 128 along K, SBO = 256)
//   T3  T1 with SBO = 0 for A: all 8-row atoms alias one atom (a "broadcast row" operand in 256 B)
//   T4  MN-major SWIZZLE_128B A operand with M = 128 = two 64-wide atoms (LBO = 16 KB), K = 128 (SBO = 1024), B MN-major N = 32
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../openvis_b200/csrc -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "ptx.cuh"
using namespace ovis;

__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode 1/2/3: D[128 x N] = A[128 x 16] * B[N x 16]^T ; mode 4: D[128 x 32] = A[128(M) x 128(K)] * B[128(K) x 32(N)], both MN-major
__global__ void __launch_bounds__(128) probe(const __half* gA, const __half* gB, float* gD, int mode, int N) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sm;                 // up to 32 KB
  uint8_t* sB = sm + 32768;         // up to 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 32768 + 16384);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  for (int i = tid; i < (32768 + 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  if (mode == 1 || mode == 3) {         // SW32 K-major: (r, k) -> (r/8)*256 + (r%8)*32 + (((k/8) ^ ((r%8)>>2)) * 16) + (k%8)*2
    const int rowsA = mode == 3 ? 8 : 128;
    for (int i = tid; i < rowsA * 16; i += 128) {
      const int r = i / 16, k = i % 16;
      *reinterpret_cast<__half*>(sA + (r / 8) * 256 + (r % 8) * 32 + (((k / 8) ^ ((r % 8) >> 2)) * 16) + (k % 8) * 2) = gA[(mode == 3 ? 0 : r) * 16 + k];
    }
    for (int i = tid; i < N * 16; i += 128) {
      const int r = i / 16, k = i % 16;
      *reinterpret_cast<__half*>(sB + (r / 8) * 256 + (r % 8) * 32 + (((k / 8) ^ ((r % 8) >> 2)) * 16) + (k % 8) * 2) = gB[r * 16 + k];
    }
  } else if (mode == 2) {               // no swizzle: (r, k) -> (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2
    for (int i = tid; i < 128 * 16; i += 128) {
      const int r = i / 16, k = i % 16;
      *reinterpret_cast<__half*>(sA + (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2) = gA[r * 16 + k];
    }
    for (int i = tid; i < N * 16; i += 128) {
      const int r = i / 16, k = i % 16;
      *reinterpret_cast<__half*>(sB + (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2) = gB[r * 16 + k];
    }
  } else {                              // MN-major SW128: element (mn, k) -> (mn/64)*AS + (k/8)*KS + (k%8)*128 + ((((mn%64)/8) ^ (k%8)) * 16) + (mn%8)*2
    const int AS = mode == 6 ? 1024 : 16384, KS = mode == 6 ? 2048 : 1024;
    for (int i = tid; i < 128 * 128; i += 128) {
      const int m = i / 128, k = i % 128;      // gA[m][k]
      *reinterpret_cast<__half*>(sA + (m / 64) * AS + (k / 8) * KS + (k % 8) * 128 + ((((m % 64) / 8) ^ (k % 8)) * 16) + (m % 8) * 2) = gA[m * 128 + k];
    }
    for (int i = tid; i < 128 * 32; i += 128) {
      const int k = i / 32, n = i % 32;        // gB[k][n]
      *reinterpret_cast<__half*>(sB + (k / 8) * 1024 + (k % 8) * 128 + (((n / 8) ^ (k % 8)) * 16) + (n % 8) * 2) = gB[k * 32 + n];
    }
  }
  fence_async_proxy();
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (tid < 32) tmem_alloc(holder, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *holder;
  if (tid == 0) {
    if (mode == 1) umma_f16(tmem, desc_generic(smem_u32(sA), 16, 256, 6), desc_generic(smem_u32(sB), 16, 256, 6), idesc(128, N, 0, 0), 0u);
    if (mode == 3) umma_f16(tmem, desc_generic(smem_u32(sA), 16, 0, 6), desc_generic(smem_u32(sB), 16, 256, 6), idesc(128, N, 0, 0), 0u);
    if (mode == 2) umma_f16(tmem, desc_generic(smem_u32(sA), 128, 256, 0), desc_generic(smem_u32(sB), 128, 256, 0), idesc(128, N, 0, 0), 0u);
    if (mode >= 4) {
      // 4: LBO = atom stride (16 KB), SBO = k-group stride (1 KB); 5: the two fields swapped on the same data;
      // 6: atoms adjacent inside a k-group: LBO = 1 KB, SBO = 2 KB;  7: as 6 with the fields swapped
      const uint32_t lbo = mode == 4 ? 16384 : mode == 5 ? 1024 : mode == 6 ? 1024 : 2048;
      const uint32_t sbo = mode == 4 ? 1024 : mode == 5 ? 16384 : mode == 6 ? 2048 : 1024;
      const uint32_t kstep = (mode == 6 || mode == 7) ? 4096 : 2048;
      for (int kk = 0; kk < 8; ++kk)
        umma_f16(tmem, desc_generic(smem_u32(sA) + kk * kstep, lbo, sbo, 2), desc_generic(smem_u32(sB) + kk * 2048, 4096, 1024, 2),
                 idesc(128, 32, 1, 1), kk > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int ncols = mode >= 4 ? 32 : N;
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)((tid / 32) * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < ncols; ++j) gD[tid * ncols + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 256);
}

int main() {
  const int N = 112;
  std::vector<__half> hA(128 * 128), hB(128 * 128);
  std::vector<float> fA(128 * 128), fB(128 * 128);
  srand(1);
  for (int i = 0; i < 128 * 128; ++i) {
    fA[i] = (float)(rand() % 17 - 8) / 8.f; fB[i] = (float)(rand() % 13 - 6) / 4.f;
    hA[i] = __float2half(fA[i]); hB[i] = __float2half(fB[i]);
  }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, 128 * 128 * 2); cudaMalloc(&dB, 128 * 128 * 2); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dA, hA.data(), 128 * 128 * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), 128 * 128 * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
  std::vector<float> hD(128 * 128);
  for (int mode = 1; mode <= 7; ++mode) {
    cudaMemset(dD, 0, 128 * 128 * 4);
    probe<<<1, 128, 52 * 1024>>>(dA, dB, dD, mode, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, err_lo = 0, err_hi = 0;
    const int ncols = mode >= 4 ? 32 : N;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < ncols; ++c) {
        double ref = 0;
        if (mode >= 4) { for (int k = 0; k < 128; ++k) ref += (double)fA[r * 128 + k] * fB[k * 32 + c]; }
        else { for (int k = 0; k < 16; ++k) ref += (double)fA[(mode == 3 ? 0 : r) * 16 + k] * fB[c * 16 + k]; }
        maxerr = fmax(maxerr, fabs(ref - hD[r * ncols + c]));
        if (r < 64) err_lo = fmax(err_lo, fabs(ref - hD[r * ncols + c])); else err_hi = fmax(err_hi, fabs(ref - hD[r * ncols + c]));
      }
    if (mode >= 4) printf("   rows 0-63 max err %.4g, rows 64-127 max err %.4g\n", err_lo, err_hi);
    printf("mode %d (%s): max err %.4g  %s\n", mode,
           mode == 1 ? "K-major SW32" : mode == 2 ? "K-major no swizzle" : mode == 3 ? "SW32, A with SBO = 0 (broadcast atom)" : mode == 4 ? "MN-major SW128 A: LBO = atom stride 16K, SBO = k-group 1K" :
           mode == 5 ? "MN-major A: same data, LBO / SBO swapped" : mode == 6 ? "MN-major A: atoms adjacent (LBO 1K, SBO 2K)" : "MN-major A: atoms adjacent, fields swapped",
           maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  }
  return 0;
}
