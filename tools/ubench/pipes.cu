// Micro-benchmark: per-SM throughput of MUFU.EX2, ALU (FSEL/LOP3) and FMA-pipe (FADD) instructions on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = seed + i * 0.01f + threadIdx.x * 1e-4f;
  unsigned m = threadIdx.x * 2654435761u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      else if (OP == 1) { x[i] = (m & (1u << i)) ? -1.f : x[i]; asm volatile("" : "+f"(x[i])); }
      else if (OP == 2) { asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed)); }
      else { asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed)); }
    }
    m = m * 3u + 1u;
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, int warps) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<148, warps * 32>>>(out, 100, 0.5f);
  cudaEventRecord(a); k<OP><<<148, warps * 32>>>(out, iters, 0.5f); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = 148.0 * warps * 32 * 16.0 * iters;
  printf("%-8s warps/SM=%2d: %.1f Gops/s total, %.2f lane-ops/ns/SM (at 1.9 GHz: %.1f lanes/clk/SM)\n", name, warps, ops / ms / 1e6,
         ops / ms / 1e6 / 148, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16}) { run<0>("ex2", w); run<1>("sel", w); run<2>("fadd", w); run<3>("fmax", w); }
  return 0;
}
