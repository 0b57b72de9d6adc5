// SAN / BriVIS side path: attention of the post-split CLIP blocks (BiasedResidualAttentionBlock, side_adapter.py:70-78)
// over the token sequence [Q SOS tokens | CLS | L patches] with the additive bias of SideAdapter._build_attn_biases
// (side_adapter.py:237-270) generated on the fly from the pooled per-head biases -- the [B*heads, Q+1+L, Q+1+L] fp32
// matrix is never materialised.  The matrix's structure is used, not just applied:
//   * columns of SOS keys carry -100 for every other row: exp(-100) is below fp32 resolution next to any finite logit,
//     so those keys are skipped; a SOS row keeps its own key (diagonal 0);
//   * SOS rows see CLS with -100 (skipped) and the L patches with pooled[b, head, q, patch];
//   * CLS / patch rows see CLS + patches with bias 0.
// Hence every row attends to at most L + 1 keys, and K / V of the CLS + patch tokens (shared by all rows of a
// (frame, head)) sit in shared memory.  d = 64; PARTS lanes per row, each with its own online softmax, merged by shuffles.
#pragma once
#include "ptx.cuh"

namespace ovis {

struct SanAttnArgs {
  const __half* qkv;     // [B*(Q+1+L)][3*heads*64]: q | k | v, biases included, unscaled
  const float* pooled;   // [B*heads][Q][L] fp32 (adaptive-max-pooled attention biases) or null (no bias)
  __half* out;           // [B*(Q+1+L)][heads*64]
  int Q, L, heads;
  float scale_log2;      // 64^-1/2 * log2(e)
};

__global__ void __launch_bounds__(256)
san_attn_kernel(const SanAttnArgs a) {
  constexpr int PARTS = 4, D = 64;
  extern __shared__ __half sm_kv[];            // K [1+L][64], V [1+L][64] of the CLS + patch tokens
  const int head = blockIdx.x, b = blockIdx.y;
  const int Lt = a.Q + 1 + a.L, W = a.heads * D, W3 = 3 * W;
  __half* sk = sm_kv;
  __half* sv = sm_kv + (1 + a.L) * D;
  const __half* base = a.qkv + (long long)b * Lt * W3;
  for (int i = threadIdx.x; i < (1 + a.L) * 8; i += blockDim.x) {
    const int tok = i >> 3, ch = i & 7;
    const __half* rowp = base + (long long)(a.Q + tok) * W3 + head * D + ch * 8;
    *reinterpret_cast<uint4*>(sk + tok * D + ch * 8) = *reinterpret_cast<const uint4*>(rowp + W);
    *reinterpret_cast<uint4*>(sv + tok * D + ch * 8) = *reinterpret_cast<const uint4*>(rowp + 2 * W);
  }
  __syncthreads();
  const float LOG2E = 1.4426950408889634f;
  const int part = threadIdx.x % PARTS;
  for (int row0 = 0; row0 < Lt; row0 += 256 / PARTS) {
    const int row_raw = row0 + threadIdx.x / PARTS;
    const bool row_ok = row_raw < Lt;                 // (whole lane groups are in or out)
    const int row = row_ok ? row_raw : Lt - 1;
    const bool sos = row < a.Q;
    const __half* qp = base + (long long)row * W3 + head * D;
    float q[D], acc[D];
#pragma unroll
    for (int d = 0; d < D; d += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(qp + d);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        q[d + 2 * e] = f.x * a.scale_log2; q[d + 2 * e + 1] = f.y * a.scale_log2;
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    auto visit = [&](const __half* kp, const __half* vp, float bias_log2) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(kp + d));
        s0 = fmaf(q[d], kf.x, s0);
        s1 = fmaf(q[d + 1], kf.y, s1);
      }
      const float s = (s0 + s1) + bias_log2;
      const float mn = fmaxf(m, s);
      const float c = exp2f(m - mn);
      const float p = exp2f(s - mn);
      l = l * c + p;
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(vp + d));
        acc[d] = fmaf(p, vf.x, acc[d] * c);
        acc[d + 1] = fmaf(p, vf.y, acc[d + 1] * c);
      }
      m = mn;
    };
    if (sos) {
      // own key (diagonal 0), then the L patches with their pooled bias; CLS (-100) is skipped
      if (part == 0) visit(qp + W, qp + 2 * W, 0.f);
      const float* pb = a.pooled ? a.pooled + (((long long)b * a.heads + head) * a.Q + row) * a.L : nullptr;
      for (int j = 1 + part; j <= a.L; j += PARTS) visit(sk + j * D, sv + j * D, pb ? __ldg(pb + j - 1) * LOG2E : 0.f);
    } else {
      for (int j = part; j <= a.L; j += PARTS) visit(sk + j * D, sv + j * D, 0.f);
    }
    // merge the key partitions of the row (adjacent lanes)
    float mt = m;
#pragma unroll
    for (int o = 1; o < PARTS; o <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
    const float sc = (m == -INFINITY) ? 0.f : exp2f(m - mt);
    l *= sc;
#pragma unroll
    for (int o = 1; o < PARTS; o <<= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float v = acc[d] * sc;
#pragma unroll
      for (int o = 1; o < PARTS; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[d] = v * inv;
    }
    if (row_ok) {      // each lane of the group writes 16 of the 64 channels
      constexpr int CH = D / PARTS;
      __half* op = a.out + ((long long)b * Lt + row) * W + head * D + part * CH;
#pragma unroll
      for (int e = 0; e < CH; e += 8) {
        float o[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          float v = acc[e + x];
#pragma unroll
          for (int pp = 1; pp < PARTS; ++pp) v = (part == pp) ? acc[pp * CH + e + x] : v;
          o[x] = v;
        }
        uint4 u;
        u.x = pack_half2(o[0], o[1]); u.y = pack_half2(o[2], o[3]); u.z = pack_half2(o[4], o[5]); u.w = pack_half2(o[6], o[7]);
        *reinterpret_cast<uint4*>(op + e) = u;
      }
    }
  }
}

// adaptive max-pool of the per-head attention biases to the CLIP grid (step 1 of _build_attn_biases,
// side_adapter.py:241-250): bias [BN][Q][h][w] -> pooled [BN][Q][gh*gw].  One thread per output element.
__global__ void __launch_bounds__(256)
san_pool_bias_kernel(const float* __restrict__ bias, float* __restrict__ pooled, long long total, int h, int w, int gh, int gw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int L = gh * gw;
  const int cell = (int)(i % L);
  const long long plane = i / L;
  const int gy = cell / gw, gx = cell % gw;
  const int y0 = (gy * h) / gh, y1 = ((gy + 1) * h + gh - 1) / gh;
  const int x0 = (gx * w) / gw, x1 = ((gx + 1) * w + gw - 1) / gw;
  const float* p = bias + plane * h * w;
  float m = -INFINITY;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) m = fmaxf(m, __ldg(p + y * w + x));
  pooled[i] = m;
}

}  // namespace ovis
