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
#include "xattn.cuh"

namespace ovis {

struct SanAttnArgs {
  const __half* qkv;     // [B*(Q+1+L)][3*heads*64]: q | k | v, biases included, unscaled
  const float* pooled;   // [B*heads][Q][L] fp32 (adaptive-max-pooled attention biases) or null (no bias)
  __half* out;           // [B*(Q+1+L)][heads*64]
  int Q, L, heads;
  float scale_log2;      // 64^-1/2 * log2(e)
  int B;                 // images (san_attn_tc_kernel's persistent grid)
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

// Tensor-core version (mma.sync.m16n8k16, fp32 accumulate): CTA = (64-row tile, head, frame), 4 warps x 16 rows.
// K / V of the CLS + patch tokens sit in shared memory (row stride 72 halves: conflict-free fragment loads and
// ldmatrix), flash-style loop over 64-key tiles.  A SOS row's own key (the diagonal 0 of the bias matrix) is the initial
// state of its online softmax: m = q.k_self, l = 1, O = v_self.  The SIMT kernel above stays as the checker (tests).
constexpr int SA_LD = 72;
constexpr int SA_KT = 64;

__global__ void __launch_bounds__(128)
san_attn_mma_kernel(const SanAttnArgs a) {
  constexpr int D = 64;
  extern __shared__ __align__(16) __half sm_kv2[];
  const int head = blockIdx.y, b = blockIdx.z;
  const int Lt = a.Q + 1 + a.L, W = a.heads * D, W3 = 3 * W;
  const int nkeys = 1 + a.L;                               // CLS + patches
  const int ntiles = (nkeys + SA_KT - 1) / SA_KT;
  __half* sK = sm_kv2;
  __half* sV = sm_kv2 + ntiles * SA_KT * SA_LD;
  const __half* base = a.qkv + (long long)b * Lt * W3;
  for (int i = threadIdx.x; i < ntiles * SA_KT * 8; i += blockDim.x) {
    const int tok = i >> 3, ch = i & 7;
    uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;        // rows past the last key are zero (P = 0 there)
    if (tok < nkeys) {
      const __half* rowp = base + (long long)(a.Q + tok) * W3 + head * D + ch * 8;
      kk = *reinterpret_cast<const uint4*>(rowp + W);
      vv = *reinterpret_cast<const uint4*>(rowp + 2 * W);
    }
    *reinterpret_cast<uint4*>(sK + tok * SA_LD + ch * 8) = kk;
    *reinterpret_cast<uint4*>(sV + tok * SA_LD + ch * 8) = vv;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane >> 2, tq = lane & 3;
  const int row_base = blockIdx.x * 64 + warp * 16;
  if (row_base >= Lt) return;
  const float LOG2E = 1.4426950408889634f;
  int rows[2];
  bool ok[2], sos[2];
#pragma unroll
  for (int hi = 0; hi < 2; ++hi) {
    const int r = row_base + quad + hi * 8;
    ok[hi] = r < Lt;
    rows[hi] = ok[hi] ? r : Lt - 1;
    sos[hi] = rows[hi] < a.Q;
  }
  // Q fragments: a[i]: row = quad + (i & 1) * 8, col = ks * 16 + tq * 2 + (i >> 1) * 8
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      qf[ks][i] = *reinterpret_cast<const uint32_t*>(base + (long long)rows[i & 1] * W3 + head * D + ks * 16 + tq * 2 + (i >> 1) * 8);

  float o[8][4], mrow[2], lrow[2];
  // initial state: SOS rows start from their own key, the other rows from nothing
#pragma unroll
  for (int hi = 0; hi < 2; ++hi) {
    const __half* qp = base + (long long)rows[hi] * W3 + head * D;
    float sp = 0.f;
#pragma unroll
    for (int d = 0; d < 16; d += 2) {                      // the quad's four lanes share the 64-dim dot product
      const float2 qv = __half22float2(*reinterpret_cast<const __half2*>(qp + tq * 16 + d));
      const float2 kv = __half22float2(*reinterpret_cast<const __half2*>(qp + W + tq * 16 + d));
      sp = fmaf(qv.x, kv.x, sp);
      sp = fmaf(qv.y, kv.y, sp);
    }
    sp += __shfl_xor_sync(0xffffffffu, sp, 1);
    sp += __shfl_xor_sync(0xffffffffu, sp, 2);
    mrow[hi] = sos[hi] ? sp * a.scale_log2 : -INFINITY;
    lrow[hi] = sos[hi] ? 1.f : 0.f;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(qp + 2 * W + dn * 8 + tq * 2));
      o[dn][hi * 2] = sos[hi] ? vv.x : 0.f;
      o[dn][hi * 2 + 1] = sos[hi] ? vv.y : 0.f;
    }
  }
  const float* pb[2];
#pragma unroll
  for (int hi = 0; hi < 2; ++hi)
    pb[hi] = (a.pooled && sos[hi]) ? a.pooled + (((long long)b * a.heads + head) * a.Q + rows[hi]) * a.L : nullptr;

  for (int t = 0; t < ntiles; ++t) {
    const __half* kt = sK + t * SA_KT * SA_LD;
    const __half* vt = sV + t * SA_KT * SA_LD;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kt + (nt * 8 + quad) * SA_LD + ks * 16 + tq * 2);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kt + (nt * 8 + quad) * SA_LD + ks * 16 + tq * 2 + 8);
        mma_16816(s[nt], qf[ks], b0, b1);
      }
    }
    // scale, bias, structural mask; row max
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int hi = i >> 1;
        const int key = t * SA_KT + nt * 8 + tq * 2 + (i & 1);      // 0 = CLS, 1.. = patches
        float v = s[nt][i] * a.scale_log2;
        if (sos[hi]) {
          if (key == 0) v = -INFINITY;                                // SOS -> CLS carries -100
          else if (pb[hi] && key < nkeys) v = fmaf(__ldg(pb[hi] + key - 1), LOG2E, v);
        }
        if (key >= nkeys) v = -INFINITY;
        s[nt][i] = v;
        mx[hi] = fmaxf(mx[hi], v);
      }
    }
    float corr[2], muse[2];
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 1));
      mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 2));
      const float mnew = fmaxf(mrow[hi], mx[hi]);
      muse[hi] = (mnew == -INFINITY) ? 0.f : mnew;
      corr[hi] = exp2f(mrow[hi] - muse[hi]);
      mrow[hi] = mnew;
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = exp2f(s[nt][i] - muse[i >> 1]);
        s[nt][i] = p;
        ls[i >> 1] += p;
      }
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      ls[hi] += __shfl_xor_sync(0xffffffffu, ls[hi], 1);
      ls[hi] += __shfl_xor_sync(0xffffffffu, ls[hi], 2);
      lrow[hi] = lrow[hi] * corr[hi] + ls[hi];
    }
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      o[dn][0] *= corr[0]; o[dn][1] *= corr[0];
      o[dn][2] *= corr[1]; o[dn][3] *= corr[1];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t pa[4];
      pa[0] = pack_half2(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack_half2(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack_half2(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack_half2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, vt + (j * 16 + (lane & 15)) * SA_LD + dn * 8);
        mma_16816(o[dn], pa, b0, b1);
      }
    }
  }
#pragma unroll
  for (int hi = 0; hi < 2; ++hi) {
    if (!ok[hi]) continue;
    const float inv = 1.f / lrow[hi];
    __half* op = a.out + ((long long)b * Lt + rows[hi]) * W + head * D + tq * 2;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn)
      *reinterpret_cast<uint32_t*>(op + dn * 8) = pack_half2(o[dn][hi * 2] * inv, o[dn][hi * 2 + 1] * inv);
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
