// HBM-bound layout / pooling kernels that feed the tcgen05 GEMMs.
#pragma once
#include "ptx.cuh"

namespace ovis {

// in [B][C][N] fp32 (NCHW, N = h*w)  ->  out [B][N][C] fp16 ("token-major", K-major GEMM operand)
// and optionally out_pos = fp16(in + pos[n][c] + pos_t[b][c]): the `key = memory + pos` operand
// (with_pos_embed, video_..._decoder.py:115-116; pos3d = pos2d + pos_z, position_encoding.py:163).
// Tile 64(n) x 64(c).  Reads are 256 B rows along n; every thread then writes 16 B (8 channels) so that a warp
// store covers four full 128-byte token rows.
__global__ void __launch_bounds__(256)
nchw_to_tokens_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, __half* __restrict__ out_pos,
                          const float* __restrict__ pos, const float* __restrict__ pos_t, int C, int N) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;     // 64 x 4
  const float* ib = in + (long long)b * C * N;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int c = ty + i * 4;
    const int n = n0 + tx;
    tile[c][tx] = (n < N && c0 + c < C) ? __ldg(ib + (long long)(c0 + c) * N + n) : 0.f;
  }
  __syncthreads();
  // item = (token, 8-channel group): 64 tokens x 8 groups = 512 items, two per thread
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int item = i * 256 + threadIdx.x;
    const int cg = item & 7, nl = item >> 3;
    const int n = n0 + nl, c = c0 + cg * 8;
    if (n >= N || c >= C) continue;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tile[cg * 8 + e][nl];
    const long long o = ((long long)b * N + n) * C + c;
    *reinterpret_cast<uint4*>(out + o) =
        make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
    if (out_pos) {
      // key operand: x + (level_embed + 2-D sine position)[n] (+ frame term of the 3-D embedding)
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + (long long)n * C + c));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + (long long)n * C + c) + 1);
      float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      if (pos_t) {
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(pos_t + (long long)b * C + c));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(pos_t + (long long)b * C + c) + 1);
        p[0] += t0.x; p[1] += t0.y; p[2] += t0.z; p[3] += t0.w; p[4] += t1.x; p[5] += t1.y; p[6] += t1.z; p[7] += t1.w;
      }
      *reinterpret_cast<uint4*>(out_pos + o) =
          make_uint4(pack_half2(v[0] + p[0], v[1] + p[1]), pack_half2(v[2] + p[2], v[3] + p[3]),
                     pack_half2(v[4] + p[4], v[5] + p[5]), pack_half2(v[6] + p[6], v[7] + p[7]));
    }
  }
}

// mask_features F [B][C][H][W] fp32 (stride-4 map)  ->
//   ft [B][H*W][C] fp16                                   (operand of the final full-resolution mask GEMM)
//   g0/g1/g2 [B][(H/s)*(W/s)][C] fp16, s = 8/4/2          (mean of the centre 2x2 pixels of every s x s block)
// The centre-2x2 mean is exactly what F.interpolate(bilinear, align_corners=False) computes for an integer
// down-scale factor s (SURVEY.md Finding 3), and the mask logit is linear in F, so
//   sign(bilinear(mask_embed . F)) == sign(mask_embed . g_l)   up to fp rounding.
// Tile: 8 rows x 32 cols of pixels x 32 channels.
__global__ void __launch_bounds__(256)
maskfeat_prep_kernel(const float* __restrict__ F, __half* __restrict__ ft, __half* __restrict__ g0,
                     __half* __restrict__ g1, __half* __restrict__ g2, int C, int H, int W) {
  __shared__ float tile[32][257];     // [channel][pixel 8x32], +1 pad: conflict-free both ways
  const int b = blockIdx.z;
  const int tiles_x = (W + 31) / 32;
  const int x0 = (blockIdx.x % tiles_x) * 32, y0 = (blockIdx.x / tiles_x) * 8;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* fb = F + ((long long)b * C + c0) * H * W;
  const bool inx = x0 + tx < W;
#pragma unroll 8
  for (int ch = 0; ch < 32; ++ch)
    tile[ch][ty * 32 + tx] = inx ? __ldg(fb + (long long)ch * H * W + (long long)(y0 + ty) * W + x0 + tx) : 0.f;
  __syncthreads();

  // ---- full-resolution token-major fp16 copy: item = (pixel, 8-channel group), 16-byte stores
  {
    __half* ob = ft + (long long)b * H * W * C + c0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int item = i * 256 + threadIdx.x;
      const int cg = item & 3, pix = item >> 2;
      const int py = pix >> 5, px = pix & 31;
      if (x0 + px < W) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = tile[cg * 8 + e][pix];
        *reinterpret_cast<uint4*>(ob + ((long long)(y0 + py) * W + x0 + px) * C + cg * 8) =
            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
      }
    }
  }
  // ---- level 2: 2x2 blocks (all four pixels are "centre"), 4 x 16 cells
  {
    const int Wc = W / 2;
    __half* ob = g2 + (long long)b * (H / 2) * Wc * C + c0;
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int item = i * 256 + threadIdx.x;
      const int ch = item & 31, cell = item >> 5;
      const int cy = cell >> 4, cx = cell & 15;
      if (x0 + 2 * cx + 1 < W) {
        const float* t = &tile[ch][(2 * cy) * 32 + 2 * cx];
        const float v = 0.25f * ((t[0] + t[1]) + (t[32] + t[33]));
        ob[((long long)(y0 / 2 + cy) * Wc + x0 / 2 + cx) * C + ch] = __float2half_rn(v);
      }
    }
  }
  // ---- level 1: 4x4 blocks, centre rows/cols 1,2 ; 2 x 8 cells
  {
    const int Wc = W / 4;
    __half* ob = g1 + (long long)b * (H / 4) * Wc * C + c0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int item = i * 256 + threadIdx.x;
      const int ch = item & 31, cell = item >> 5;
      const int cy = cell >> 3, cx = cell & 7;
      if (x0 + 4 * cx + 3 < W) {
        const float* t = &tile[ch][(4 * cy + 1) * 32 + 4 * cx + 1];
        const float v = 0.25f * ((t[0] + t[1]) + (t[32] + t[33]));
        ob[((long long)(y0 / 4 + cy) * Wc + x0 / 4 + cx) * C + ch] = __float2half_rn(v);
      }
    }
  }
  // ---- level 0: 8x8 blocks, centre rows/cols 3,4 ; 1 x 4 cells
  if (threadIdx.x < 128) {
    const int Wc = W / 8;
    __half* ob = g0 + (long long)b * (H / 8) * Wc * C + c0;
    const int ch = threadIdx.x & 31, cx = threadIdx.x >> 5;
    if (x0 + 8 * cx + 7 < W) {
      const float* t = &tile[ch][3 * 32 + 8 * cx + 3];
      const float v = 0.25f * ((t[0] + t[1]) + (t[32] + t[33]));
      ob[((long long)(y0 / 8) * Wc + x0 / 8 + cx) * C + ch] = __float2half_rn(v);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Query state initialisation (frame_..._decoder.py:73-74, 140): per (group, query) row
//   z32 = query_feat[q]; z16 = fp16(z32); ze16 = fp16(query_feat[q] + query_embed[q]);
//   d = LayerNorm_dec(query_feat[q]) -> d16 / d32.      One warp per row, C = 256.
__global__ void __launch_bounds__(256)
init_queries_kernel(const float* __restrict__ qfeat, const float* __restrict__ qembed, const float* __restrict__ g,
                    const float* __restrict__ bta, float* __restrict__ z32, __half* __restrict__ z16,
                    __half* __restrict__ ze16, float* __restrict__ d32, __half* __restrict__ d16, int Q, int rows) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int q = row % Q;
  float x[8], e[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = __ldg(qfeat + q * 256 + lane + i * 32);
    e[i] = __ldg(qembed + q * 256 + lane + i * 32);
    s += x[i];
  }
  const float mean = warp_sum(s) * (1.f / 256.f);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = x[i] - mean; sq += d * d; }
  const float rstd = rsqrtf(warp_sum(sq) * (1.f / 256.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = lane + i * 32;
    const long long o = (long long)row * 256 + c;
    const float d = (x[i] - mean) * rstd * __ldg(g + c) + __ldg(bta + c);
    z32[o] = x[i];
    z16[o] = __float2half_rn(x[i]);
    ze16[o] = __float2half_rn(x[i] + e[i]);
    d32[o] = d;
    d16[o] = __float2half_rn(d);
  }
}

// Tail of the split linear+LayerNorm path used when a projection has too few rows to fill the GPU with row tiles
// (the Video decoders' 100..400 query rows): z = sum_s part[s][row] + bias + resid[row] -> LayerNorm (ln1) -> y
// [-> second LayerNorm (ln2) -> d].  Same outputs as the fused GEMM epilogue (EPI_LN).  One warp per row, C = 256:
// lane l owns columns 4l..4l+3 and 128+4l..128+4l+3 (two coalesced float4 per operand).
struct LnReduceArgs {
  const float* part;      // [S][part_stride rows][256] fp32 partial products (split-K slices)
  int S;
  long long part_stride;  // rows between slices
  const float* bias;      // [256]
  const float* resid;     // [rows][256]
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;   // null -> no second norm
  const float* pe; int pe_period;
  float* y32; __half* y16; __half* ype16;
  float* d32; __half* d16;
  int rows;
};

// one warp finishes one row.  COHERENT: partials / residual are read through L2 (ld.global.cg) -- for callers that run in
// the same launch as the producers of those rows (the wide query-side chain, csrc/chain.cuh)
template <bool COHERENT>
__device__ __forceinline__ void ln_reduce_row(const LnReduceArgs& a, const int row, const int lane) {
  const int c0 = lane * 4, c1 = 128 + lane * 4;
  auto ld4 = [](const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
  auto ld4c = [](const float* p) {
    return COHERENT ? __ldcg(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
  };
  float z[8];
  {
    const float4 b0 = ld4(a.bias + c0), b1 = ld4(a.bias + c1);
    const float4 r0 = ld4c(a.resid + (long long)row * 256 + c0), r1 = ld4c(a.resid + (long long)row * 256 + c1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < a.S; ++s) {
      const float* p = a.part + ((long long)s * a.part_stride + row) * 256;
      const float4 p0 = ld4c(p + c0), p1 = ld4c(p + c1);
      acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w;
      acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
    }
    z[0] = acc[0] + b0.x + r0.x; z[1] = acc[1] + b0.y + r0.y; z[2] = acc[2] + b0.z + r0.z; z[3] = acc[3] + b0.w + r0.w;
    z[4] = acc[4] + b1.x + r1.x; z[5] = acc[5] + b1.y + r1.y; z[6] = acc[6] + b1.z + r1.z; z[7] = acc[7] + b1.w + r1.w;
  }
  auto layer_norm = [&](float (&v)[8], const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.f / 256.f);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; sq += d * d; }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / 256.f) + 1e-5f);
    const float4 g0 = ld4(g + c0), g1 = ld4(g + c1), b0 = ld4(b + c0), b1 = ld4(b + c1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
  };
  auto st32 = [&](float* base, const float (&v)[8]) {
    *reinterpret_cast<float4*>(base + (long long)row * 256 + c0) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(base + (long long)row * 256 + c1) = make_float4(v[4], v[5], v[6], v[7]);
  };
  auto st16 = [&](__half* base, const float (&v)[8]) {
    *reinterpret_cast<uint2*>(base + (long long)row * 256 + c0) = make_uint2(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]));
    *reinterpret_cast<uint2*>(base + (long long)row * 256 + c1) = make_uint2(pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
  };
  layer_norm(z, a.ln1_g, a.ln1_b);
  if (a.y32) st32(a.y32, z);
  if (a.y16) st16(a.y16, z);
  if (a.ype16 && a.pe) {
    const float* pe = a.pe + (long long)(row % a.pe_period) * 256;
    const float4 p0 = ld4(pe + c0), p1 = ld4(pe + c1);
    const float e[8] = {z[0] + p0.x, z[1] + p0.y, z[2] + p0.z, z[3] + p0.w, z[4] + p1.x, z[5] + p1.y, z[6] + p1.z, z[7] + p1.w};
    st16(a.ype16, e);
  }
  if (a.ln2_g) {
    layer_norm(z, a.ln2_g, a.ln2_b);
    if (a.d32) st32(a.d32, z);
    if (a.d16) st16(a.d16, z);
  }
}

__global__ void __launch_bounds__(256)
ln_reduce_kernel(const LnReduceArgs a) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  ln_reduce_row<false>(a, row, threadIdx.x & 31);
}

// Row-wise LayerNorm (optional) + L2 normalisation (optional) to fp16 / fp32, one warp per row.
//   mode bit0: LayerNorm with (g, b), eps 1e-5        (SideAdapter ln_post, side_adapter.py:203)
//   mode bit1: divide by the L2 norm                  (ClipAdapter.normalize adapter.py:118-119; F.normalize side_adapter.py:205)
//   mode bit2: do NOT divide; write the row's sum of squares to ss[row] instead (the logits GEMM's epilogue divides:
//              ovis_linear_rowscale_f16).  ss_zero (optional, [rows]): set to 0 (accumulator of a following GEMM epilogue).
__global__ void __launch_bounds__(256)
rownorm_kernel(const float* __restrict__ in, const float* __restrict__ g, const float* __restrict__ bta,
               float* __restrict__ out32, __half* __restrict__ out16, int rows, int D, int mode, float* __restrict__ ss,
               float* __restrict__ ss_zero, int group_rows, int group_stride) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // input rows may be the first group_rows rows of consecutive groups of group_stride rows (the SOS tokens of every frame
  // inside the [SOS | CLS | patches] token matrix); outputs are dense
  const long long in_row = group_rows > 0 ? (long long)(row / group_rows) * group_stride + row % group_rows : row;
  const float* x = in + in_row * D;
  float mean = 0.f, rstd = 1.f;
  if (mode & 1) {
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += x[c];
    mean = warp_sum(s) / D;
    float sq = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = x[c] - mean; sq += d * d; }
    rstd = rsqrtf(warp_sum(sq) / D + 1e-5f);
  }
  float inv = 1.f;
  if (ss_zero && lane == 0) ss_zero[row] = 0.f;
  if (mode & 4) {
    float sq = 0.f;
    for (int c = lane; c < D; c += 32) {
      float v = x[c];
      if (mode & 1) v = (v - mean) * rstd * g[c] + bta[c];
      v = __half2float(__float2half_rn(v));           // the norm of what the GEMM will actually multiply
      sq += v * v;
    }
    sq = warp_sum(sq);
    if (lane == 0 && ss) ss[row] = sq;
  } else if (mode & 2) {
    float sq = 0.f;
    for (int c = lane; c < D; c += 32) {
      float v = x[c];
      if (mode & 1) v = (v - mean) * rstd * g[c] + bta[c];
      sq += v * v;
    }
    // F.normalize clamps the norm at 1e-12; x / x.norm() does not - identical for non-degenerate rows
    inv = 1.f / fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
  }
  for (int c = lane; c < D; c += 32) {
    float v = x[c];
    if (mode & 1) v = (v - mean) * rstd * g[c] + bta[c];
    v *= inv;
    if (out32) out32[(long long)row * D + c] = v;
    if (out16) out16[(long long)row * D + c] = __float2half_rn(v);
  }
}

// fp32 -> fp16 cast (weights / small activations)
__global__ void cast_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

// OpenVIS.open_vocabulary_inference tail (openvis/openvis.py:123-141): per query, mean of the CLIP logits over the
// frames whose mask is non-empty, then softmax over the vocabulary.  logits [T][Q][K], valid [T][Q] -> probs [Q][K]
// (rows of queries with no valid frame are zero and flagged in qvalid).  One CTA per query.
__global__ void __launch_bounds__(256)
clip_aggregate_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ valid,
                      float* __restrict__ probs, unsigned char* __restrict__ qvalid, int T, int Q, int K) {
  extern __shared__ float sm_acc[];    // [K]
  __shared__ float red[32];
  __shared__ int s_cnt;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) {
    int c = 0;
    for (int t = 0; t < T; ++t) c += valid[t * Q + q] ? 1 : 0;
    s_cnt = c;
  }
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt == 0) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) probs[(long long)q * K + k] = 0.f;
    if (threadIdx.x == 0) qvalid[q] = 0;
    return;
  }
  float lmax = -INFINITY;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t)
      if (valid[t * Q + q]) s += logits[((long long)t * Q + q) * K + k];
    s /= cnt;
    sm_acc[k] = s;
    lmax = fmaxf(lmax, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lmax;
  __syncthreads();
  lmax = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) lmax = fmaxf(lmax, red[i]);
  __syncthreads();
  float lsum = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float e = __expf(sm_acc[k] - lmax);
    sm_acc[k] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
  const float inv = 1.f / tot;
  for (int k = threadIdx.x; k < K; k += blockDim.x) probs[(long long)q * K + k] = sm_acc[k] * inv;
  if (threadIdx.x == 0) qvalid[q] = 1;
}

// SideAdapter._build_attn_biases (side_adapter.py:237-270): adaptive max-pool of the per-head attention biases to the
// CLIP grid fused with the construction of the additive [Q+1+L]^2 matrix (L = gh*gw).
//   bias [B][n][Q][h][w] fp32  ->  out [B*n][Q+1+L][Q+1+L] fp32
// One CTA per (row of the matrix, b*n).
__global__ void __launch_bounds__(256)
san_attn_bias_kernel(const float* __restrict__ bias, float* __restrict__ out, int Q, int h, int w, int gh, int gw) {
  const int L = gh * gw, n = Q + 1 + L;
  const int row = blockIdx.x;
  const long long bn = blockIdx.y;
  float* orow = out + (bn * n + row) * n;
  for (int col = threadIdx.x; col < n; col += blockDim.x) {
    float v;
    if (col < Q) v = (row < Q && col == row) ? 0.f : -100.f;
    else if (col == Q) v = row < Q ? -100.f : 0.f;
    else if (row >= Q) v = 0.f;
    else {
      const int cell = col - Q - 1;
      const int gy = cell / gw, gx = cell % gw;
      const int y0 = (gy * h) / gh, y1 = ((gy + 1) * h + gh - 1) / gh;
      const int x0 = (gx * w) / gw, x1 = ((gx + 1) * w + gw - 1) / gw;
      const float* p = bias + (bn * Q + row) * h * w;
      float m = -INFINITY;
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) m = fmaxf(m, __ldg(p + y * w + x));
      v = m;
    }
    orow[col] = v;
  }
}

}  // namespace ovis
