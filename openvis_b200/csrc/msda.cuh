// Multi-scale deformable attention, forward: the reference's only native op,
// ms_deformable_im2col_gpu_kernel (openvis/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:243-305),
// called by MSDeformAttnFunction.forward (ops/functions/ms_deform_attn_func.py:32-40).
//   out[b][q][m][c] = sum_{l, p} w[b][q][m][l][p] * bilinear(value_l[b][:, m, c], loc[b][q][m][l][p])
// with h_im = loc_y * H_l - 0.5, w_im = loc_x * W_l - 0.5, zero contribution from corners outside the map and from
// samples outside (-1, H) x (-1, W).
//
// The reference maps one thread to one output channel and re-reads the sampling location / weight in every thread.
// Here a group of D/4 lanes owns one (b, q, head): every lane gathers float4 channel quads, so one bilinear corner of a
// head is a single contiguous 16 * D/4-byte read (128 B for D = 32) issued by D/4 lanes instead of D, and the L * P
// (location, weight) triples are read once per group and broadcast with shuffles.  The value map of a frame (20 MB at
// 736 x 1280) lives in L2, so the op is gather-bound there.
#pragma once
#include <cuda_fp16.h>
#include "ptx.cuh"

namespace ovis {

struct MsdaArgs {
  const float* value;        // [N][S][M][D]
  const long long* shapes;   // [L][2] (H, W)  -- int64 like the reference
  const long long* start;    // [L] level start index
  const float* loc;          // [N][Lq][M][L][P][2] (x, y) in [0, 1]
  const float* weight;       // [N][Lq][M][L][P]
  float* out;                // [N][Lq][M*D]
  int N, S, M, D, Lq, L, P;
};

// LPG lanes per (b, q, m) group, 4 channels per lane: D == 4 * LPG
template <int LPG>
__global__ void __launch_bounds__(256)
msda_forward_vec_kernel(const MsdaArgs a) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPG;                                    // channel quad within the head
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
  const long long groups = (long long)a.N * a.Lq * a.M;
  const bool ok = group < groups;
  const long long gi = ok ? group : groups - 1;
  const int m = (int)(gi % a.M);
  const long long bq = gi / a.M;
  const int b = (int)(bq / a.Lq);
  const int LP = a.L * a.P;
  const float* locp = a.loc + gi * LP * 2;
  const float* wp = a.weight + gi * LP;
  const int row_stride = a.M * a.D;                              // floats between consecutive spatial positions
  const float* vb = a.value + (long long)b * a.S * row_stride + m * a.D + sub * 4;
  const unsigned gmask = 0xffffffffu;
  const int gbase = lane - sub;                                  // first lane of this group
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = 0; i0 < LP; i0 += LPG) {
    // each lane of the group fetches one (x, y, w) triple; they are handed round with shuffles
    const int mine = i0 + sub;
    float lx = 0.f, ly = 0.f, lw = 0.f;
    if (mine < LP) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(locp) + mine);
      lx = xy.x; ly = xy.y; lw = __ldg(wp + mine);
    }
    const int cnt = min(LPG, LP - i0);
    for (int j = 0; j < cnt; ++j) {
      const float x = __shfl_sync(gmask, lx, gbase + j);
      const float y = __shfl_sync(gmask, ly, gbase + j);
      const float w = __shfl_sync(gmask, lw, gbase + j);
      const int l = (i0 + j) / a.P;
      const int H = (int)__ldg(a.shapes + 2 * l), W = (int)__ldg(a.shapes + 2 * l + 1);
      const float h_im = y * H - 0.5f, w_im = x * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
        const float lh = h_im - h_low, lw2 = w_im - w_low;
        const float hh = 1.f - lh, hw = 1.f - lw2;
        const float* vl = vb + (long long)__ldg(a.start + l) * row_stride;
        float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
        if (h_low >= 0 && w_low >= 0) v1 = __ldg(reinterpret_cast<const float4*>(vl + ((long long)h_low * W + w_low) * row_stride));
        if (h_low >= 0 && w_low + 1 <= W - 1) v2 = __ldg(reinterpret_cast<const float4*>(vl + ((long long)h_low * W + w_low + 1) * row_stride));
        if (h_low + 1 <= H - 1 && w_low >= 0) v3 = __ldg(reinterpret_cast<const float4*>(vl + ((long long)(h_low + 1) * W + w_low) * row_stride));
        if (h_low + 1 <= H - 1 && w_low + 1 <= W - 1) v4 = __ldg(reinterpret_cast<const float4*>(vl + ((long long)(h_low + 1) * W + w_low + 1) * row_stride));
        const float w1 = hh * hw, w2 = hh * lw2, w3 = lh * hw, w4 = lh * lw2;       // (ms_deform_im2col_cuda.cuh:60-62)
        acc.x += w * (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x);
        acc.y += w * (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y);
        acc.z += w * (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z);
        acc.w += w * (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w);
      }
    }
  }
  if (ok) *reinterpret_cast<float4*>(a.out + gi * a.D + sub * 4) = acc;
}

// Any D: one thread per output channel, the reference's mapping (used for channel counts that are not 4 * 2^k <= 128).
__global__ void __launch_bounds__(256)
msda_forward_scalar_kernel(const MsdaArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.N * a.Lq * a.M * a.D;
  if (idx >= total) return;
  const int c = (int)(idx % a.D);
  const long long gi = idx / a.D;
  const int m = (int)(gi % a.M);
  const int b = (int)(gi / a.M / a.Lq);
  const int LP = a.L * a.P;
  const int row_stride = a.M * a.D;
  const float* vb = a.value + (long long)b * a.S * row_stride + m * a.D + c;
  float acc = 0.f;
  for (int i = 0; i < LP; ++i) {
    const int l = i / a.P;
    const int H = (int)a.shapes[2 * l], W = (int)a.shapes[2 * l + 1];
    const float x = a.loc[(gi * LP + i) * 2], y = a.loc[(gi * LP + i) * 2 + 1], w = a.weight[gi * LP + i];
    const float h_im = y * H - 0.5f, w_im = x * W - 0.5f;
    if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
      const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
      const float lh = h_im - h_low, lw2 = w_im - w_low, hh = 1.f - lh, hw = 1.f - lw2;
      const float* vl = vb + (long long)a.start[l] * row_stride;
      float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
      if (h_low >= 0 && w_low >= 0) v1 = vl[((long long)h_low * W + w_low) * row_stride];
      if (h_low >= 0 && w_low + 1 <= W - 1) v2 = vl[((long long)h_low * W + w_low + 1) * row_stride];
      if (h_low + 1 <= H - 1 && w_low >= 0) v3 = vl[((long long)(h_low + 1) * W + w_low) * row_stride];
      if (h_low + 1 <= H - 1 && w_low + 1 <= W - 1) v4 = vl[((long long)(h_low + 1) * W + w_low + 1) * row_stride];
      acc += w * (hh * hw * v1 + hh * lw2 * v2 + lh * hw * v3 + lh * lw2 * v4);
    }
  }
  a.out[idx] = acc;
}

// MSDeformAttn.forward between the query projections and the sampling op (ops/modules/ms_deform_attn.py:104-117):
// proj [rows][M * L * P * 3] fp32 = one GEMM's output, columns [0, 2 * MLP) the sampling offsets laid out (m, l, p, xy),
// columns [2 * MLP, 3 * MLP) the attention logits (m, l, p).  One thread per (row, head): softmax over the L * P logits,
// sampling_locations = reference_points + offsets / (W_l, H_l)   (2-d reference points), or
//                    = reference_points[:2] + offsets / P * reference_points[2:] * 0.5   (reference boxes).
struct MsdaPrepArgs {
  const float* proj;
  const float* ref;          // [rows][L][ref_dim]
  const long long* shapes;   // [L][2] (H, W)
  float* loc;                // [rows][M][L][P][2]
  float* weight;             // [rows][M][L][P]
  long long rows;
  int M, L, P, ref_dim;
};

__global__ void __launch_bounds__(256)
msda_prepare_kernel(const MsdaPrepArgs a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.rows * a.M) return;
  const int m = (int)(idx % a.M);
  const long long row = idx / a.M;
  const int LP = a.L * a.P, MLP = a.M * LP;
  const float* off = a.proj + row * 3ll * MLP + (long long)m * LP * 2;
  const float* lg = a.proj + row * 3ll * MLP + 2ll * MLP + (long long)m * LP;
  float mx = -INFINITY;
  for (int i = 0; i < LP; ++i) mx = fmaxf(mx, lg[i]);
  float sum = 0.f;
  for (int i = 0; i < LP; ++i) sum += expf(lg[i] - mx);
  const float inv = 1.f / sum;
  float* wo = a.weight + (row * a.M + m) * LP;
  float* lo = a.loc + (row * a.M + m) * LP * 2;
  for (int l = 0; l < a.L; ++l) {
    const float* r = a.ref + (row * a.L + l) * a.ref_dim;
    const float Hl = (float)a.shapes[2 * l], Wl = (float)a.shapes[2 * l + 1];
    for (int p = 0; p < a.P; ++p) {
      const int i = l * a.P + p;
      wo[i] = expf(lg[i] - mx) * inv;
      const float ox = off[2 * i], oy = off[2 * i + 1];
      if (a.ref_dim == 2) {
        lo[2 * i] = r[0] + ox / Wl;
        lo[2 * i + 1] = r[1] + oy / Hl;
      } else {
        lo[2 * i] = r[0] + ox / (float)a.P * r[2] * 0.5f;
        lo[2 * i + 1] = r[1] + oy / (float)a.P * r[3] * 0.5f;
      }
    }
  }
}

// MSDeformAttn.forward from the two projection GEMMs to the operand of output_proj in ONE kernel (the module path of the
// pixel decoder's encoder, ops/modules/ms_deform_attn.py:98-122): 4 lanes own a (row, head); they softmax the head's L * P
// logits and form the sampling locations in registers (what msda_prepare_kernel writes to memory), then gather the fp16
// value map as 16-byte channel octets (one bilinear corner of a head = one contiguous 64-byte read) with fp32 weights and
// accumulators, and store the fp16 rows the output projection consumes.  D = 32 channels per head, L * P <= 16.
struct MsdaFusedArgs {
  const __half* value;       // [N][S][M][32] fp16 (value_proj output)
  const float* proj;         // [N * Lq][3 * M * L * P] fp32: offsets (m, l, p, xy) | logits (m, l, p)
  const float* ref;          // [N or 1][Lq][L][ref_dim]
  long long ref_bs;          // floats between the reference points of consecutive samples (0: shared)
  const long long* shapes;   // [L][2] (H, W)
  const long long* start;    // [L]
  __half* out;               // [N * Lq][M * 32] fp16
  int N, S, M, Lq, L, P, ref_dim;
};

__global__ void __launch_bounds__(256)
msda_fused_kernel(const MsdaFusedArgs a) {
  constexpr int LPG = 4, D = 32;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LPG - 1);
  const int gbase = lane - sub;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
  const long long groups = (long long)a.N * a.Lq * a.M;
  const bool ok = group < groups;
  const long long gi = ok ? group : groups - 1;
  const int m = (int)(gi % a.M);
  const long long row = gi / a.M;                 // (sample, query)
  const int b = (int)(row / a.Lq);
  const long long q = row - (long long)b * a.Lq;
  const int LP = a.L * a.P, MLP = a.M * LP;
  const float* off = a.proj + row * 3ll * MLP + (long long)m * LP * 2;
  const float* lg = a.proj + row * 3ll * MLP + 2ll * MLP + (long long)m * LP;
  const float* rp = a.ref + (long long)b * a.ref_bs + q * a.L * a.ref_dim;
  // point i lives in lane i % 4, slot i / 4
  float lw[4], lx[4], ly[4];
  float mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int i = s * LPG + sub;
    lw[s] = -INFINITY; lx[s] = 0.f; ly[s] = 0.f;
    if (i < LP) {
      lw[s] = __ldg(lg + i);
      const float2 o = __ldg(reinterpret_cast<const float2*>(off) + i);
      const int l = i / a.P;
      const float* r = rp + l * a.ref_dim;
      if (a.ref_dim == 2) {
        lx[s] = __ldg(r) + o.x / (float)__ldg(a.shapes + 2 * l + 1);
        ly[s] = __ldg(r + 1) + o.y / (float)__ldg(a.shapes + 2 * l);
      } else {
        lx[s] = __ldg(r) + o.x / (float)a.P * __ldg(r + 2) * 0.5f;
        ly[s] = __ldg(r + 1) + o.y / (float)a.P * __ldg(r + 3) * 0.5f;
      }
      mx = fmaxf(mx, lw[s]);
    }
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    lw[s] = (s * LPG + sub < LP) ? expf(lw[s] - mx) : 0.f;
    sum += lw[s];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.f / sum;
  const int row_stride = a.M * D;                 // halves between consecutive positions
  const __half* vb = a.value + (long long)b * a.S * row_stride + m * D + sub * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    if (s * LPG >= LP) break;
#pragma unroll
    for (int j = 0; j < LPG; ++j) {
      const float x = __shfl_sync(0xffffffffu, lx[s], gbase + j);
      const float y = __shfl_sync(0xffffffffu, ly[s], gbase + j);
      const float w = __shfl_sync(0xffffffffu, lw[s], gbase + j) * inv;
      const int i = s * LPG + j;
      if (i >= LP) continue;
      const int l = i / a.P;
      const int H = (int)__ldg(a.shapes + 2 * l), W = (int)__ldg(a.shapes + 2 * l + 1);
      const float h_im = y * H - 0.5f, w_im = x * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
        const float lh = h_im - h_low, lw2 = w_im - w_low;
        const float hh = 1.f - lh, hw = 1.f - lw2;
        const __half* vl = vb + (long long)__ldg(a.start + l) * row_stride;
        uint4 c[4];
        float cw[4] = {w * hh * hw, w * hh * lw2, w * lh * hw, w * lh * lw2};
        const bool in[4] = {h_low >= 0 && w_low >= 0, h_low >= 0 && w_low + 1 <= W - 1, h_low + 1 <= H - 1 && w_low >= 0,
                            h_low + 1 <= H - 1 && w_low + 1 <= W - 1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          c[k] = make_uint4(0u, 0u, 0u, 0u);
          if (in[k]) c[k] = __ldg(reinterpret_cast<const uint4*>(vl + ((long long)(h_low + (k >> 1)) * W + w_low + (k & 1)) * row_stride));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2* hp = reinterpret_cast<const __half2*>(&c[k]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            acc[2 * e] += cw[k] * f.x;
            acc[2 * e + 1] += cw[k] * f.y;
          }
        }
      }
    }
  }
  if (ok) {
    const __half2 h0 = __floats2half2_rn(acc[0], acc[1]), h1 = __floats2half2_rn(acc[2], acc[3]);
    const __half2 h2 = __floats2half2_rn(acc[4], acc[5]), h3 = __floats2half2_rn(acc[6], acc[7]);
    uint4 u;
    u.x = *reinterpret_cast<const unsigned*>(&h0);
    u.y = *reinterpret_cast<const unsigned*>(&h1);
    u.z = *reinterpret_cast<const unsigned*>(&h2);
    u.w = *reinterpret_cast<const unsigned*>(&h3);
    *reinterpret_cast<uint4*>(a.out + gi * D + sub * 8) = u;
  }
}

}  // namespace ovis
