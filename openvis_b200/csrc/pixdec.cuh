// Pixel-decoder glue around the tcgen05 GEMMs and the deformable-attention kernel:
// GroupNorm(32, 256) on token-major maps, bilinear top-down addition, the 3x3 convolution's operand unfold, and the
// token-major -> NCHW store of the maps the reference API returns.
// Reference: openvis/modeling/pixel_decoder/msdeformattn.py:227-236 (input_proj = 1x1 conv + GN), :278-299 (lateral /
// output convolutions with GN), :329-380 (forward_features).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ovis {

constexpr int PD_C = 256;        // conv_dim of every shipped config
constexpr int PD_GROUPS = 32;    // GroupNorm(32, conv_dim): 8 channels per group

// Per (sample, group) sum and sum of squares of x [B][S][256] fp32 -> stats [B][32][2] (double, zeroed by the caller).
// Thread = (row slot, channel quad); both quads of a group sit in neighbouring lanes.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int S, int rows_per_cta) {
  __shared__ float red[4][PD_GROUPS][2];
  const int b = blockIdx.y;
  const int q = threadIdx.x & 63, slot = threadIdx.x >> 6;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(S, r_begin + rows_per_cta);
  const float4* base = reinterpret_cast<const float4*>(x + (long long)b * S * PD_C) + q;
  float s = 0.f, ss = 0.f;
#pragma unroll 4
  for (int r = r_begin + slot; r < r_end; r += 4) {
    const float4 v = __ldg(base + (long long)r * (PD_C / 4));
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  if (!(q & 1)) {
    red[slot][q >> 1][0] = s;
    red[slot][q >> 1][1] = ss;
  }
  __syncthreads();
  if (threadIdx.x < 2 * PD_GROUPS) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    const double v = (double)red[0][g][k] + (double)red[1][g][k] + (double)red[2][g][k] + (double)red[3][g][k];
    atomicAdd(stats + ((long long)b * PD_GROUPS + g) * 2 + k, v);
  }
}

struct GnApplyArgs {
  const float* x;         // [B][S][256] fp32 (GEMM output)
  const double* stats;    // [B][32][2]
  const float* gamma;
  const float* beta;
  float eps;
  int S, W;               // S = H * W positions per sample; W needed only with `add`
  int relu;
  // optional term added AFTER the normalisation: a map of hs x ws positions, bilinearly resized to the S positions
  // (align_corners = False; identity when the sizes agree).  Element (b, c, p) at add[b * add_bs + c * add_cs + p * add_ps].
  const float* add;
  long long add_bs, add_cs, add_ps;
  int hs, ws;
  // outputs: row (b, r) lands at b * out_bs + out_off + r (rows of 256)
  float* out32;
  __half* out16;
  long long out_bs, out_off;
};

// Thread = (position, group of 8 channels).
__global__ void __launch_bounds__(256)
gn_apply_kernel(const GnApplyArgs a) {
  const int b = blockIdx.y;
  const int g = threadIdx.x & 31;
  const double cnt = (double)a.S * (PD_C / PD_GROUPS);
  const double m = a.stats[((long long)b * PD_GROUPS + g) * 2] / cnt;
  double var = a.stats[((long long)b * PD_GROUPS + g) * 2 + 1] / cnt - m * m;
  if (var < 0.0) var = 0.0;
  const float mean = (float)m, rstd = (float)(1.0 / sqrt(var + (double)a.eps));
  float ga[8], be[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + g * 8)), g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + g * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(a.beta + g * 8) + 1);
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ga[e] *= rstd;
      be[e] -= mean * ga[e];
    }
  }
  const float sy = a.add ? (float)a.hs / (float)(a.S / a.W) : 0.f, sx = a.add ? (float)a.ws / (float)a.W : 0.f;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < a.S; r += gridDim.x * 8) {
    const float4* xp = reinterpret_cast<const float4*>(a.x + ((long long)b * a.S + r) * PD_C + g * 8);
    const float4 v0 = __ldg(xp), v1 = __ldg(xp + 1);
    float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = v[e] * ga[e] + be[e];
    if (a.add) {
      // F.interpolate(mode="bilinear", align_corners=False): source index = scale * (dst + 0.5) - 0.5, clamped at 0
      const int Y = r / a.W, X = r - Y * a.W;
      const float fy = fmaxf(sy * ((float)Y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * ((float)X + 0.5f) - 0.5f, 0.f);
      const int y0 = min((int)fy, a.hs - 1), x0 = min((int)fx, a.ws - 1);
      const int y1 = min(y0 + 1, a.hs - 1), x1 = min(x0 + 1, a.ws - 1);
      const float ly = fy - (float)y0, lx = fx - (float)x0;
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      const float* ab = a.add + (long long)b * a.add_bs + (long long)(g * 8) * a.add_cs;
      const long long p00 = (long long)(y0 * a.ws + x0) * a.add_ps, p01 = (long long)(y0 * a.ws + x1) * a.add_ps;
      const long long p10 = (long long)(y1 * a.ws + x0) * a.add_ps, p11 = (long long)(y1 * a.ws + x1) * a.add_ps;
      if (a.add_cs == 1) {
        const float4* q00 = reinterpret_cast<const float4*>(ab + p00);
        const float4* q01 = reinterpret_cast<const float4*>(ab + p01);
        const float4* q10 = reinterpret_cast<const float4*>(ab + p10);
        const float4* q11 = reinterpret_cast<const float4*>(ab + p11);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 c00 = __ldg(q00 + h), c01 = __ldg(q01 + h), c10 = __ldg(q10 + h), c11 = __ldg(q11 + h);
          v[h * 4 + 0] += w00 * c00.x + w01 * c01.x + w10 * c10.x + w11 * c11.x;
          v[h * 4 + 1] += w00 * c00.y + w01 * c01.y + w10 * c10.y + w11 * c11.y;
          v[h * 4 + 2] += w00 * c00.z + w01 * c01.z + w10 * c10.z + w11 * c11.z;
          v[h * 4 + 3] += w00 * c00.w + w01 * c01.w + w10 * c10.w + w11 * c11.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float* ac = ab + (long long)e * a.add_cs;
          v[e] += w00 * __ldg(ac + p00) + w01 * __ldg(ac + p01) + w10 * __ldg(ac + p10) + w11 * __ldg(ac + p11);
        }
      }
    }
    if (a.relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    const long long orow = (long long)b * a.out_bs + a.out_off + r;
    if (a.out32) {
      float4* o = reinterpret_cast<float4*>(a.out32 + orow * PD_C + g * 8);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (a.out16) {
      const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
      const __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
      uint4 u;
      u.x = *reinterpret_cast<const unsigned*>(&h0);
      u.y = *reinterpret_cast<const unsigned*>(&h1);
      u.z = *reinterpret_cast<const unsigned*>(&h2);
      u.w = *reinterpret_cast<const unsigned*>(&h3);
      *reinterpret_cast<uint4*>(a.out16 + orow * PD_C + g * 8) = u;
    }
  }
}

// in [B][in_bs rows][C] fp32 (rows in_off .. in_off + N of every sample) -> out [B][C][N] fp32 (NCHW).  32 x 32 tiles.
__global__ void __launch_bounds__(256)
tokens_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int N, long long in_bs, long long in_off) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const float* ib = in + ((long long)b * in_bs + in_off) * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty + i * 8, c = c0 + tx;
    tile[ty + i * 8][tx] = (n < N && c < C) ? __ldg(ib + (long long)n * C + c) : 0.f;
  }
  __syncthreads();
  float* ob = out + (long long)b * C * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, n = n0 + tx;
    if (c < C && n < N) ob[(long long)c * N + n] = tile[tx][ty + i * 8];
  }
}

// A-operand of a 3x3 / padding 1 convolution as a GEMM: in [B][H][W][C] fp16 -> out [B*H*W][9*C] fp16, tap-major
// (ky, kx, c), zeros outside the map.  One thread per (position, tap, 8 channels) = one 16-byte copy.
__global__ void __launch_bounds__(256)
conv3x3_unfold_kernel(const __half* __restrict__ in, __half* __restrict__ out, int H, int W, int C8, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  long long t = i / C8;
  const int tap = (int)(t % 9);
  t /= 9;                                   // position index (b, y, x)
  const int x = (int)(t % W);
  const long long by = t / W;
  const int y = (int)(by % H);
  const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (yy >= 0 && yy < H && xx >= 0 && xx < W)
    v = __ldg(reinterpret_cast<const uint4*>(in) + ((by - y + yy) * W + xx) * C8 + c8);
  reinterpret_cast<uint4*>(out)[i] = v;
}

// ---- token-major hand-off from the pixel decoder to the masked decoder (no NCHW fp32 round trip) ------------------------
// g_l of the decoder's mask head (bilinear down-sampling by an integer factor = mean of the centre 2x2
// pixels of each s x s block) from the token-major fp16 mask features: ft [B][H][W][256] -> g [B][H/s][W/s][256], fp32 mean.
__global__ void __launch_bounds__(256)
tokens_pool_kernel(const __half* __restrict__ ft, __half* __restrict__ g, int H, int W, int s, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, y, x, 8-channel group)
  if (i >= total) return;
  const int c8 = (int)(i & 31);
  long long t = i >> 5;
  const int Ws = W / s, Hs = H / s;
  const int x = (int)(t % Ws);
  t /= Ws;
  const int y = (int)(t % Hs);
  const long long b = t / Hs;
  const int y0 = y * s + s / 2 - 1, x0 = x * s + s / 2 - 1;
  const uint4* base = reinterpret_cast<const uint4*>(ft) + ((b * H + y0) * (long long)W + x0) * 32 + c8;
  const uint4 v00 = __ldg(base), v01 = __ldg(base + 32), v10 = __ldg(base + (long long)W * 32), v11 = __ldg(base + (long long)W * 32 + 32);
  uint4 o;
  const __half2* a = reinterpret_cast<const __half2*>(&v00);
  const __half2* bq = reinterpret_cast<const __half2*>(&v01);
  const __half2* c = reinterpret_cast<const __half2*>(&v10);
  const __half2* d = reinterpret_cast<const __half2*>(&v11);
  __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fa = __half22float2(a[e]), fb = __half22float2(bq[e]), fc = __half22float2(c[e]), fd = __half22float2(d[e]);
    po[e] = __floats2half2_rn(0.25f * ((fa.x + fb.x) + (fc.x + fd.x)), 0.25f * ((fa.y + fb.y) + (fc.y + fd.y)));
  }
  reinterpret_cast<uint4*>(g)[i] = o;
}

// key operand of the decoder's cross-attention from token-major fp16 features: xp[b][n][:] = fp16(xt[b][n][:] + pos[n][:] + pos_t[b][:])
// (level embedding + 2-D sine position per token, frame term of the 3-D embedding per frame; pos_t may be null)
__global__ void __launch_bounds__(256)
tokens_add_pos_kernel(const __half* __restrict__ xt, const float* __restrict__ pos, const float* __restrict__ pos_t,
                      __half* __restrict__ xp, int N, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, n, 8-channel group)
  if (i >= total) return;
  const int c8 = (int)(i & 31);
  const long long bn = i >> 5;
  const int n = (int)(bn % N);
  const long long b = bn / N;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(xt) + i);
  const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + (long long)n * 256 + c8 * 8));
  const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + (long long)n * 256 + c8 * 8) + 1);
  float p[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
  if (pos_t) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(pos_t + b * 256 + c8 * 8));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(pos_t + b * 256 + c8 * 8) + 1);
    p[0] += t0.x; p[1] += t0.y; p[2] += t0.z; p[3] += t0.w; p[4] += t1.x; p[5] += t1.y; p[6] += t1.z; p[7] += t1.w;
  }
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  uint4 o;
  __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(h[e]);
    po[e] = __floats2half2_rn(f.x + p[2 * e], f.y + p[2 * e + 1]);
  }
  reinterpret_cast<uint4*>(xp)[i] = o;
}

}  // namespace ovis
