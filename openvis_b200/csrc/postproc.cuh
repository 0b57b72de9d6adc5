// Device-side post-processing of the decoder outputs: what VideoMaskFormer.postprocess +
// inference_video do on the host side of the reference (openvis/modeling/video_maskformer.py:215-229, 262-298;
// OpenVIS.forward openvis.py:87-96): bilinear x4 up-sampling of the stride-4 mask logits to the padded input size,
// top-10 (query, class) selection, crop to the un-padded image, bilinear resize to the output size, `> 0`.
// Here the two interpolations are composed per output pixel for the SELECTED queries only, thresholded and bit-packed:
// the [Q, T, Hp, Wp] fp32 intermediate (13 GB for cfg 2) is never written and the device-to-host copy shrinks from one
// byte per pixel to one bit.
#pragma once
#include "ptx.cuh"

namespace ovis {

// top-k over the flattened [Q*K] score matrix (scores.flatten(0, 1).topk(10, sorted=False), video_maskformer.py:268)
// plus labels, query indices and the per-query entropy (:271).  One CTA; k <= KL <= 32.  Output sorted by score
// (descending); ties resolve to the lower flat index.
// One pass: every thread keeps the KL best of its strided slice in a sorted register list (an element enters only when it
// beats the list's tail, so the insertion network rarely runs); then k rounds of a block-wide arg-max over the list heads,
// the winner popping its head.  (The first version made k passes over all Q*K scores: 556 us at K = 1196.)
template <int KL, int THREADS>
__global__ void __launch_bounds__(THREADS)
topk_scores_kernel(const float* __restrict__ scores, int Q, int K, int k, float* __restrict__ out_scores,
                   int* __restrict__ out_query, int* __restrict__ out_label, float* __restrict__ out_entropy) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float w_val[32];
  __shared__ int w_idx[32];
  const int n = Q * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lv[KL];
  int li[KL];
#pragma unroll
  for (int j = 0; j < KL; ++j) { lv[j] = -INFINITY; li[j] = 0x7fffffff; }
  for (int i = threadIdx.x; i < n; i += THREADS) {
    float cv = __ldg(scores + i);
    if (cv != cv) cv = INFINITY;                  // torch.topk orders NaN as the largest value
    if (cv > lv[KL - 1]) {                        // strict: of equal values the earlier (lower) index stays ahead
      int ci = i;
#pragma unroll
      for (int j = 0; j < KL; ++j) {
        if (cv > lv[j]) {
          const float tv = lv[j]; lv[j] = cv; cv = tv;
          const int ti = li[j]; li[j] = ci; ci = ti;
        }
      }
    }
  }
  for (int r = 0; r < k; ++r) {
    float best = lv[0];
    int bi = li[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { w_val[warp] = best; w_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      float v = lane < THREADS / 32 ? w_val[lane] : -INFINITY;
      int i = lane < THREADS / 32 ? w_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
      }
      if (lane == 0) { s_val[r] = v; s_idx[r] = i; }
    }
    __syncthreads();
    if (li[0] == s_idx[r]) {                      // the owner pops its head
#pragma unroll
      for (int j = 0; j + 1 < KL; ++j) { lv[j] = lv[j + 1]; li[j] = li[j + 1]; }
      lv[KL - 1] = -INFINITY; li[KL - 1] = 0x7fffffff;
    }
  }
  // outputs + entropy of the selected queries' score rows: sum(-s * log s)
  for (int r = warp; r < k; r += THREADS / 32) {
    const int flat = min(s_idx[r], n - 1);        // (k <= n is checked by the host; never index past the matrix)
    const int q = flat / K;
    float e = 0.f;
    for (int c = lane; c < K; c += 32) {
      const float s = __ldg(scores + (long long)q * K + c);
      e += -s * logf(s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) {
      out_scores[r] = __ldg(scores + flat);
      out_query[r] = q;
      out_label[r] = flat - q * K;
      out_entropy[r] = e;
    }
  }
}

struct MaskPostArgs {
  const float* masks;     // [Q][..][h4][w4] stride-4 logits (pred_masks[0]); frame t of query q at q * q_stride + t * h4 * w4
  long long q_stride;     // elements between consecutive queries (T * h4 * w4 when the clip is the whole tensor)
  const int* query;       // [n_sel] selected query per output plane
  uint32_t* bits;         // [n_sel][T][out_h][words] ; bit x%32 of word x/32 = (resized logit > 0)
  int n_sel, T, h4, w4;
  int pad_h, pad_w;       // padded network input size (first interpolation target)
  int img_h, img_w;       // un-padded image size (crop)
  int out_h, out_w;       // output size (second interpolation target)
  int words;              // ceil(out_w / 32)
};

// F.interpolate(mode="bilinear", align_corners=False) source index: scale * (dst + 0.5) - 0.5 clamped at 0
__device__ __forceinline__ void bilin_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * (dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - i0;
}

// Composite weights of the two chained bilinear interpolations along one axis for output index `dst`:
// out = sum_{j<3} w[j] * L[min(base + j, in_size - 1)].  The two intermediate taps are adjacent pixels of the x4
// up-sampled image, so their four source taps fall on at most three consecutive low-resolution samples.
__device__ __forceinline__ void composite_taps(int dst, float s2, int mid_size, float s1, int in_size, int& base, float (&w)[3]) {
  int i0, i1;
  float l;
  bilin_src(dst, s2, mid_size, i0, i1, l);
  int a0, a1, b0, b1;
  float wa, wb;
  bilin_src(i0, s1, in_size, a0, a1, wa);
  bilin_src(i1, s1, in_size, b0, b1, wb);
  base = a0;
  w[0] = w[1] = w[2] = 0.f;
  const float c[4] = {(1.f - l) * (1.f - wa), (1.f - l) * wa, l * (1.f - wb), l * wb};
  const int idx[4] = {0, a1 - a0, b0 - a0, b1 - a0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    w[0] += idx[k] == 0 ? c[k] : 0.f;
    w[1] += idx[k] == 1 ? c[k] : 0.f;
    w[2] += idx[k] >= 2 ? c[k] : 0.f;
  }
}

// thread = one output column of an MP_ROWS-row strip; warp = 32 consecutive x -> one ballot word per output row.
// The horizontal pass H[r] = sum_c wx[c] L[r][c] is kept for the three low-resolution rows the current output row
// needs and slides down with it, so each output pixel costs about one cached load instead of sixteen.  Per output row a
// thread issues ONE shared load (row base + the three vertical weights as a float4), three multiply-adds, the ballot and a
// shared store; the strip's words are staged in shared memory and written as 32-byte row segments (the first version
// looked up four scalars and wrote one 4-byte word per warp and row from 32-row strips: 0.47 ms per 720 x 1280 clip).
constexpr int MP_ROWS = 120;
__global__ void __launch_bounds__(256)
mask_postprocess_kernel(const MaskPostArgs a) {
  const int plane = blockIdx.z;                               // (selected query, frame)
  const int sel = plane / a.T, t = plane - sel * a.T;
  const int ox = blockIdx.x * 256 + threadIdx.x;
  const int oy0 = blockIdx.y * MP_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wx0 = blockIdx.x * 8;                             // first output word of this CTA
  const float s2y = (float)a.img_h / a.out_h, s2x = (float)a.img_w / a.out_w;     // second resize (image -> output)
  const float s1y = (float)a.h4 / a.pad_h, s1x = (float)a.w4 / a.pad_w;           // first resize (stride 4 -> padded)
  // vertical taps of the strip's rows: computed once per CTA
  __shared__ float4 s_tab[MP_ROWS];                           // (row base as int bits, wy0, wy1, wy2)
  __shared__ uint32_t s_out[MP_ROWS][8];
  if (threadIdx.x < MP_ROWS) {
    int nb;
    float wy[3];
    composite_taps(min(oy0 + (int)threadIdx.x, a.out_h - 1), s2y, a.img_h, s1y, a.h4, nb, wy);
    s_tab[threadIdx.x] = make_float4(__int_as_float(nb), wy[0], wy[1], wy[2]);
  }
  __syncthreads();
  const int oy1 = min(oy0 + MP_ROWS, a.out_h);
  if (wx0 + warp < a.words) {                                 // (whole warps beyond the row only help with the final copy)
    const int q = __ldg(a.query + sel);
    const float* L = a.masks + (long long)q * a.q_stride + (long long)t * a.h4 * a.w4;
    const bool x_ok = ox < a.out_w;
    int cbase;
    float wxc[3];
    composite_taps(x_ok ? ox : a.out_w - 1, s2x, a.img_w, s1x, a.w4, cbase, wxc);
    const int c0 = cbase, c1 = min(cbase + 1, a.w4 - 1), c2 = min(cbase + 2, a.w4 - 1);
    auto hrow = [&](int r) {
      const float* p = L + (long long)min(r, a.h4 - 1) * a.w4;
      return wxc[0] * __ldg(p + c0) + wxc[1] * __ldg(p + c1) + wxc[2] * __ldg(p + c2);
    };
    int rbase = -1000;
    float H0 = 0.f, H1 = 0.f, H2 = 0.f;
    for (int oy = oy0; oy < oy1; ++oy) {
      const float4 tb = s_tab[oy - oy0];
      const int nb = __float_as_int(tb.x);
      if (nb != rbase) {
        if (nb == rbase + 1) { H0 = H1; H1 = H2; H2 = hrow(nb + 2); }
        else { H0 = hrow(nb); H1 = hrow(nb + 1); H2 = hrow(nb + 2); }
        rbase = nb;
      }
      const float v = tb.y * H0 + tb.z * H1 + tb.w * H2;
      const uint32_t word = __ballot_sync(0xffffffffu, x_ok && v > 0.f);
      if (lane == 0) s_out[oy - oy0][warp] = word;
    }
  }
  __syncthreads();
  uint32_t* out = a.bits + ((long long)plane * a.out_h + oy0) * a.words + wx0;
  const int nw = min(8, a.words - wx0);
  for (int i = threadIdx.x; i < (oy1 - oy0) * 8; i += 256) {
    const int r = i >> 3, w = i & 7;
    if (w < nw) out[(long long)r * a.words + w] = s_out[r][w];
  }
}

// Fast path of the same post-processing for the common evaluation case: output size = image size (the second interpolation
// is the identity) and an exact x4 first interpolation (padded size = 4 x the stride-4 map).  Then every output row / column
// sits between two low-resolution samples with one of four fixed fractions: output index 4k + 2 + j (j = 0..3) reads samples
// (k, k + 1) with weight (2j + 1) / 8, indices clamped at the borders -- exactly F.interpolate's align_corners = False
// arithmetic for a scale of 1/4.  A thread owns an output column; per low-resolution row pair it loads two samples, forms
// H1 and D = H1 - H0, and the four output rows cost one multiply-add, one compare and one ballot each (6-7 instructions
// per row against 26 in the general kernel, which is issue-bound at 79 % of the issue slots).
constexpr int MPX_PAIRS = 30;                                  // low-resolution row pairs per CTA = 120 output rows
__global__ void __launch_bounds__(256)
mask_postprocess_x4_kernel(const MaskPostArgs a) {
  const int plane = blockIdx.z;
  const int sel = plane / a.T, t = plane - sel * a.T;
  const int ox = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wx0 = blockIdx.x * 8;
  const int k0 = (int)blockIdx.y * MPX_PAIRS - 1;             // first pair of this CTA (pair -1 = rows 0, 1: both clamp to sample 0)
  const int k1 = min(k0 + MPX_PAIRS, a.h4);                   // pairs k0 .. k1 - 1; the last pair is (h4 - 1, h4 - 1)
  const int row0 = 4 * k0 + 2;                                // first output row of the CTA's window (may be -2)
  __shared__ __align__(16) uint32_t s_out[8][4 * MPX_PAIRS];
  if (wx0 + warp < a.words) {
    const int q = __ldg(a.query + sel);
    const float* L = a.masks + (long long)q * a.q_stride + (long long)t * a.h4 * a.w4;
    const bool x_ok = ox < a.out_w;
    const int xs = x_ok ? ox : a.out_w - 1;
    const int c = (xs + 2) / 4 - 1;                           // floor((x - 2) / 4) for x >= -2
    const float lx = (float)(2 * ((xs + 2) & 3) + 1) * 0.125f;
    const float* p0 = L + max(c, 0);
    const float* p1 = L + min(c + 1, a.w4 - 1);
    auto ldrow = [&](int r, float& v0, float& v1) {
      const int ro = min(max(r, 0), a.h4 - 1) * a.w4;
      v0 = __ldg(p0 + ro);
      v1 = __ldg(p1 + ro);
    };
    float u0, u1, n0, n1;
    ldrow(k0, u0, u1);
    ldrow(k0 + 1, n0, n1);
    float H0 = fmaf(lx, u1 - u0, u0);
    for (int k = k0; k < k1; ++k) {
      float m0, m1;
      ldrow(k + 2, m0, m1);                                   // the samples of the NEXT pair are in flight during this one
      const float H1 = fmaf(lx, n1 - n0, n0);
      n0 = m0; n1 = m1;
      const float D = H1 - H0;
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = __ballot_sync(0xffffffffu, x_ok && fmaf((float)(2 * j + 1) * 0.125f, D, H0) > 0.f);
      if (lane == 0) *reinterpret_cast<uint4*>(&s_out[warp][4 * (k - k0)]) = make_uint4(w[0], w[1], w[2], w[3]);
      H0 = H1;
    }
  }
  __syncthreads();
  const int nrows = 4 * (k1 - k0);
  const int nw = min(8, a.words - wx0);
  for (int i = threadIdx.x; i < nrows * 8; i += 256) {
    const int r = i >> 3, w = i & 7;
    const int oy = row0 + r;
    if (w < nw && oy >= 0 && oy < a.out_h) a.bits[((long long)plane * a.out_h + oy) * a.words + wx0 + w] = s_out[w][r];
  }
}

}  // namespace ovis
