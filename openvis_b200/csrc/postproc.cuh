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
// plus labels, query indices and the per-query entropy (:271).  One CTA; k <= 32.  Output sorted by score (descending);
// ties resolve to the lower flat index.
__global__ void __launch_bounds__(1024)
topk_scores_kernel(const float* __restrict__ scores, int Q, int K, int k, float* __restrict__ out_scores,
                   int* __restrict__ out_query, int* __restrict__ out_label, float* __restrict__ out_entropy) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float w_val[32];
  __shared__ int w_idx[32];
  const long long n = (long long)Q * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    long long bi = -1;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      bool taken = false;
      for (int j = 0; j < r; ++j) taken |= (s_idx[j] == (int)i);
      const float v = __ldg(scores + i);
      if (!taken && (v > best || (v == best && (bi < 0 || i < bi)))) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (ov > best || (ov == best && (bi < 0 || oi < bi)))) { best = ov; bi = oi; }
    }
    if (lane == 0) { w_val[warp] = best; w_idx[warp] = (int)bi; }
    __syncthreads();
    if (warp == 0) {
      float v = lane < (int)(blockDim.x >> 5) ? w_val[lane] : -INFINITY;
      int i = lane < (int)(blockDim.x >> 5) ? w_idx[lane] : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (oi >= 0 && (ov > v || (ov == v && (i < 0 || oi < i)))) { v = ov; i = oi; }
      }
      if (lane == 0) { s_val[r] = v; s_idx[r] = i; }
    }
    __syncthreads();
  }
  // outputs + entropy of the selected queries' score rows: sum(-s * log s)
  for (int r = warp; r < k; r += (int)(blockDim.x >> 5)) {
    const int flat = s_idx[r];
    const int q = flat / K;
    float e = 0.f;
    for (int c = lane; c < K; c += 32) {
      const float s = __ldg(scores + (long long)q * K + c);
      e += -s * logf(s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) {
      out_scores[r] = s_val[r];
      out_query[r] = q;
      out_label[r] = flat - q * K;
      out_entropy[r] = e;
    }
  }
}

struct MaskPostArgs {
  const float* masks;     // [Q][T][h4][w4] stride-4 logits (pred_masks[0])
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

// thread = output pixel, warp = 32 consecutive x of one output row -> one ballot word
__global__ void __launch_bounds__(256)
mask_postprocess_kernel(const MaskPostArgs a) {
  const int wpr = a.words;                                   // warps per output row
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long total_warps = (long long)a.n_sel * a.T * a.out_h * wpr;
  if (warp_global >= total_warps) return;
  const int lane = threadIdx.x & 31;
  const int wx = (int)(warp_global % wpr);
  long long r = warp_global / wpr;
  const int oy = (int)(r % a.out_h); r /= a.out_h;
  const int t = (int)(r % a.T);
  const int sel = (int)(r / a.T);
  const int ox = wx * 32 + lane;
  const int q = __ldg(a.query + sel);
  const float* L = a.masks + ((long long)q * a.T + t) * a.h4 * a.w4;
  const float s2y = (float)a.img_h / a.out_h, s2x = (float)a.img_w / a.out_w;     // second resize (image -> output)
  const float s1y = (float)a.h4 / a.pad_h, s1x = (float)a.w4 / a.pad_w;           // first resize (stride 4 -> padded)
  bool pos = false;
  if (ox < a.out_w) {
    int iy[2], ix[2];
    float ly, lx;
    bilin_src(oy, s2y, a.img_h, iy[0], iy[1], ly);
    bilin_src(ox, s2x, a.img_w, ix[0], ix[1], lx);
    // intermediate (up-sampled, cropped) image at the 2 x 2 taps
    float I[2][2];
#pragma unroll
    for (int jy = 0; jy < 2; ++jy) {
      int y0, y1; float wy;
      bilin_src(iy[jy], s1y, a.h4, y0, y1, wy);
#pragma unroll
      for (int jx = 0; jx < 2; ++jx) {
        int x0, x1; float wx1;
        bilin_src(ix[jx], s1x, a.w4, x0, x1, wx1);
        const float v00 = __ldg(L + y0 * a.w4 + x0), v01 = __ldg(L + y0 * a.w4 + x1);
        const float v10 = __ldg(L + y1 * a.w4 + x0), v11 = __ldg(L + y1 * a.w4 + x1);
        I[jy][jx] = (1.f - wy) * ((1.f - wx1) * v00 + wx1 * v01) + wy * ((1.f - wx1) * v10 + wx1 * v11);
      }
    }
    const float v = (1.f - ly) * ((1.f - lx) * I[0][0] + lx * I[0][1]) + ly * ((1.f - lx) * I[1][0] + lx * I[1][1]);
    pos = v > 0.f;
  }
  const uint32_t word = __ballot_sync(0xffffffffu, pos);
  if (lane == 0) a.bits[warp_global] = word;
}

}  // namespace ovis
