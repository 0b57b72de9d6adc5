// OpenVIS crop classifier, front end: ClipAdapter._preprocess_image + the input half of
// encode_image (reference: openvis/modeling/clip_adapter/adapter.py:73-116, 140-144) and the token assembly in front of the
// CLIP visual transformer (third_parties/mask_adapted_clip/.../model.py:327-342).
//
//   mask_boxes_kernel     valid flag + bounding box of (mask > 0.5) for every (frame, query)        adapter.py:84-92
//                         (detectron2 BitMasks.get_bounding_boxes: [x_min, y_min, x_max + 1, y_max + 1])
//   crop_blend_kernel     square crop box, roi_align of the frame and of the soft mask to R x R, blend  adapter.py:93-113
//                         (torchvision roi_align: spatial_scale 1, sampling_ratio -1 = adaptive, aligned = False)
//   patchify_kernel       / 255, CLIP mean / std, 16 x 16 patches as rows of the patch-embedding GEMM   adapter.py:141-142
//   clip_embed_kernel     [class_embedding | conv1 tokens] + positional_embedding -> ln_pre             model.py:340-342
//
// The reference runs roi_align in fp16 on the GPU (`frames.half()`, `ind_boxes.half()`): inputs and outputs are rounded to
// fp16 here as well, but sample positions, bilinear weights and the bin average are fp32 (fp16 positions quantise to 0.5-1
// pixel beyond x = 1024; the oracle is the reference's own code evaluated in fp32).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ovis {

// masks are addressed as masks[t * stride_t + n * stride_n + y * W + x]: [T][N][H][W] soft masks as ClipAdapter.forward takes
// them, or the decoder's [N][T][H][W] logits directly (`logits` = 1: the sigmoid of openvis.py:118 is applied on load; the
// threshold test sigmoid(x) > thresh becomes x > logit(thresh)), without the transposed fp32 copy the reference makes.
__global__ void __launch_bounds__(256)
mask_boxes_kernel(const float* __restrict__ masks, int N, long long stride_t, long long stride_n, int H, int W, float thresh,
                  int* __restrict__ boxes, unsigned char* __restrict__ valid) {
  const long long m = blockIdx.x;
  const float* p = masks + (m / N) * stride_t + (m % N) * stride_n;
  int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
  const long long total = (long long)H * W;
  if ((W & 3) == 0 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const int w4 = W >> 2;
    for (long long i = threadIdx.x; i < total / 4; i += blockDim.x) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(p) + i);
      const int y = (int)(i / w4), x = (int)(i % w4) * 4;
      const bool b0 = v.x > thresh, b1 = v.y > thresh, b2 = v.z > thresh, b3 = v.w > thresh;
      if (b0 | b1 | b2 | b3) {
        const int lo = b0 ? x : b1 ? x + 1 : b2 ? x + 2 : x + 3;
        const int hi = b3 ? x + 3 : b2 ? x + 2 : b1 ? x + 1 : x;
        x0 = min(x0, lo); x1 = max(x1, hi);
        y0 = min(y0, y); y1 = max(y1, y);
      }
    }
  } else {
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
      if (__ldcs(p + i) > thresh) {
        const int y = (int)(i / W), x = (int)(i % W);
        x0 = min(x0, x); x1 = max(x1, x);
        y0 = min(y0, y); y1 = max(y1, y);
      }
    }
  }
  x0 = __reduce_min_sync(0xffffffffu, x0); y0 = __reduce_min_sync(0xffffffffu, y0);
  x1 = __reduce_max_sync(0xffffffffu, x1); y1 = __reduce_max_sync(0xffffffffu, y1);
  __shared__ int s[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s[warp][0] = x0; s[warp][1] = y0; s[warp][2] = x1; s[warp][3] = y1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      x0 = min(x0, s[w][0]); y0 = min(y0, s[w][1]);
      x1 = max(x1, s[w][2]); y1 = max(y1, s[w][3]);
    }
    const bool any = x1 >= 0;
    valid[m] = any ? 1 : 0;
    reinterpret_cast<int4*>(boxes)[m] = any ? make_int4(x0, y0, x1 + 1, y1 + 1) : make_int4(0, 0, 0, 0);
  }
}

struct CropArgs {
  const float* frames;   // [T][3][H][W], 0..255
  const float* masks;    // soft masks (after the sigmoid) or logits, see mask_boxes_kernel
  long long stride_t, stride_n;
  int logits;
  const int* ids;        // [M][2] (frame, query) of the valid regions, row-major order of `valid`
  const int* boxes;      // [T * N][4]
  __half* regions;       // [M][3][R][R]
  int N, H, W, R;
};

__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }

__global__ void __launch_bounds__(256)
crop_blend_kernel(const CropArgs a) {
  const int m = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= a.R * a.R) return;
  const int oy = pix / a.R, ox = pix % a.R;
  const int t = a.ids[2 * m], n = a.ids[2 * m + 1];
  const int4 b = reinterpret_cast<const int4*>(a.boxes)[(long long)t * a.N + n];
  // xyxy -> square box anchored at the top-left corner, side = max(width, height)  (adapter.py:94-100)
  const int side = max(b.z - b.x, b.w - b.y);
  const float roi_w = fmaxf((float)side, 1.f);           // torchvision: aligned = False forces rois of at least 1 x 1
  const float bin = roi_w / (float)a.R;
  const int grid = (int)ceilf(roi_w / (float)a.R);        // sampling_ratio = -1: ceil(roi / pooled) samples per bin side
  const float inv_count = 1.f / (float)max(grid * grid, 1);
  const long long plane = (long long)a.H * a.W;
  const float* f0 = a.frames + (long long)t * 3 * plane;
  const float* mk = a.masks + t * a.stride_t + n * a.stride_n;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accm = 0.f;
  for (int iy = 0; iy < grid; ++iy) {
    float y = (float)b.y + oy * bin + ((float)iy + 0.5f) * bin / (float)grid;
    for (int ix = 0; ix < grid; ++ix) {
      float x = (float)b.x + ox * bin + ((float)ix + 0.5f) * bin / (float)grid;
      if (y < -1.f || y > (float)a.H || x < -1.f || x > (float)a.W) continue;     // outside the image: contributes 0
      float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
      int yl = (int)yy, xl = (int)xx, yh, xh;
      if (yl >= a.H - 1) { yh = yl = a.H - 1; yy = (float)yl; } else yh = yl + 1;
      if (xl >= a.W - 1) { xh = xl = a.W - 1; xx = (float)xl; } else xh = xl + 1;
      const float ly = yy - yl, lx = xx - xl, hy = 1.f - ly, hx = 1.f - lx;
      const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
      const long long i1 = (long long)yl * a.W + xl, i2 = (long long)yl * a.W + xh, i3 = (long long)yh * a.W + xl,
                      i4 = (long long)yh * a.W + xh;
      acc0 += w1 * h_round(__ldg(f0 + i1)) + w2 * h_round(__ldg(f0 + i2)) + w3 * h_round(__ldg(f0 + i3)) + w4 * h_round(__ldg(f0 + i4));
      acc1 += w1 * h_round(__ldg(f0 + plane + i1)) + w2 * h_round(__ldg(f0 + plane + i2)) + w3 * h_round(__ldg(f0 + plane + i3)) +
              w4 * h_round(__ldg(f0 + plane + i4));
      acc2 += w1 * h_round(__ldg(f0 + 2 * plane + i1)) + w2 * h_round(__ldg(f0 + 2 * plane + i2)) +
              w3 * h_round(__ldg(f0 + 2 * plane + i3)) + w4 * h_round(__ldg(f0 + 2 * plane + i4));
      float m1 = __ldg(mk + i1), m2 = __ldg(mk + i2), m3 = __ldg(mk + i3), m4 = __ldg(mk + i4);
      if (a.logits) {
        m1 = 1.f / (1.f + __expf(-m1)); m2 = 1.f / (1.f + __expf(-m2));
        m3 = 1.f / (1.f + __expf(-m3)); m4 = 1.f / (1.f + __expf(-m4));
      }
      accm += w1 * h_round(m1) + w2 * h_round(m2) + w3 * h_round(m3) + w4 * h_round(m4);
    }
  }
  // both roi_align outputs are fp16 tensors; blend = mask_region * region (+ (1 - mask_region) * 0)   (adapter.py:113)
  const float mr = h_round(accm * inv_count);
  __half* o = a.regions + ((long long)m * 3) * a.R * a.R + pix;
  o[0] = __float2half_rn(mr * h_round(acc0 * inv_count));
  o[(long long)a.R * a.R] = __float2half_rn(mr * h_round(acc1 * inv_count));
  o[2ll * a.R * a.R] = __float2half_rn(mr * h_round(acc2 * inv_count));
}

// regions [M][3][R][R] fp16 (0..255) -> rows of the patch-embedding GEMM: out[(m * gp * gp + py * gp + px)][c * P * P + ky * P + kx]
// = (v / 255 - mean[c]) / std[c]   (encode_image: image / 255, bicubic resize to R x R = identity at this size, Normalize;
// conv1 weight [width][3][P][P] flattened is the GEMM's B operand)
__global__ void __launch_bounds__(256)
patchify_kernel(const __half* __restrict__ regions, __half* __restrict__ out, int R, int P, float m0, float m1, float m2,
                float is0, float is1, float is2, long long rows) {
  const long long row = blockIdx.x;
  if (row >= rows) return;
  const int gp = R / P;
  const int px = (int)(row % gp), py = (int)((row / gp) % gp);
  const long long m = row / (gp * gp);
  const int K = 3 * P * P;
  for (int e = threadIdx.x; e < K; e += blockDim.x) {
    const int c = e / (P * P), ky = (e / P) % P, kx = e % P;
    const float v = __half2float(regions[((m * 3 + c) * R + py * P + ky) * R + px * P + kx]);
    const float mean = c == 0 ? m0 : c == 1 ? m1 : m2, is = c == 0 ? is0 : c == 1 ? is1 : is2;
    out[row * K + e] = __float2half_rn((v * (1.f / 255.f) - mean) * is);
  }
}

// X[m * (1 + Lp) + tok] = ln_pre((tok == 0 ? class_embedding : patch_tokens[m * Lp + tok - 1]) + positional_embedding[tok])
// one warp per row, width <= 1024 and a multiple of 32
__global__ void __launch_bounds__(256)
clip_embed_kernel(const float* __restrict__ patch_tokens, const float* __restrict__ cls, const float* __restrict__ pos,
                  const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ X, int Lp, int width,
                  long long rows) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int tok = (int)(row % (1 + Lp));
  const long long m = row / (1 + Lp);
  const float* src = tok == 0 ? cls : patch_tokens + (m * Lp + tok - 1) * width;
  const int per = width / 32;
  float v[32];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i < per) {
      const int c = i * 32 + lane;
      v[i] = src[c] + __ldg(pos + (long long)tok * width + c);
      sum += v[i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)width;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) { const float d = v[i] - mean; sq += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)width + 1e-5f);
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      X[row * width + c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
}

}  // namespace ovis
