// TMA-fed version of maskfeat_prep_kernel (prep.cuh): same outputs, bit for bit.
//   F [B][C][H][W] fp32  ->  ft [B][H*W][C] fp16 and the centre-2x2 means g2 / g1 / g0 (blocks of 2 / 4 / 8 pixels).
// HBM-bound (6 bytes per input element + the pooled maps), so the kernel is organised around the memory system:
// persistent CTAs, a 4-stage ring of 32 KB input tiles (32 channels x 8 rows x 32 columns) filled by bulk tensor loads,
// the transposition done shared -> registers -> 64B-swizzled shared staging, and bulk tensor stores that write whole
// 64-byte channel runs of 256 pixels per instruction (the neighbouring channel blocks of the same pixels are handled by
// the neighbouring CTAs at the same time, so L2 assembles full 512-byte token rows).
#pragma once
#include "ptx.cuh"

namespace ovis {

constexpr int PT_STAGES = 4;
constexpr int PT_IN_BYTES = 32 * 8 * 32 * 4;                 // 32 KB
constexpr int PT_FT_BYTES = 256 * 64;                        // [8][32] pixels x 32 ch fp16
constexpr int PT_G2_BYTES = 64 * 64, PT_G1_BYTES = 16 * 64, PT_G0_BYTES = 4 * 64;
constexpr int PT_STG_BYTES = PT_FT_BYTES + PT_G2_BYTES + 1024 + 1024;     // g1, g0 padded to the swizzle period
constexpr int PT_SMEM = PT_STAGES * PT_IN_BYTES + 2 * PT_STG_BYTES + 1024 + 256;

struct PrepTmaMaps {
  CUtensorMap in, ft, g2, g1, g0;
};

// 16-byte chunk `chunk` of 64-byte row `row` under CU_TENSOR_MAP_SWIZZLE_64B (address bits [4,5] ^= bits [7,8])
__device__ __forceinline__ uint32_t sw64(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }

__global__ void __launch_bounds__(256, 1)
maskfeat_prep_tma_kernel(const __grid_constant__ PrepTmaMaps maps, int B, int C, int H, int W) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_in = smem;
  uint8_t* s_stg = smem + PT_STAGES * PT_IN_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_stg + 2 * PT_STG_BYTES);

  const int tiles_x = (W + 31) / 32, tiles_y = H / 8, cblocks = C / 32;
  const long long total = (long long)B * tiles_y * tiles_x * cblocks;
  const int tid = threadIdx.x;

  auto issue_load = [&](long long tile, int s) {
    const int cb = (int)(tile % cblocks);
    long long r = tile / cblocks;
    const int tx = (int)(r % tiles_x); r /= tiles_x;
    const int ty = (int)(r % tiles_y);
    const int b = (int)(r / tiles_y);
    mbar_arrive_expect_tx(&full[s], PT_IN_BYTES);
    tma_load_4d(s_in + s * PT_IN_BYTES, &maps.in, &full[s], tx * 32, ty * 8, cb * 32, b);
  };

  if (tid == 0) {
    tma_prefetch_desc(&maps.in); tma_prefetch_desc(&maps.ft); tma_prefetch_desc(&maps.g2);
    tma_prefetch_desc(&maps.g1); tma_prefetch_desc(&maps.g0);
    for (int s = 0; s < PT_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < PT_STAGES - 1; ++s) {
      const long long tile = blockIdx.x + (long long)s * gridDim.x;
      if (tile < total) issue_load(tile, s);
    }
  }

  for (long long it = 0;; ++it) {
    const long long tile = blockIdx.x + it * gridDim.x;
    if (tile >= total) break;
    const int s = (int)(it % PT_STAGES);
    const float* in = reinterpret_cast<const float*>(s_in + s * PT_IN_BYTES);       // [32 ch][8 y][32 x]
    uint8_t* stg = s_stg + (it & 1) * PT_STG_BYTES;
    const uint32_t ft_s = smem_u32(stg), g2_s = ft_s + PT_FT_BYTES, g1_s = g2_s + PT_G2_BYTES, g0_s = g1_s + 1024;
    mbar_wait(&full[s], (uint32_t)((it / PT_STAGES) & 1));

    // ---- full-resolution token-major copy: item = (pixel, 8-channel group); lanes run along x (conflict-free loads)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int item = i * 256 + tid;
      const int pix = item & 255, cg = item >> 8;
      const float* p = in + cg * 8 * 256 + pix;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = p[e * 256];
      st_shared_v4(ft_s + sw64(pix, cg), make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]),
                                                    pack_half2(v[4], v[5]), pack_half2(v[6], v[7])));
    }
    // ---- level 2: 2x2 blocks, 4 x 16 cells; item = (cell, channel group)
    {
      const int cell = tid & 63, cg = tid >> 6;
      const int cy = cell >> 4, cx = cell & 15;
      const float* p = in + cg * 8 * 256 + (2 * cy) * 32 + 2 * cx;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 a = *reinterpret_cast<const float2*>(p + e * 256);
        const float2 c = *reinterpret_cast<const float2*>(p + e * 256 + 32);
        v[e] = 0.25f * ((a.x + a.y) + (c.x + c.y));
      }
      st_shared_v4(g2_s + sw64(cell, cg), make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]),
                                                     pack_half2(v[4], v[5]), pack_half2(v[6], v[7])));
    }
    // ---- level 1: 4x4 blocks (centre rows / columns 1, 2), 2 x 8 cells: threads 0..63
    //      level 0: 8x8 blocks (centre rows / columns 3, 4), 1 x 4 cells: threads 64..79
    if (tid < 80) {
      const bool l1 = tid < 64;
      const int t2 = l1 ? tid : tid - 64;
      const int cell = l1 ? (t2 & 15) : (t2 & 3), cg = l1 ? (t2 >> 4) : (t2 >> 2);
      const int off = l1 ? ((4 * (cell >> 3) + 1) * 32 + 4 * (cell & 7) + 1) : (3 * 32 + 8 * cell + 3);
      const float* p = in + cg * 8 * 256 + off;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.25f * ((p[e * 256] + p[e * 256 + 1]) + (p[e * 256 + 32] + p[e * 256 + 33]));
      st_shared_v4((l1 ? g1_s : g0_s) + sw64(cell, cg), make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]),
                                                                   pack_half2(v[4], v[5]), pack_half2(v[6], v[7])));
    }
    fence_async_proxy();                       // staging writes -> visible to the bulk stores
    if (tid == 0) tma_store_wait_read0();      // the other staging buffer (previous tile's stores) has been read
    __syncthreads();
    if (tid == 0) {
      const int cb = (int)(tile % cblocks);
      long long r = tile / cblocks;
      const int tx = (int)(r % tiles_x); r /= tiles_x;
      const int ty = (int)(r % tiles_y);
      const int b = (int)(r / tiles_y);
      tma_store_4d(&maps.ft, stg, cb * 32, tx * 32, ty * 8, b);
      tma_store_4d(&maps.g2, stg + PT_FT_BYTES, cb * 32, tx * 16, ty * 4, b);
      tma_store_4d(&maps.g1, stg + PT_FT_BYTES + PT_G2_BYTES, cb * 32, tx * 8, ty * 2, b);
      tma_store_4d(&maps.g0, stg + PT_FT_BYTES + PT_G2_BYTES + 1024, cb * 32, tx * 4, ty, b);
      tma_store_commit();
      // the stage consumed one iteration ago is free (every thread passed this barrier after reading it)
      const long long nt = blockIdx.x + (it + PT_STAGES - 1) * gridDim.x;
      if (nt < total) issue_load(nt, (int)((it + PT_STAGES - 1) % PT_STAGES));
    }
  }
  if (tid == 0) tma_store_wait0();
}

}  // namespace ovis

namespace ovis {

// TMA-fed version of nchw_to_tokens_f16_kernel (prep.cuh): x [B][C][h][w] fp32 -> xt [B][h*w][C] fp16 and, when `pos_cn`
// is given, xp = fp16(x + (pos_cn[c][n] + pos_t[b][c])).  Same ring / staging scheme as above.  The position table does
// not depend on the frame, so a CTA walks ONE spatial tile through a run of frames: its 32 position values per thread
// are read once (channel-major table = coalesced along x) and stay in registers, the per-frame terms of the run sit in
// shared memory, and the steady-state loop touches global memory only through the bulk tensor loads and stores.
struct TokTmaMaps {
  CUtensorMap in, xt, xp;
};
constexpr int TT_STG_BYTES = 2 * PT_FT_BYTES;                 // xt tile + xp tile
constexpr int TT_MAX_RUN = 64;                                // frames per run (pos_t staging: 64 x 32 floats)
constexpr int TT_SMEM = PT_STAGES * PT_IN_BYTES + 2 * TT_STG_BYTES + TT_MAX_RUN * 32 * 4 + 1024 + 256;

__global__ void __launch_bounds__(256, 1)
tokens_prep_tma_kernel(const __grid_constant__ TokTmaMaps maps, const float* __restrict__ pos_cn,
                       const float* __restrict__ pos_t, int B, int C, int H, int W, int runs, int run_len) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_in = smem;
  uint8_t* s_stg = smem + PT_STAGES * PT_IN_BYTES;
  float* s_pt = reinterpret_cast<float*>(s_stg + 2 * TT_STG_BYTES);           // [run_len][32]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_stg + 2 * TT_STG_BYTES + TT_MAX_RUN * 32 * 4);

  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8, cblocks = C / 32;
  const long long spatial = (long long)tiles_y * tiles_x * cblocks;
  const long long items = spatial * runs;                     // item = (spatial tile, run of frames); channel block fastest
  const int tid = threadIdx.x;
  const long long N = (long long)H * W;

  auto decode = [&](long long item, int& cb, int& tx, int& ty, int& b0, int& b1) {
    const long long sp = item % spatial;
    const int run = (int)(item / spatial);
    cb = (int)(sp % cblocks);
    const long long r = sp / cblocks;
    tx = (int)(r % tiles_x);
    ty = (int)(r / tiles_x);
    b0 = run * run_len;
    b1 = min(B, b0 + run_len);
  };
  // producer cursor (thread 0): next (item, frame) to request
  long long p_item = blockIdx.x;
  int p_b = 0, p_b1 = 0, p_cb = 0, p_tx = 0, p_ty = 0;
  long long p_count = 0;
  auto producer_issue = [&]() {                               // returns after issuing one load (if any work is left)
    if (p_item >= items) return;
    if (p_b >= p_b1) return;
    const int s = (int)(p_count % PT_STAGES);
    mbar_arrive_expect_tx(&full[s], PT_IN_BYTES);
    tma_load_4d(s_in + s * PT_IN_BYTES, &maps.in, &full[s], p_tx * 32, p_ty * 8, p_cb * 32, p_b);
    ++p_count;
    if (++p_b >= p_b1) {
      p_item += gridDim.x;
      if (p_item < items) decode(p_item, p_cb, p_tx, p_ty, p_b, p_b1);
    }
  };

  if (tid == 0) {
    tma_prefetch_desc(&maps.in); tma_prefetch_desc(&maps.xt); tma_prefetch_desc(&maps.xp);
    for (int s = 0; s < PT_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
    if (p_item < items) decode(p_item, p_cb, p_tx, p_ty, p_b, p_b1);
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < PT_STAGES - 1; ++s) producer_issue();
  }

  long long it = 0;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    int cb, tx, ty, b0, b1;
    decode(item, cb, tx, ty, b0, b1);
    // ---- per-item state: position values of this thread's four (pixel, channel group) items, frame terms of the run
    float pv[4][8];
    if (pos_cn) {
      __syncthreads();                                        // previous item's readers of s_pt are done
      for (int i = tid; i < (b1 - b0) * 32; i += 256)
        s_pt[i] = pos_t ? __ldg(pos_t + (long long)(b0 + (i >> 5)) * C + cb * 32 + (i & 31)) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pix = tid, cg = i;                          // item index i * 256 + tid
        const int x = tx * 32 + (pix & 31), y = ty * 8 + (pix >> 5);
        const bool ok = x < W && y < H;
        const float* pp = pos_cn + (long long)(cb * 32 + cg * 8) * N + (long long)y * W + x;
#pragma unroll
        for (int e = 0; e < 8; ++e) pv[i][e] = ok ? __ldg(pp + (long long)e * N) : 0.f;
      }
      __syncthreads();
    }
    for (int b = b0; b < b1; ++b, ++it) {
      const int s = (int)(it % PT_STAGES);
      const float* in = reinterpret_cast<const float*>(s_in + s * PT_IN_BYTES);       // [32 ch][8 y][32 x]
      uint8_t* stg = s_stg + (it & 1) * TT_STG_BYTES;
      const uint32_t xt_s = smem_u32(stg), xp_s = xt_s + PT_FT_BYTES;
      mbar_wait(&full[s], (uint32_t)((it / PT_STAGES) & 1));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pix = tid, cg = i;
        const float* p = in + cg * 8 * 256 + pix;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = p[e * 256];
        st_shared_v4(xt_s + sw64(pix, cg), make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]),
                                                      pack_half2(v[4], v[5]), pack_half2(v[6], v[7])));
        if (pos_cn) {
          const float4 t0 = *reinterpret_cast<const float4*>(s_pt + (b - b0) * 32 + cg * 8);
          const float4 t1 = *reinterpret_cast<const float4*>(s_pt + (b - b0) * 32 + cg * 8 + 4);
          const float tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += pv[i][e] + tt[e];
          st_shared_v4(xp_s + sw64(pix, cg), make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]),
                                                        pack_half2(v[4], v[5]), pack_half2(v[6], v[7])));
        }
      }
      fence_async_proxy();
      if (tid == 0) tma_store_wait_read0();
      __syncthreads();
      if (tid == 0) {
        tma_store_4d(&maps.xt, stg, cb * 32, tx * 32, ty * 8, b);
        if (pos_cn) tma_store_4d(&maps.xp, stg + PT_FT_BYTES, cb * 32, tx * 32, ty * 8, b);
        tma_store_commit();
        producer_issue();     // refills the stage consumed one iteration ago (every thread passed the barrier since)
      }
    }
  }
  if (tid == 0) tma_store_wait0();
}

}  // namespace ovis
