// Attention of the CLIP residual blocks on tcgen05 / TMEM / TMA (round 2; same arithmetic and bias structure as
// san_attn_mma_kernel in san_attn.cuh, which stays as the checker and serves more than 255 patches).
//
// Token sequence of one image: [Q SOS tokens | CLS | L patches] (Q = 0 for the plain CLIP visual tower of the crop
// classifier), d = 64 per head.  Every row attends to the CLS + patch keys (SOS rows: not CLS, patches with the pooled
// additive bias, plus their own key); see san_attn.cuh for the structure of SideAdapter._build_attn_biases that this uses.
//
// One CTA per (image, head): K and V of the 1 + L keys are loaded ONCE (two 128-row TMA boxes each, 128-byte swizzle) and
// stay in shared memory while the CTA walks the image's 128-row query tiles:
//   S   = Q_tile K^T        UMMA 128 x N x 16 (N = keys rounded up to 16, <= 256), 4 K-steps, fp32 in TMEM columns [0, N)
//   P   = softmax rows      thread = query row, the two warps of a TMEM lane quarter split the row's key chunks: pass 1 takes
//                           the row maximum (SOS warps: scale, bias, mask, biased scores parked back in TMEM; plain warps:
//                           the raw maximum only), pass 2 exponentiates (one FFMA + MUFU.EX2 per element), sums and writes
//                           fp16 P as a K-major 128B-swizzled A operand [128][256]
//   O   = P V               UMMA 128 x 64 x 16 with V as an MN-major B operand (its natural [key][d] layout), N/16 K-steps
// The Q tile is double-buffered, so the next tile's TMA load and score product run under the current tile's softmax / output.
#pragma once
#include "ptx.cuh"
#include "xattn_tc.cuh"
#include "san_attn.cuh"

namespace ovis {

constexpr int ST_THREADS = 320;                 // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: softmax
constexpr int ST_TILE = 128 * 128;              // one 128-row x 64-column fp16 box
constexpr int ST_OFF_Q = 0;                     // 2 Q tiles
constexpr int ST_OFF_K = 2 * ST_TILE;           // keys 0..255
constexpr int ST_OFF_V = 4 * ST_TILE;
constexpr int ST_OFF_P = 6 * ST_TILE;           // [4 key atoms of 64][128 rows][128 B]
constexpr int ST_OFF_BAR = 10 * ST_TILE;
constexpr int ST_SMEM = ST_OFF_BAR + 256 + 2048 /*row max / sum exchange*/ + 1024;
constexpr int ST_MAX_KEYS = 256;

__global__ void __launch_bounds__(ST_THREADS, 1)
san_attn_tc_kernel(const __grid_constant__ CUtensorMap tm, const SanAttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ST_OFF_BAR);
  uint64_t* kv_full = bars;          // K and V resident
  uint64_t* q_full = bars + 1;       // [2]
  uint64_t* q_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;       // scores of the tile in TMEM
  uint64_t* p_full = bars + 6;       // P written by the four softmax warps
  uint64_t* o_full = bars + 7;       // output accumulator complete
  uint64_t* k_free = bars + 8;       // the item's last score product has retired: K may be overwritten
  uint64_t* v_free = bars + 9;       // the item's last PV product has retired: V may be overwritten
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lt = a.Q + 1 + a.L, W = a.heads * 64, W3 = 3 * W;
  const int nkeys = 1 + a.L;
  const int nk16 = (nkeys + 15) >> 4;               // K-steps of the PV product; the score product's N = 16 * nk16
  const int mtiles = (Lt + 127) >> 7;
  const int items = a.B * a.heads;
  // persistent CTAs: item = (image, head), walked with the grid stride; all pipeline parities count tiles / items per CTA

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(o_full, 1);
    mbar_init(k_free, 1);
    mbar_init(v_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t t_s = tmem_base, t_o = tmem_base + 256u;

  if (warp == 0) {
    if (lane == 0) {
      int t = 0, it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int head = item % a.heads;
        const long long row0 = (long long)(item / a.heads) * Lt;
        // (the expect-tx arrival must not run ahead of the previous item's phase: with two query tiles per item the Q ring
        //  alone would let this thread start the next item while the previous K / V bytes are still in flight)
        mbar_wait(k_free, (uint32_t)((it & 1) ^ 1));
        mbar_arrive_expect_tx(kv_full, 4 * ST_TILE);
        for (int i = 0; i < 2; ++i)
          tma_load_2d(smem + ST_OFF_K + i * ST_TILE, &tm, kv_full, W + head * 64, (int)(row0 + a.Q + i * 128));
        mbar_wait(v_free, (uint32_t)((it & 1) ^ 1));
        for (int i = 0; i < 2; ++i)
          tma_load_2d(smem + ST_OFF_V + i * ST_TILE, &tm, kv_full, 2 * W + head * 64, (int)(row0 + a.Q + i * 128));
        for (int m = 0; m < mtiles; ++m, ++t) {
          const int buf = t & 1;
          mbar_wait(&q_empty[buf], (uint32_t)(((t >> 1) & 1) ^ 1));
          mbar_arrive_expect_tx(&q_full[buf], ST_TILE);
          tma_load_2d(smem + ST_OFF_Q + buf * ST_TILE, &tm, &q_full[buf], head * 64, (int)(row0 + m * 128));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = xt_idesc(128, nk16 * 16, 0);
      constexpr uint32_t idesc_o = xt_idesc(128, 64, 1);          // B (V) MN-major
      const uint64_t kdesc = umma_desc_k_sw128(smem_u32(smem + ST_OFF_K));
      const uint64_t vdesc = umma_desc_mn_sw128(smem_u32(smem + ST_OFF_V));
      int t = 0, it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        mbar_wait(kv_full, (uint32_t)(it & 1));
        for (int m = 0; m < mtiles; ++m, ++t) {
          const int buf = t & 1;
          mbar_wait(&q_full[buf], (uint32_t)((t >> 1) & 1));
          tc_fence_after();
          const uint64_t qdesc = umma_desc_k_sw128(smem_u32(smem + ST_OFF_Q + buf * ST_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(t_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_full);
          umma_commit(&q_empty[buf]);
          if (m == mtiles - 1) umma_commit(k_free);
          mbar_wait(p_full, (uint32_t)(t & 1));
          tc_fence_after();
          for (int j = 0; j < nk16; ++j) {
            const uint64_t pdesc = umma_desc_k_sw128(smem_u32(smem + ST_OFF_P + (j >> 2) * ST_TILE)) + 2 * (j & 3);
            umma_f16(t_o, pdesc, vdesc + (uint64_t)(j * (2048 >> 4)), idesc_o, j > 0 ? 1u : 0u);
          }
          umma_commit(o_full);
          if (m == mtiles - 1) umma_commit(v_free);
        }
      }
    }
  } else {
    // Eight softmax warps: the two warps of a TMEM lane quarter split the key chunks (and the output channels) of a row;
    // they exchange the row maximum and the row sum through shared memory around two named barriers.
    const int quarter = warp & 3;                       // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;
    const int rt = quarter * 32 + lane;                 // row within the tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const int nchunks = (nk16 * 16 + 31) >> 5;
    const int c_lo = half == 0 ? 0 : (nchunks + 1) >> 1, c_hi = half == 0 ? (nchunks + 1) >> 1 : nchunks;
    const uint32_t p_row = smem_u32(smem + ST_OFF_P) + (uint32_t)rt * 128u;
    float* x_max = reinterpret_cast<float*>(smem + ST_OFF_BAR + 256);      // [2][128]
    float* x_sum = x_max + 256;                                            // [2][128]
    int t = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x)
    for (int m = 0; m < mtiles; ++m, ++t) {
      const int head = item % a.heads, b = item / a.heads;
      const long long row0 = (long long)b * Lt;
      const int row = m * 128 + rt;
      const bool ok = row < Lt;
      const bool sos = ok && row < a.Q;
      const bool warp_sos = m * 128 + quarter * 32 < a.Q;              // warp-uniform: some row of this warp is a SOS row
      const __half* qp = a.qkv + (row0 + (ok ? row : 0)) * W3 + head * 64;
      // a SOS row's own key (the diagonal 0 of the bias matrix): score now, value in the epilogue
      float s_self = -INFINITY;
      if (sos) {
        float sp = 0.f;
#pragma unroll
        for (int d = 0; d < 64; d += 8) {
          const uint4 qu = *reinterpret_cast<const uint4*>(qp + d);
          const uint4 ku = *reinterpret_cast<const uint4*>(qp + W + d);
          const __half2* q2 = reinterpret_cast<const __half2*>(&qu);
          const __half2* k2 = reinterpret_cast<const __half2*>(&ku);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 qf = __half22float2(q2[e]), kf = __half22float2(k2[e]);
            sp = fmaf(qf.x, kf.x, sp);
            sp = fmaf(qf.y, kf.y, sp);
          }
        }
        s_self = sp * a.scale_log2;
      }
      const float* pb = (a.pooled && sos) ? a.pooled + (((long long)b * a.heads + head) * a.Q + row) * a.L : nullptr;

      mbar_wait(s_full, (uint32_t)(t & 1));
      tc_fence_after();
      uint32_t v[32];
      float mx = s_self;                                  // in scaled (base-2) units
      if (warp_sos) {
        // general rows: scale, bias, structural mask; the biased scores are parked back in TMEM for pass 2
        for (int c = c_lo; c < c_hi; ++c) {
          tmem_ld_32x32(t_s + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = c * 32 + j;                   // 0 = CLS, 1.. = patches
            float x = __uint_as_float(v[j]) * a.scale_log2;
            if (sos) {
              if (pb && key >= 1 && key < nkeys) x = fmaf(__ldg(pb + key - 1), LOG2E, x);
              if (key == 0) x = -INFINITY;                // SOS -> CLS carries -100
            }
            if (key >= nkeys) x = -INFINITY;              // padding keys (and stale TMEM columns past N)
            v[j] = __float_as_uint(x);
            mx = fmaxf(mx, x);
          }
          tmem_st_32x32(t_s + lane_off + c * 32, v);
        }
        tmem_st_wait();
      } else {
        // plain rows (CLS / patches, the whole CLIP tower): only the maximum of the raw scores is needed here
        float raw = -INFINITY;
        for (int c = c_lo; c < c_hi; ++c) {
          tmem_ld_32x32(t_s + lane_off + c * 32, v);
          tmem_ld_wait();
          if (c * 32 + 32 <= nkeys) {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw = fmaxf(raw, __uint_as_float(v[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw = fmaxf(raw, c * 32 + j < nkeys ? __uint_as_float(v[j]) : -INFINITY);
          }
        }
        mx = raw * a.scale_log2;                           // (scale > 0)
      }
      x_max[half * 128 + rt] = mx;
      named_bar_sync(1, 256);
      mx = fmaxf(mx, x_max[(half ^ 1) * 128 + rt]);
      if (mx == -INFINITY) mx = 0.f;                      // rows past the sequence end only
      float l = half == 0 ? fast_ex2(s_self - mx) : 0.f;  // own key of a SOS row (0 otherwise)
      const float sc = warp_sos ? 1.f : a.scale_log2;     // pass 2 reads biased scores (SOS warps) or raw ones
      for (int c = c_lo; c < c_hi; ++c) {
        tmem_ld_32x32(t_s + lane_off + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
        const bool tail = !warp_sos && c * 32 + 32 > nkeys;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = fast_ex2(fmaf(__uint_as_float(v[j]), sc, -mx)), p1 = fast_ex2(fmaf(__uint_as_float(v[j + 1]), sc, -mx));
          if (tail) {
            p0 = c * 32 + j < nkeys ? p0 : 0.f;
            p1 = c * 32 + j + 1 < nkeys ? p1 : 0.f;
          }
          l += p0 + p1;
          pk[j >> 1] = pack_half2(p0, p1);
        }
        // keys 32c .. 32c+31 of row rt: key atom c / 2, 16-byte chunks (c & 1) * 4 .. + 3, XOR-swizzled with the row
        const uint32_t base = p_row + (uint32_t)(c >> 1) * ST_TILE;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = (c & 1) * 4 + q4;
          st_shared_v4(base + (uint32_t)((chunk ^ (rt & 7)) << 4), make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]));
        }
      }
      x_sum[half * 128 + rt] = l;
      fence_async_proxy();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      named_bar_sync(2, 256);
      l += x_sum[(half ^ 1) * 128 + rt];

      mbar_wait(o_full, (uint32_t)(t & 1));
      tc_fence_after();
      uint32_t o[32];                                     // this warp's 32 of the 64 output channels
      tmem_ld_32x32(t_o + lane_off + half * 32, o);
      tmem_ld_wait();
      if (ok) {
        const float inv = 1.f / l;
        const float p_self = sos ? fast_ex2(s_self - mx) : 0.f;
        __half* op = a.out + (row0 + row) * W + head * 64 + half * 32;
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(o[d + e]);
          if (sos) {
            const uint4 vu = *reinterpret_cast<const uint4*>(qp + 2 * W + half * 32 + d);
            const __half2* v2 = reinterpret_cast<const __half2*>(&vu);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 vf = __half22float2(v2[e]);
              f[2 * e] = fmaf(p_self, vf.x, f[2 * e]);
              f[2 * e + 1] = fmaf(p_self, vf.y, f[2 * e + 1]);
            }
          }
          uint4 u;
          u.x = pack_half2(f[0] * inv, f[1] * inv); u.y = pack_half2(f[2] * inv, f[3] * inv);
          u.z = pack_half2(f[4] * inv, f[5] * inv); u.w = pack_half2(f[6] * inv, f[7] * inv);
          *reinterpret_cast<uint4*>(op + d) = u;
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ovis
