// Masked flash cross-attention on tcgen05 / TMEM / TMA for the Mask2Former-style decoder
// (reference: CrossAttentionLayer.forward_post -> nn.MultiheadAttention with a bool attn_mask,
//  video_mask2former_transformer_decoder.py:110-122; all-masked-row rule frame_..._decoder.py:87).
//
// Shape of the problem: <= 128 queries per CTA (one UMMA M tile), 8 heads of d = 32 that all read the same key/value
// rows [keys][256] (head h = columns 32h..32h+31), up to ~5e5 keys per group, one shared mask bit per (query, key).
//
// One CTA = (key split, query tile, group).  It walks its key range in 64-key tiles; every tile is processed as four
// "head pairs" (one 128-byte-swizzled 64-channel TMA box of K and of V each), the two heads of a pair being handled
// concurrently by two softmax warpgroups:
//
//   warp 0      TMA producer: Q once (4 boxes, resident), then K/V boxes through a 4-stage ring
//   warp 1      MMA issuer:   S_w = Q_h K_h^T  (UMMA 128x64x16 x2, K-major A and B)      -> TMEM S[w][buf]
//                             O_h += P_w V_h   (UMMA 128x32x16 x4, A = P from smem, B = V MN-major) -> TMEM O[h]
//   warps 2-5   softmax warpgroup A (even heads), warps 6-9 warpgroup B (odd heads): thread = query row;
//               tcgen05.ld S row, expand the mask predicate from the packed sign bits, online softmax in base 2 with
//               lazy rescaling of the TMEM-resident O accumulator, P -> fp16 -> 128B-swizzled smem operand.
//
// TMEM (512 columns): S[2 wg][2 buf] x 64 fp32 columns = 256, O[8 heads] x 32 = 256.
// The split's un-normalised O and (max, sum) go to the same partial buffers as the round-1 kernel and are merged by
// xattn_combine_kernel.
#pragma once
#include "ptx.cuh"

namespace ovis {

struct XattnTcArgs {
  const uint32_t* bits;        // [G][W][q_stride], bit = 1 -> blocked
  const unsigned char* flags;  // [G][q_stride], 1 -> row has an unblocked key
  float* o_part;               // [G][S][8][q_pad][32]
  float* ml_part;              // [G][S][8][q_pad][2]
  int Q, q_pad, q_stride;
  int keys, W, splits, chunk;  // chunk: keys per split, multiple of 64
  long long* trace;            // tools/trace_xattn.py: clock64 stamps of block 0, [6 roles][64 steps][8 events], else null
  const uint32_t* skipmap;     // xattn_tc2 only: [G][qtiles][map_words], bit t = every row of the query tile blocks all 64 keys of tile t; or null
  int map_words;
};

constexpr int XT_KT = 64;                       // keys per tile
constexpr int XT_STAGES = 4;
constexpr int XT_Q_BYTES = 4 * 128 * 128;       // 4 k-blocks [128 rows][64 ch] fp16
constexpr int XT_KV_STAGE = 2 * XT_KT * 128;    // K box + V box
constexpr int XT_P_BYTES = 128 * 128;           // [128 q][64 keys] fp16
constexpr int XT_STATS_BYTES = 256 * 8 * 4;
constexpr int XT_SMEM = XT_Q_BYTES + XT_STAGES * XT_KV_STAGE + 4 * XT_P_BYTES + XT_STATS_BYTES + 1024 + 512;
constexpr int XT_THREADS = 320;

// kind::f16 instruction descriptors (cute::UMMA::InstrDescriptor): fp16 x fp16 -> fp32
__host__ __device__ constexpr uint32_t xt_idesc(int M, int N, int b_mn_major) {
  return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major operand, 128-byte swizzle: rows of 64 fp16 along MN (128 B) per K index, 8-K atoms of 1024 B.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(4096 >> 4) << 16;     // LBO: stride between 64-wide MN atoms (unused: N = 32 <= 64)
  d |= (uint64_t)(1024 >> 4) << 32;     // SBO: stride between 8-K atoms
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float fast_ex2(float x) {      // MUFU.EX2, flush-to-zero: exactly what a probability needs
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(XT_THREADS, 1)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const XattnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + XT_Q_BYTES;
  uint8_t* sP = sKV + XT_STAGES * XT_KV_STAGE;                       // [wg][buf][16 KB]
  float* stats = reinterpret_cast<float*>(sP + 4 * XT_P_BYTES);                   // [256 threads][4 max, 4 sum]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * XT_P_BYTES + XT_STATS_BYTES);
  uint64_t* q_full = bars;               // [1]
  uint64_t* full = bars + 1;             // [4]   TMA -> MMA
  uint64_t* empty = bars + 5;            // [4]   MMA -> TMA
  uint64_t* s_full = bars + 9;           // [2 wg][2 buf]  MMA -> softmax
  uint64_t* s_empty = bars + 13;         // [2][2]         softmax -> MMA (S consumed)
  uint64_t* p_full = bars + 17;          // [2][2]         softmax -> MMA (P written)
  uint64_t* p_empty = bars + 21;         // [2][2]         MMA -> softmax (PV retired)
  uint64_t* done = bars + 25;            // [1]            all MMAs retired
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, qt = blockIdx.y, g = blockIdx.z;
  const int k_begin = split * a.chunk;
  const int k_end = min(k_begin + a.chunk, a.keys);
  const int ntiles = (k_end - k_begin + XT_KT - 1) / XT_KT;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < XT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);        // one elected arrival per softmax warp
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, XT_Q_BYTES);
      for (int j = 0; j < 4; ++j) tma_load_2d(sQ + j * 16384, &tmQ, q_full, j * 64, g * a.Q + qt * 128);
      for (int t = 0; t < ntiles; ++t) {
        const int krow = g * a.keys + k_begin + t * XT_KT;
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&empty[j], (uint32_t)((t & 1) ^ 1));
          uint8_t* st = sKV + j * XT_KV_STAGE;
          mbar_arrive_expect_tx(&full[j], XT_KV_STAGE);
          tma_load_2d(st, &tmK, &full[j], j * 64, krow);
          tma_load_2d(st + XT_KT * 128, &tmV, &full[j], j * 64, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = xt_idesc(128, XT_KT, 0);     // S: M128 N64, A,B K-major
      constexpr uint32_t idesc_o = xt_idesc(128, 32, 1);        // O: M128 N32, B (V) MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t kv_addr = smem_u32(sKV);
      const uint32_t p_addr = smem_u32(sP);
      mbar_wait(q_full, 0);
      tc_fence_after();
      const int nsteps = ntiles * 4;
      // step i = (tile t, pair j): S for step i is issued before PV of step i-1 so the softmax warps never starve
      for (int i = 0; i <= nsteps; ++i) {
        if (i < nsteps) {
          const int t = i >> 2, j = i & 3, b = j & 1;
          mbar_wait(&full[j], (uint32_t)(t & 1));
          tc_fence_after();
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            // S buffer (w, b) was last used by step i-2: wait until that read has finished
            mbar_wait(&s_empty[w * 2 + b], (uint32_t)((((i >> 1) & 1)) ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)((w * 2 + b) * 64);
            const uint64_t adesc = umma_desc_k_sw128(q_addr + j * 16384 + w * 64);
            const uint64_t bdesc = umma_desc_k_sw128(kv_addr + j * XT_KV_STAGE + w * 64);
            umma_f16(d_tmem, adesc, bdesc, idesc_s, 0u);
            umma_f16(d_tmem, adesc + 2, bdesc + 2, idesc_s, 1u);
            umma_commit(&s_full[w * 2 + b]);
          }
        }
        if (i > 0) {
          const int ip = i - 1;
          const int t = ip >> 2, j = ip & 3, b = j & 1;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            mbar_wait(&p_full[w * 2 + b], (uint32_t)((ip >> 1) & 1));
            tc_fence_after();
            const int h = 2 * j + w;
            const uint32_t d_tmem = tmem_base + 256u + (uint32_t)(h * 32);
            const uint64_t adesc = umma_desc_k_sw128(p_addr + (w * 2 + b) * XT_P_BYTES);
            const uint64_t bdesc = umma_desc_mn_sw128(kv_addr + j * XT_KV_STAGE + XT_KT * 128 + w * 64);
#pragma unroll
            for (int kk = 0; kk < XT_KT / 16; ++kk)
              umma_f16(d_tmem, adesc + 2 * kk, bdesc + (uint64_t)(kk * (2048 >> 4)), idesc_o, (t > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&p_empty[w * 2 + b]);
          }
          umma_commit(&empty[j]);           // K/V stage free once both PV products have retired
        }
      }
      umma_commit(done);
    }
  } else {
    // ============================ softmax warpgroups ============================
    const int w = (warp - 2) >> 2;                       // 0: even heads, 1: odd heads
    const int quarter = warp & 3;                        // TMEM lane quarter of this warp
    const int r = quarter * 32 + lane;                   // query row within the tile
    const int q = qt * 128 + r;
    const bool q_ok = q < a.Q;
    const bool use_mask = q_ok && (a.flags[(long long)g * a.q_stride + q] != 0);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t* bits_q = a.bits + (long long)g * a.W * a.q_stride + q;
    uint8_t* p_base = sP + (w * 2) * XT_P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;

    // running (max, sum) of this row for the warpgroup's four heads live in shared memory so that the head-pair loop
    // can stay rolled: fully unrolled, the loop body (4 x ~550 instructions) overflows the instruction cache and the
    // kernel becomes fetch-bound (ncu: stall_no_inst 34 %) whatever the arithmetic does
    float* m_run = stats + (((warp - 2) * 32 + lane) * 8);
    float* l_run = m_run + 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) { m_run[j] = -INFINITY; l_run[j] = 0.f; }

    // mask words of a 64-key tile (shared by all heads); the next tile's words are prefetched one tile ahead
    auto load_words = [&](int t, uint32_t (&dst)[2]) {
      const int kb = k_begin + t * XT_KT;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int wi = (kb >> 5) + x;
        dst[x] = (use_mask && t < ntiles && wi < a.W) ? __ldg(bits_q + (long long)wi * a.q_stride) : 0u;
      }
    };
    uint32_t nw[2];
    load_words(0, nw);
    for (int t = 0; t < ntiles; ++t) {
      const int kb = k_begin + t * XT_KT;
      uint32_t mw[2];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int nvalid = k_end - (kb + x * 32);
        const uint32_t inval = nvalid >= 32 ? 0u : (nvalid <= 0 ? 0xffffffffu : ~((1u << nvalid) - 1u));
        mw[x] = nw[x] | inval;
      }
      load_words(t + 1, nw);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int b = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);       // (step >> 1) & 1 with step = 4t + j
        const int bi = w * 2 + b;
        mbar_wait(&s_full[bi], ph);
        tc_fence_after();
        uint32_t sv[64];
        const uint32_t s_addr = tmem_base + lane_off + (uint32_t)(bi * 64);
        tmem_ld_32x32_nowait(s_addr, sv);
        tmem_ld_32x32_nowait(s_addr + 32, sv + 32);
        tmem_ld_wait();
        reg_fence32(sv);
        reg_fence32(sv + 32);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[bi]);
        // Fused optimistic pass: mask select + max (ALU pipe) interleaved element by element with exp2 against the
        // *stale* reference max (MUFU pipe), so both pipes work at the same time inside one warp.  With lazy rescaling
        // the stale max is the one that will be used anyway unless the tile max exceeds it by more than 2^8 (or this is
        // the row's first unblocked tile); only then the probabilities are recomputed (rare after the first tiles).
        const float m_old = m_run[j];
        const float m_opt = (m_old == -INFINITY) ? 0.f : m_old;
        float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {                     // 8 chunks of 8 keys = 16 bytes
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = c * 8 + e;
            float sc = __uint_as_float(sv[k]);
            if (mw[k >> 5] & (1u << (k & 31))) sc = -INFINITY;
            sv[k] = __float_as_uint(sc);
            mxa[k & 3] = fmaxf(mxa[k & 3], sc);
            p[e] = fast_ex2(sc - m_opt);
            ls[e & 3] += p[e];
          }
          pk[c * 4 + 0] = pack_half2(p[0], p[1]); pk[c * 4 + 1] = pack_half2(p[2], p[3]);
          pk[c * 4 + 2] = pack_half2(p[4], p[5]); pk[c * 4 + 3] = pack_half2(p[6], p[7]);
        }
        const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
        float m_use = m_old;
        bool grow = false;
        if (mx > m_old + 8.f || m_old == -INFINITY) {
          m_use = fmaxf(m_old, mx);
          grow = (m_old != -INFINITY) && (m_use != m_old);
          if (m_use != m_old) {                            // reference changed: recompute this tile's probabilities
            ls[0] = ls[1] = ls[2] = ls[3] = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float p[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                p[e] = fast_ex2(__uint_as_float(sv[c * 8 + e]) - m_use);
                ls[e & 3] += p[e];
              }
              pk[c * 4 + 0] = pack_half2(p[0], p[1]); pk[c * 4 + 1] = pack_half2(p[2], p[3]);
              pk[c * 4 + 2] = pack_half2(p[4], p[5]); pk[c * 4 + 3] = pack_half2(p[6], p[7]);
            }
          }
        }
        const float f = grow ? fast_ex2(m_old - m_use) : 1.f;
        // the P buffer (and the O accumulator of this head) are free once the PV product of step i-2 has retired
        mbar_wait(&p_empty[bi], ph ^ 1u);
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          uint32_t ov[32];
          const uint32_t o_addr = tmem_base + lane_off + 256u + (uint32_t)((2 * j + w) * 32);
          tmem_ld_32x32_nowait(o_addr, ov);
          tmem_ld_wait();
          reg_fence32(ov);
#pragma unroll
          for (int c = 0; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * f);
          tmem_st_32x32(o_addr, ov);
          tmem_st_wait();
        }
        m_run[j] = m_use;
        l_run[j] = l_run[j] * f + ((ls[0] + ls[1]) + (ls[2] + ls[3]));
        uint8_t* prow = p_base + b * XT_P_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        fence_async_proxy();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[bi]);
      }
    }
    // ---- all MMAs retired: write the split's partial O (un-normalised) and (max, sum)
    mbar_wait(done, 0);
    tc_fence_after();
    const long long pb = (((long long)g * a.splits + split) * 8) * a.q_pad + q;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int h = 2 * j + w;
      uint32_t ov[32];
      tmem_ld_32x32_nowait(tmem_base + lane_off + 256u + (uint32_t)(h * 32), ov);
      tmem_ld_wait();
      reg_fence32(ov);
      if (q < a.q_pad) {
        float* op = a.o_part + (pb + (long long)h * a.q_pad) * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          *reinterpret_cast<float4*>(op + c) = make_float4(__uint_as_float(ov[c]), __uint_as_float(ov[c + 1]),
                                                           __uint_as_float(ov[c + 2]), __uint_as_float(ov[c + 3]));
        *reinterpret_cast<float2*>(a.ml_part + (pb + (long long)h * a.q_pad) * 2) = make_float2(m_run[j], l_run[j]);
      }
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
