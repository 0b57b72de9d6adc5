// Masked flash cross-attention on tcgen05 / TMEM / TMA, second layout (reference: CrossAttentionLayer.forward_post ->
// nn.MultiheadAttention with a bool attn_mask, video_mask2former_transformer_decoder.py:110-122; all-masked-row rule
// frame_mask2former_transformer_decoder.py:87).
//
// What the first layout (xattn_tc.cuh) taught (profiles/experiments/xattn_tc_r1_findings.md + tools/ubench/softmax_loop.cu):
// the softmax instruction stream alone runs at the MUFU.EX2 rate (~14 elements/clk/SM, 130 us for the cfg-2 level-2
// launch) once 16 warps per SM execute it, but the kernel needed 350 us because only 8 softmax warps were resident and
// each of them serialises TMEM load -> mask/max -> exp2 -> store -> fence -> arrive.  So this layout is built for
// thread-level parallelism and a deep K/V ring instead of a smarter inner loop:
//
//   CTA = (key chunk, head PAIR, query tile, group): Q is one 16 KB box, a 64-key K/V stage is 16 KB, so the ring is
//         8 stages deep (three tiles in flight cover the HBM latency at the target rate of ~0.6 us per tile);
//   four softmax warpgroups (16 warps, 96 registers: 5 warps on one scheduler cap it): warpgroup (w, b) owns head w of the pair and the key tiles of
//         parity b, with its own S buffer, P buffer, O accumulator and running (max, sum) -- the two parities of a head
//         are simply two interleaved key splits, merged with the others by xattn_combine_kernel;
//   three single-thread roles on their own warps: TMA producer, S = Q K^T issuer, O += P V issuer; the issuers serve
//         the warpgroups out of order, so S for a warpgroup's next tile is issued the moment the warpgroup has drained its S buffer into registers.
//
// TMEM (512 columns allocated): S[4 wg] x 64 fp32 columns at 0, O[4 wg] x 32 at 256.
#pragma once
#include "ptx.cuh"
#include "xattn_tc.cuh"

namespace ovis {

constexpr int X2_KT = 64;                        // keys per tile
constexpr int X2_STAGES = 8;
constexpr int X2_Q_BYTES = 128 * 128;            // [128 rows][64 ch] fp16 (one head pair)
constexpr int X2_KV_STAGE = 2 * X2_KT * 128;     // K box + V box
constexpr int X2_P_BYTES = 128 * 128;            // [128 q][64 keys] fp16
// Tile skipping ("skips fully-masked key tiles"): xattn_skipmap_kernel marks the 64-key tiles that every row of a query
// tile blocks (rows under the all-masked-row rule block nothing).  Each CTA then builds the list of the tiles it has to
// visit -- its equal share of the *surviving* tiles of the group, so the chunks stay balanced whatever the mask looks
// like -- and every role walks that list instead of a contiguous key range: skipped tiles cost no TMA load, no MMA, no
// softmax step.  With nothing to skip the list is the old contiguous chunk (dense masks: +1 small launch per call).
constexpr int X2_MAP_WORDS = 512;                // skip bitmap words per (group, query tile): up to 16384 key tiles
constexpr int X2_LIST_MAX = 7168;                // tile list entries per CTA (uint16)
constexpr int X2_LIST_BYTES = X2_MAP_WORDS * 4 + X2_LIST_MAX * 2;
constexpr int X2_SMEM = X2_Q_BYTES + X2_STAGES * X2_KV_STAGE + 4 * X2_P_BYTES + 1024 + 512 + X2_LIST_BYTES;
constexpr int X2_THREADS = 19 * 32;

// clock64 timeline of block 0 (tools/trace_xattn.py): compiled in only with -DOVIS_XATTN_TRACE_BUILD
#ifdef OVIS_XATTN_TRACE_BUILD
#define X2_TRACE(role, step, ev)                                                                   \
  do {                                                                                            \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (step) < 64)          \
      a.trace[((role) * 64 + (step)) * 8 + (ev)] = clock64();                                     \
  } while (0)
#else
#define X2_TRACE(role, step, ev) do { } while (0)
#endif

// bit t of map[g][qt][t / 32] = every row of query tile qt blocks all keys of the 64-key tile t.  A row under the
// all-masked-row rule (flag 0: it attends to every key, frame_..._decoder.py:87) blocks nothing; rows past Q do not count.
// grid (map words, query tiles, G), 128 threads = one per row of the query tile.
__global__ void __launch_bounds__(128)
xattn_skipmap_kernel(const uint32_t* __restrict__ bits, const unsigned char* __restrict__ flags, uint32_t* __restrict__ map,
                     int Q, int q_stride, int keys, int W, int map_words) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  __shared__ uint32_t s_and[4];
  const int wi = blockIdx.x, qt = blockIdx.y, g = blockIdx.z;
  const int q = qt * 128 + threadIdx.x;
  const bool counts = q < Q;
  const bool masked = counts && flags[(long long)g * q_stride + q] != 0;
  const uint32_t* bq = bits + (long long)g * W * q_stride + q;
  // bit j of `mine`: this row blocks every key of tile wi * 32 + j.  The 64 mask words are fetched in four batches of 16
  // independent loads (a load -> test -> barrier loop per tile made this kernel 25 us long).
  uint32_t mine = masked ? 0u : (counts ? 0u : 0xffffffffu);
  if (masked) {
#pragma unroll
    for (int b4 = 0; b4 < 4; ++b4) {
      uint32_t wv[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const int w = (wi * 32 + b4 * 8) * 2 + x;
        const int nvalid = keys - w * 32;
        const uint32_t inval = nvalid >= 32 ? 0u : (nvalid <= 0 ? 0xffffffffu : ~((1u << nvalid) - 1u));
        wv[x] = (w < W ? __ldg(bq + (long long)w * q_stride) : 0u) | inval;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if ((wv[2 * j] & wv[2 * j + 1]) == 0xffffffffu) mine |= 1u << (b4 * 8 + j);
    }
  }
  mine = __reduce_and_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0) s_and[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t out = s_and[0] & s_and[1] & s_and[2] & s_and[3];
    const int total_tiles = (keys + X2_KT - 1) / X2_KT;
    const int left = total_tiles - wi * 32;                     // tiles covered by this word
    if (left < 32) out &= left <= 0 ? 0u : ((1u << left) - 1u);
    map[((long long)g * gridDim.y + qt) * map_words + wi] = out;
  }
}

__global__ void __launch_bounds__(X2_THREADS, 1)
xattn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const XattnTcArgs a) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + X2_Q_BYTES;
  uint8_t* sP = sKV + X2_STAGES * X2_KV_STAGE;                       // [wg][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * X2_P_BYTES);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* full = bars + 1;                    // [8]  TMA -> S issuer
  uint64_t* empty = bars + 9;                   // [8]  PV issuer -> TMA
  uint64_t* s_full = bars + 17;                 // [4 wg]  S issuer -> softmax
  uint64_t* s_empty = bars + 21;                // [4]     softmax -> S issuer (S drained into registers)
  uint64_t* p_full = bars + 25;                 // [4]     softmax -> PV issuer (P written)
  uint64_t* p_empty = bars + 29;                // [4]     PV issuer -> softmax (PV retired: P buffer and O free)
  uint64_t* done = bars + 33;                   // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 34);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk_id = blockIdx.x >> 2, hp = blockIdx.x & 3, qt = blockIdx.y, g = blockIdx.z;
  uint32_t* wpre = reinterpret_cast<uint32_t*>(sP + 4 * X2_P_BYTES + 512);      // [X2_MAP_WORDS] prefix of surviving tiles
  uint16_t* tlist = reinterpret_cast<uint16_t*>(wpre + X2_MAP_WORDS);           // [X2_LIST_MAX] this CTA's tiles
  bool use_list = a.skipmap != nullptr;
  int k_begin = chunk_id * a.chunk;
  int k_end = min(k_begin + a.chunk, a.keys);
  int ntiles = (k_end - k_begin + X2_KT - 1) / X2_KT;
  if (use_list) {
    // surviving tiles of this (group, query tile): per-word counts -> exclusive prefix -> this chunk's share -> list
    const uint32_t* map = a.skipmap + ((long long)g * gridDim.y + qt) * a.map_words;
    const int total_tiles = (a.keys + X2_KT - 1) / X2_KT;
    const int words = (total_tiles + 31) >> 5;
    // (xattn_skipmap_kernel leaves the bits past the last tile clear, so ~word over-counts there: masked off here)
    const uint32_t last_mask = (total_tiles & 31) ? ((1u << (total_tiles & 31)) - 1u) : 0xffffffffu;
    for (int i = threadIdx.x; i < X2_MAP_WORDS; i += X2_THREADS)
      wpre[i] = i < words ? (uint32_t)__popc(~__ldg(map + i) & (i == words - 1 ? last_mask : 0xffffffffu)) : 0u;
    __syncthreads();
    if (warp == 0) {                                     // exclusive scan of 512 counts: 16 per lane
      uint32_t loc[16], sum = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) { loc[j] = wpre[lane * 16 + j]; sum += loc[j]; }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      uint32_t run = inc - sum;
#pragma unroll
      for (int j = 0; j < 16; ++j) { wpre[lane * 16 + j] = run; run += loc[j]; }
      if (lane == 31) tmem_holder[1] = inc;                               // total number of surviving tiles
    }
    __syncthreads();
    const int cnt = (int)tmem_holder[1];
    const int chunks = (int)(gridDim.x >> 2);
    const int per = (cnt + chunks - 1) / chunks;
    const int lo = chunk_id * per, hi = min(cnt, lo + per);
    ntiles = max(0, hi - lo);
    k_end = a.keys;
    if (cnt == total_tiles) {             // nothing to skip: the share is a contiguous run of tiles, no list needed
      use_list = false;
      k_begin = lo * X2_KT;
    } else {
      k_begin = 0;
    }
    for (int i = threadIdx.x; use_list && i < words; i += X2_THREADS) {
      uint32_t w = ~__ldg(map + i) & (i == words - 1 ? last_mask : 0xffffffffu);
      int r = (int)wpre[i];
      if (r >= hi || r + __popc(w) <= lo) continue;
      while (w) {
        const int bit = __ffs(w) - 1;
        w &= w - 1u;
        if (r >= lo && r < hi) tlist[r - lo] = (uint16_t)(i * 32 + bit);
        ++r;
      }
    }
    // (the __syncthreads() after the barrier / TMEM set-up below publishes the list)
  }
  const bool listed = use_list;
  const int k_first = k_begin, k_last = k_end;      // (by-value copies: the roles' loops keep them in registers)
  auto tile_key = [=](int t) -> int { return listed ? (int)tlist[t] * X2_KT : k_first + t * X2_KT; };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < X2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);        // one elected arrival per softmax warp
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc(tmem_holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 16) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, X2_Q_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, hp * 64, g * a.Q + qt * 128);
      for (int t = 0; t < ntiles; ++t) {
        const int st = t % X2_STAGES;
        mbar_wait(&empty[st], (uint32_t)(((t / X2_STAGES) & 1) ^ 1));
        uint8_t* dst = sKV + st * X2_KV_STAGE;
        const int krow = g * a.keys + tile_key(t);
        mbar_arrive_expect_tx(&full[st], X2_KV_STAGE);
        tma_load_2d(dst, &tmK, &full[st], hp * 64, krow);
        tma_load_2d(dst + X2_KT * 128, &tmV, &full[st], hp * 64, krow);
      }
    }
  } else if (warp == 17) {
    // ============================ S = Q K^T issuer ============================
    // Serves the four warpgroups out of order: whichever has drained its S buffer (and whose K tile has landed) gets
    // its next S first, so one slow warpgroup never delays the others.
    if (lane == 0) {
      constexpr uint32_t idesc_s = xt_idesc(128, X2_KT, 0);     // M128 N64, A and B K-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t kv_addr = smem_u32(sKV);
      mbar_wait(q_full, 0);
      int nxt[4];                          // per warpgroup: local index n of its next tile (t = 2n + b)
#pragma unroll
      for (int i = 0; i < 4; ++i) nxt[i] = 0;
      int remaining = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) remaining += (ntiles - (i >> 1) + 1) >> 1;
      while (remaining > 0) {
        bool any = false;
#pragma unroll
        for (int wg = 0; wg < 4; ++wg) {
          const int w = wg & 1, b = wg >> 1;
          const int n = nxt[wg], t = 2 * n + b;
          if (t >= ntiles) continue;
          const int st = t % X2_STAGES;
          if (!mbar_test_wait(&s_empty[wg], (uint32_t)((n & 1) ^ 1))) continue;
          if (!mbar_test_wait(&full[st], (uint32_t)((t / X2_STAGES) & 1))) continue;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(wg * 64);
          const uint64_t adesc = umma_desc_k_sw128(q_addr + w * 64);
          const uint64_t bdesc = umma_desc_k_sw128(kv_addr + st * X2_KV_STAGE + w * 64);
          umma_f16(d_tmem, adesc, bdesc, idesc_s, 0u);
          umma_f16(d_tmem, adesc + 2, bdesc + 2, idesc_s, 1u);
          umma_commit(&s_full[wg]);
          X2_TRACE(4, n, wg);                                  // role 4 = S issuer: event index = warpgroup
          nxt[wg] = n + 1;
          --remaining;
          any = true;
        }
        if (!any) {
          // nothing ready: sleep on the barrier of the warpgroup that is furthest behind (mbarrier.try_wait suspends
          // the thread, so the issuer does not take issue slots from the softmax warps of its scheduler)
          int wmin = -1, tmin = 0x7fffffff;
#pragma unroll
          for (int wg = 0; wg < 4; ++wg) {
            const int t = 2 * nxt[wg] + (wg >> 1);
            if (t < ntiles && t < tmin) { tmin = t; wmin = wg; }
          }
          if (wmin >= 0) {
            const int n = tmin >> 1;
            if (mbar_try_wait(&full[tmin % X2_STAGES], (uint32_t)((tmin / X2_STAGES) & 1)))
              mbar_try_wait(&s_empty[wmin], (uint32_t)((n & 1) ^ 1));
          }
        }
      }
    }
  } else if (warp == 18) {
    // ============================ O += P V issuer (out of order as well) ============================
    if (lane == 0) {
      constexpr uint32_t idesc_o = xt_idesc(128, 32, 1);        // M128 N32, B (V) MN-major
      const uint32_t kv_addr = smem_u32(sKV);
      const uint32_t p_addr = smem_u32(sP);
      int nxt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) nxt[i] = 0;
      int remaining = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) remaining += (ntiles - (i >> 1) + 1) >> 1;
      uint32_t half_done = 0;              // bit st: one of the two heads of the tile in stage st has been issued
      while (remaining > 0) {
        bool any = false;
#pragma unroll
        for (int wg = 0; wg < 4; ++wg) {
          const int w = wg & 1, b = wg >> 1;
          const int n = nxt[wg], t = 2 * n + b;
          if (t >= ntiles) continue;
          if (!mbar_test_wait(&p_full[wg], (uint32_t)(n & 1))) continue;
          tc_fence_after();
          const int st = t % X2_STAGES;
          const uint32_t d_tmem = tmem_base + 256u + (uint32_t)(wg * 32);
          const uint64_t adesc = umma_desc_k_sw128(p_addr + wg * X2_P_BYTES);
          const uint64_t bdesc = umma_desc_mn_sw128(kv_addr + st * X2_KV_STAGE + X2_KT * 128 + w * 64);
#pragma unroll
          for (int kk = 0; kk < X2_KT / 16; ++kk)
            umma_f16(d_tmem, adesc + 2 * kk, bdesc + (uint64_t)(kk * (2048 >> 4)), idesc_o, (n > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&p_empty[wg]);
          X2_TRACE(5, n, wg);                                  // role 5 = PV issuer
          if (half_done & (1u << st)) {     // both heads of this tile issued: the K/V stage is free once they retire
            umma_commit(&empty[st]);
            half_done &= ~(1u << st);
          } else {
            half_done |= 1u << st;
          }
          nxt[wg] = n + 1;
          --remaining;
          any = true;
        }
        if (!any) {
          int wmin = -1, tmin = 0x7fffffff;
#pragma unroll
          for (int wg = 0; wg < 4; ++wg) {
            const int t = 2 * nxt[wg] + (wg >> 1);
            if (t < ntiles && t < tmin) { tmin = t; wmin = wg; }
          }
          if (wmin >= 0) mbar_try_wait(&p_full[wmin], (uint32_t)((tmin >> 1) & 1));
        }
      }
      umma_commit(done);
    }
  } else {
    // ============================ softmax warpgroups ============================
    const int wg = warp >> 2;                            // 0..3
    const int w = wg & 1, b = wg >> 1;                   // head within the pair, key-tile parity
    const int quarter = warp & 3;                        // TMEM lane quarter of this warp
    const int r = quarter * 32 + lane;                   // query row within the tile
    const int q = qt * 128 + r;
    const bool use_mask = (q < a.Q) && (a.flags[(long long)g * a.q_stride + q] != 0);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t* bits_q = a.bits + (long long)g * a.W * a.q_stride + q;
    uint8_t* prow = sP + wg * X2_P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
    const uint32_t s_addr = tmem_base + lane_off + (uint32_t)(wg * 64);
    const uint32_t o_addr = tmem_base + lane_off + 256u + (uint32_t)(wg * 32);
    float m_run = -INFINITY, l_run = 0.f;

    auto load_words = [&](int t, uint32_t (&dst)[2]) {
      const int kb = t < ntiles ? tile_key(t) : 0;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int wi = (kb >> 5) + x;
        dst[x] = (use_mask && t < ntiles && wi < a.W) ? __ldg(bits_q + (long long)wi * a.q_stride) : 0u;
      }
    };
    uint32_t nw[2];
    load_words(b, nw);
    for (int t = b, n = 0; t < ntiles; t += 2, ++n) {
      const int kb = tile_key(t);
      uint32_t mw[2] = {nw[0], nw[1]};
      if (kb + X2_KT > k_last) {                          // only the last tile of the group can have keys past the end
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int nvalid = k_last - (kb + x * 32);
          const uint32_t inval = nvalid >= 32 ? 0u : (nvalid <= 0 ? 0xffffffffu : ~((1u << nvalid) - 1u));
          mw[x] |= inval;
        }
      }
      load_words(t + 2, nw);
      const uint32_t ph = (uint32_t)(n & 1);
      const bool tr = (quarter == 0 && lane == 0);
      if (tr) X2_TRACE(wg, n, 0);                           // step begins
      mbar_wait(&s_full[wg], ph);
      tc_fence_after();
      if (tr) X2_TRACE(wg, n, 1);                           // S available
      // Online softmax in base 2 with a lazily updated reference max (one 32-key half at a time to stay inside the
      // register budget): probabilities are computed optimistically against the running reference; only when a half
      // would overflow (or the row had no unblocked key so far) it is recomputed against a new reference and
      // everything accumulated so far is rescaled.
      uint32_t pk[32];
      float m_cur = m_run;
      float f_all = 1.f;                                  // rescale factor for l_run / O from before this tile
      bool grow_any = false;
      float sum_tile = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        // 32 S columns at a time (registers: 19 warps leave 96 per thread); the S buffer is handed back to the issuer
        // as soon as the second half sits in registers.  Keep this loop unrolled and branch-free up to the TMEM load:
        // ptxas then overlaps the second half's load with the first half's arithmetic.  A warp-uniform "whole half is
        // masked -> skip" test in front of the load cost 17 % on dense masks (837 -> 981 us), and a rolled version
        // (one copy of the half, P stored per half) 12 % (835 -> 935 us); neither is instruction-cache related.
        uint32_t sv[32];
        const uint32_t word = mw[hf];
        __syncwarp();                                     // (re-converged after the per-lane recompute branch)
        tmem_ld_32x32_nowait(s_addr + hf * 32, sv);
        tmem_ld_wait();
        reg_fence32(sv);
        if (hf == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[wg]);
        }
        const float m_opt = (m_cur == -INFINITY) ? 0.f : m_cur;
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = e + u;
            float sc = __uint_as_float(sv[k]);
            if (word & (1u << k)) sc = -INFINITY;
            p[u] = fast_ex2(sc - m_opt);
            ls[k & 3] += p[u];
          }
          pk[hf * 16 + (e >> 1)] = pack_half2(p[0], p[1]);
        }
        // No running max in the fast path: the half's sum bounds every probability (all are >= 0), so while it stays
        // <= 2^12 nothing can overflow fp16 and the stale reference is as good as the true max.  Otherwise (or for the
        // row's first unblocked keys) find the max, move the reference and redo the half.
        const float hsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
        if (!(hsum <= 4096.f) || (m_cur == -INFINITY && word != 0xffffffffu)) {
          float mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float sc = __uint_as_float(sv[e]);
            if (word & (1u << e)) sc = -INFINITY;
            sv[e] = __float_as_uint(sc);
            mx = fmaxf(mx, sc);
          }
          const float m_new = fmaxf(m_cur, mx);
          ls[0] = ls[1] = ls[2] = ls[3] = 0.f;
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = fast_ex2(__uint_as_float(sv[e]) - m_new);
            const float p1 = fast_ex2(__uint_as_float(sv[e + 1]) - m_new);
            ls[e & 3] += p0;
            ls[(e + 1) & 3] += p1;
            pk[hf * 16 + (e >> 1)] = pack_half2(p0, p1);
          }
          if (m_cur != -INFINITY) {
            const float f = fast_ex2(m_cur - m_new);
            f_all *= f;
            grow_any = true;
            sum_tile *= f;
            if (hf == 1) {                                // first half was packed against the old reference
              const __half2 fh = __float2half2_rn(f);
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                __half2 v = *reinterpret_cast<__half2*>(&pk[c]);
                v = __hmul2(v, fh);
                pk[c] = *reinterpret_cast<uint32_t*>(&v);
              }
            }
          }
          m_cur = m_new;
        }
        sum_tile += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      }
      if (tr) X2_TRACE(wg, n, 2);                           // softmax arithmetic done
      // the P buffer and the O accumulator of this warpgroup are free once the PV product of its previous tile retired
      mbar_wait(&p_empty[wg], ph ^ 1u);
      if (tr) X2_TRACE(wg, n, 3);                           // P buffer free
      tc_fence_after();
      if (__any_sync(0xffffffffu, grow_any)) {
        uint32_t ov[32];
        tmem_ld_32x32_nowait(o_addr, ov);
        tmem_ld_wait();
        reg_fence32(ov);
#pragma unroll
        for (int c = 0; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * f_all);
        tmem_st_32x32(o_addr, ov);
        tmem_st_wait();
      }
      m_run = m_cur;
      l_run = l_run * f_all + sum_tile;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      fence_async_proxy();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[wg]);
      if (tr) X2_TRACE(wg, n, 4);                           // P handed over
    }
    // ---- all MMAs retired: write this (head, parity) partial: un-normalised O and (max, sum)
    mbar_wait(done, 0);
    tc_fence_after();
    const int h = 2 * hp + w;
    const int split = 2 * chunk_id + b;
    const long long pr = ((((long long)g * a.splits + split) * 8) + h) * a.q_pad + q;
    uint32_t ov[32];
    if (b < ntiles) {                                   // (a warpgroup without tiles never had its accumulator written)
      tmem_ld_32x32_nowait(o_addr, ov);
      tmem_ld_wait();
      reg_fence32(ov);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) ov[c] = 0u;
    }
    if (q < a.q_pad) {
      float* op = a.o_part + pr * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(op + c) = make_float4(__uint_as_float(ov[c]), __uint_as_float(ov[c + 1]),
                                                         __uint_as_float(ov[c + 2]), __uint_as_float(ov[c + 3]));
      *reinterpret_cast<float2*>(a.ml_part + pr * 2) = make_float2(m_run, l_run);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ovis
