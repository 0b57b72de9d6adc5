// Masked flash cross-attention on tcgen05 / TMEM / TMA, third layout: TRANSPOSED scores (reference:
// CrossAttentionLayer.forward_post -> nn.MultiheadAttention with a bool attn_mask,
// video_mask2former_transformer_decoder.py:110-122; all-masked-row rule frame_mask2former_transformer_decoder.py:87).
//
// Why transposed.  xattn_tc2 (thread = query row) executes ~9 instructions per probability (mask select, subtract, exp,
// running sum, pack, and ~4 of per-step overhead on 64-key steps) and pads Q = 100 queries to a 128-row UMMA tile; it sits
// at ~60 % of the exp-pipe (MUFU.EX2) bound.  Here S^T = K Q^T is computed instead (UMMA M = 128 keys, N = the queries
// rounded up to 16: 112 for Q = 100), a thread owns one KEY and walks the query columns:
//   * the softmax reference is folded into the score product: one extra K = 16 slab multiplies a constant "ones" operand
//     (one 256-byte SWIZZLE_32B atom broadcast to all 128 rows with SBO = 0) with a per-query [-m_ref] column, so the
//     accumulator already holds s - m_ref and the exponential needs no subtraction;
//   * the row sums come from the tensor core as well: the B operand of O += P V is [V_h | 1] (the head's V box plus a
//     constant "ones" atom), which leaves sum_k p in column 32 of the 48-column accumulator;
//   * what is left per probability: mask select (R2P + FSEL), MUFU.EX2, half a saturating pack, 1/8 of a 16-byte store;
//   * 12.5 % fewer exponentials (112 instead of 128 columns) and 128-key steps (half the per-step fixed cost).
// The reference is per (CTA, head, query) and constant while the chunk is processed, so the softmax warpgroups
// accumulate into ONE [O | l] accumulator per head and a CTA emits one partial per head (merged by xattn_combine_kernel).
//
// Robustness without a running max.  m_ref comes from a "max pass" over the FIRST key tile of the chunk (masked column
// maxima by warp-wide CREDUX.MAX + shared-memory atomics) plus a margin of 2 (probabilities of that tile <= 2^-2; for a
// query with no unblocked key in that tile: the unmasked maximum - 2).  Probabilities are packed with saturation, so
// nothing becomes inf / NaN, and after the chunk the row sums l (exact fp32 sums of the fp16 probabilities the PV product
// used) tell whether the reference was good: l >= 2^15 may hide a saturated probability (s - m_ref >= 16), l < 2^-3 means
// the probabilities sat in fp16's subnormal range (or all underflowed: l = 0 although unblocked keys exist).  Such a row
// gets a new reference FROM ITS SUM, m_ref += log2(l) - 3 (l = 0: m_ref -= 22), and the CTA runs the exponential pass
// again (rows that were fine keep their reference and reproduce their result); every retry moves a bad row by at least 12
// binary orders towards its window, six retries cover any fp16-representable score range.  With the dense ~50 % masks of
// the benchmark no CTA retries; with very sparse masks a CTA typically retries once.
//
// CTA = (key chunk, head PAIR, query tile of <= 128, group), 12 warps:
//   warps 0-7   two softmax warpgroups (thread = key row of a 128-key tile); unit u = 2 * tile + head goes to warpgroup
//               u % 2, its S^T to TMEM buffer u % 3 and its P^T to shared-memory buffer u % 3: with three rotating buffers
//               for two consumers the scores of a warpgroup's next unit are always ready and its P^T buffer always free,
//               and the chunks of consecutive units are processed as one stream (the TMEM load of the next 16 columns --
//               also across a unit boundary -- is in flight while the current ones are computed)
//   warp 8      TMA producer for Q (once) and the K ring (2 x 16 KB: a K stage is free as soon as both heads' S^T are issued)
//   warp 11     TMA producer for the V ring (3 x 16 KB, two 64B-swizzled per-head boxes per stage); both producers pull
//               their tiles into L2 six tiles ahead (cp.async.bulk.prefetch.tensor)
//   warp 9      S^T issuer: 2 x UMMA 128xNx16 (K-major K and Q) + the reference slab
//   warp 10     PV issuer: [O_h | l_h] += P [V_h | 1] (8 x UMMA 128x48x16, A = P^T MN-major SWIZZLE_128B with two 64-wide
//               atoms; B MN-major SWIZZLE_64B with two 32-wide atoms: the head's V box and a constant atom whose column 0
//               is 1, so the row sums cost no second pass over P^T)
// TMEM (512 columns): S^T[3] x 128 fp32 columns at 0, [O | l][2 heads] x 48 at 384 (column 32 = row sums).
//
// Measured on B200 (tools/prof_xattn_t.py, tools/ubench/umma_rate.cu, profiles/experiments/xattn_tc3_r2_findings.md):
//   * a UMMA 128xNx16 occupies the tensor pipe for max(56, N / 2) cycles whatever the operand layouts, so the eight
//     N = 48 PV products of a unit cost 448 cycles and the three S^T products 180: 628 of the ~1400 cycles a unit takes;
//   * the exponentials need 896 cycles per unit on the MUFU pipe, the K / V stream ~800 at the HBM peak (the pipeline
//     with all arithmetic removed runs at 5.6 TB/s);
//   * config-2 level-2 launch (1 clip, 529 920 keys): 195 us (xattn_tc2: 253); four clips: 669 us (941).
#pragma once
#include <type_traits>
#include "ptx.cuh"
#include "xattn_tc.cuh"

namespace ovis {

struct XattnT3Args {
  const uint32_t* bits_t;      // [G][keys][qw]: bit q%32 of word q/32 = key blocked for query q
  const unsigned char* flags;  // [G][q_stride], 1 -> row has an unblocked key (0: the row attends to every key)
  float* o_part;               // [G][S][8][q_pad][32]
  float* ml_part;              // [G][S][8][q_pad][2]
  int Q, q_pad, q_stride, qw;
  int keys, splits, chunk;     // splits = chunks; chunk: keys per chunk, multiple of 128
  const uint32_t* skipmap;     // [G][qtiles][map_words]: bit t = every (masked) row of the query tile blocks all 128 keys of tile t; or null
  int map_words;
  int* stats;                  // optional [2]: CTAs launched, CTAs that took the retry path (tests / profiling); or null
  long long* trace;            // tools/trace_xattn_t.py: clock64 stamps of block 0, [7 roles][128 steps][8 events]; or null
};

#ifdef OVIS_XATTN_TRACE_BUILD
#define X3_TRACE(role, step, ev)                                                                   \
  do {                                                                                            \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (step) < 128)         \
      a.trace[((role) * 128 + (step)) * 8 + (ev)] = clock64();                                    \
  } while (0)
#else
#define X3_TRACE(role, step, ev) do { } while (0)
#endif

constexpr int X3_KT = 128;                        // keys per tile
constexpr int X3_KS = 2, X3_VS = 3;               // ring depths (K stages turn over as soon as both heads' S^T are issued)
#ifndef X3_NWG_DEF
#define X3_NWG_DEF 2
#endif
constexpr int X3_NWG = X3_NWG_DEF;                 // softmax warpgroups (warps 0-7)
constexpr int X3_NB = 3;                           // S^T buffers in TMEM = P^T buffers in shared memory, rotating with the unit number
constexpr int X3_WARP_K = 4 * X3_NWG, X3_WARP_S = X3_WARP_K + 1, X3_WARP_PV = X3_WARP_K + 2, X3_WARP_V = X3_WARP_K + 3;
constexpr int X3_PF = 6;                           // tiles the TMA producers prefetch into L2 ahead of their loads
constexpr int X3_THREADS = (4 * X3_NWG + 4) * 32; // 384
constexpr int X3_TILE_BYTES = X3_KT * 128;        // one K or V box: [128 keys][64 ch] fp16
constexpr int X3_P_BYTES = 2 * 16 * 1024;         // P^T: [2 query atoms][16 key groups][8 keys][128 B]
constexpr int X3_MAP_WORDS = 256;                 // skip bitmap words per (group, query tile): up to 8192 key tiles
constexpr int X3_LIST_MAX = 2048;                 // tile list entries per CTA (uint16)
// shared memory map (offsets from the 1024-aligned base)
constexpr int X3_OFF_Q = 0;
constexpr int X3_OFF_K = X3_OFF_Q + 16384;
constexpr int X3_OFF_V = X3_OFF_K + X3_KS * X3_TILE_BYTES;
constexpr int X3_OFF_P = X3_OFF_V + X3_VS * X3_TILE_BYTES;
constexpr int X3_OFF_MREF = X3_OFF_P + X3_NB * X3_P_BYTES;       // 2 heads x [128 q][32 B] SWIZZLE_32B
constexpr int X3_OFF_E = X3_OFF_MREF + 2 * 4096;                 // [128 keys][64 B] SWIZZLE_64B MN-major, column 0 = 1: the "ones" atom of V
constexpr int X3_OFF_ONES = X3_OFF_E + 8192;                     // [8 rows][32 B] SWIZZLE_32B, column 0 = 1
constexpr int X3_OFF_SMAX = X3_OFF_ONES + 256;                   // int [2 kinds][2 heads][128]: ordered-int column maxima
constexpr int X3_OFF_REF = X3_OFF_SMAX + 2 * 2 * 128 * 4;        // float [2 heads][128]: the reference in use
constexpr int X3_OFF_BAR = X3_OFF_REF + 2 * 128 * 4;             // 32 mbarriers + control words
constexpr int X3_OFF_LIST = X3_OFF_BAR + 512;                    // prefix words + tile list
constexpr int X3_SMEM = X3_OFF_LIST + X3_MAP_WORDS * 4 + X3_LIST_MAX * 2 + 1024 /*align slack*/;
static_assert(X3_SMEM <= 232448, "xattn_tc3 shared memory");

__device__ __forceinline__ uint64_t x3_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;        // 0 none, 2 SWIZZLE_128B, 4 SWIZZLE_64B, 6 SWIZZLE_32B
  return d;
}
__host__ __device__ constexpr uint32_t x3_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld_32x16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void reg_fence16(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {     // overflow clamps to 65504 instead of inf
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 2^x on the FMA / ALU pipes (the exp pipe, MUFU.EX2, is this kernel's binding unit: 16 results per clock and SM, while
// the FMA pipe has cycles to spare): round-to-nearest split x = n + f with the 1.5 * 2^23 trick, cubic minimax polynomial
// for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, a sixth of an fp16 half-ulp), exponent added as an integer.
// -inf (masked) clamps to 2^-126, which packs to 0.
__device__ __forceinline__ float poly_ex2(float x) {
  x = fminf(fmaxf(x, -126.f), 126.f);
  const float rr = x + 12582912.f;
  const float f = x - (rr - 12582912.f);
  float p = 0.0551716685f;
  p = fmaf(p, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(rr) << 23));
}
#ifndef X3_POLY_MOD
#define X3_POLY_MOD 0        // every X3_POLY_MOD-th exponential of a thread on the FMA pipe; 0 = all on MUFU (measured: 4 -> +2 %, 3 -> +5 %, 2 -> +12 % TIME, the loop is latency-bound, not MUFU-bound; profiles/experiments)
#endif
__device__ __forceinline__ float warp_max_f32(float x) {
  float y;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ int ordered_int(float f) {            // monotone float -> int map (for shared-memory atomicMax)
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// bit t of map[g][qt][t / 32] = every row of query tile qt that uses its mask blocks all 128 keys of tile t (and no row of
// the tile is under the all-masked-row rule, which attends to everything).  Built from the 32-key block ANDs the mask
// GEMM's epilogue wrote (EPI_SIGNBITS_T): grid (map words, query tiles, G), 32 threads = one key tile each.
__global__ void __launch_bounds__(32)
xattn_t3_skipmap_kernel(const uint32_t* __restrict__ blockand, const unsigned char* __restrict__ flags, uint32_t* __restrict__ map,
                        int Q, int q_stride, int qw, int keys, int W, int map_words) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  const int wi = blockIdx.x, qt = blockIdx.y, g = blockIdx.z, lane = threadIdx.x;
  const int nq = min(128, Q - qt * 128);
  // queries of this tile (4 words); a row with flag 0 ignores its mask: nothing can be skipped then
  uint32_t qmask[4];
  bool all_active = true;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int n = nq - 32 * x;
    qmask[x] = n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
  }
  for (int q = lane; q < nq; q += 32) all_active &= flags[(long long)g * q_stride + qt * 128 + q] != 0;
  all_active = __all_sync(0xffffffffu, all_active);
  const int t = wi * 32 + lane;
  const int total_tiles = (keys + X3_KT - 1) / X3_KT;
  bool skip = all_active && t < total_tiles;
  if (skip) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int blk = t * 4 + b;
      if (blk >= W) break;                                  // past the last (partial) 32-key block: nothing there
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(blockand + ((long long)g * W + blk) * qw + qt * 4));
      skip = skip && ((w.x & qmask[0]) == qmask[0]) && ((w.y & qmask[1]) == qmask[1]) && ((w.z & qmask[2]) == qmask[2]) &&
             ((w.w & qmask[3]) == qmask[3]);
    }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, skip);
  if (lane == 0) map[((long long)g * gridDim.y + qt) * map_words + wi] = word;
}

// NCH: 16-column chunks of the query tile known at compile time (7: Q = 100, 8: Q = 128, 5: the second tile of Q = 200);
// 0 = taken from the arguments at run time (any other Q).
template <int NCH>
__global__ void __launch_bounds__(X3_THREADS, 1)
xattn_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const XattnT3Args a) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + X3_OFF_Q;
  uint8_t* sK = smem + X3_OFF_K;
  uint8_t* sV = smem + X3_OFF_V;
  uint8_t* sP = smem + X3_OFF_P;
  uint8_t* sMref = smem + X3_OFF_MREF;
  uint8_t* sE = smem + X3_OFF_E;
  uint8_t* sOnes = smem + X3_OFF_ONES;
  int* smax = reinterpret_cast<int*>(smem + X3_OFF_SMAX);          // [kind: 0 masked, 1 unmasked][head][128]
  float* sref = reinterpret_cast<float*>(smem + X3_OFF_REF);      // [head][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + X3_OFF_BAR);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [3] TMA -> S issuer
  uint64_t* k_empty = bars + 4;            // [3] S issuer (commit) -> TMA
  uint64_t* v_full = bars + 7;             // [3] TMA -> PV issuer
  uint64_t* v_empty = bars + 10;           // [3] PV issuer (commit) -> TMA
  uint64_t* s_full = bars + 13;            // [3 wg] S issuer (commit) -> softmax
  uint64_t* s_empty = bars + 16;           // [3]    softmax -> S issuer (S drained into registers)
  uint64_t* p_full = bars + 19;            // [3]    softmax -> PV issuer (P^T written)
  uint64_t* p_empty = bars + 22;           // [3]    PV issuer (commit) -> softmax (P buffer free)
  uint64_t* done = bars + 25;              // [1]    PV issuer (commit): all products of the pass retired
  uint32_t* ctl = reinterpret_cast<uint32_t*>(bars + 26);          // [0] TMEM base, [1] surviving tile count, [2..5] active words,
                                                                    // [6..9] any-unblocked words
  uint32_t* wpre = reinterpret_cast<uint32_t*>(smem + X3_OFF_LIST);
  uint16_t* tlist = reinterpret_cast<uint16_t*>(wpre + X3_MAP_WORDS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int chunk_id = blockIdx.x >> 2, hp = blockIdx.x & 3, qt = blockIdx.y, g = blockIdx.z;
  const int nq = min(128, a.Q - qt * 128);              // queries of this tile
  const int N = (nq + 15) & ~15;                        // UMMA N of S^T
  const int nchunks = NCH ? NCH : (N >> 4);              // 16-column steps of the softmax loop

  // ---- this CTA's key tiles: a contiguous chunk, or its share of the surviving tiles (tile skipping)
  bool use_list = a.skipmap != nullptr;
  int k_begin = chunk_id * a.chunk;
  int k_end = min(k_begin + a.chunk, a.keys);
  int ntiles = (k_end - k_begin + X3_KT - 1) / X3_KT;
  if (use_list) {
    const uint32_t* map = a.skipmap + ((long long)g * gridDim.y + qt) * a.map_words;
    const int total_tiles = (a.keys + X3_KT - 1) / X3_KT;
    const int words = (total_tiles + 31) >> 5;
    const uint32_t last_mask = (total_tiles & 31) ? ((1u << (total_tiles & 31)) - 1u) : 0xffffffffu;
    for (int i = tid; i < X3_MAP_WORDS; i += X3_THREADS)
      wpre[i] = i < words ? (uint32_t)__popc(~__ldg(map + i) & (i == words - 1 ? last_mask : 0xffffffffu)) : 0u;
    __syncthreads();
    if (warp == 0) {                                     // exclusive scan of 256 counts: 8 per lane
      uint32_t loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = wpre[lane * 8 + j]; sum += loc[j]; }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      uint32_t run = inc - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) { wpre[lane * 8 + j] = run; run += loc[j]; }
      if (lane == 31) ctl[1] = inc;
    }
    __syncthreads();
    const int cnt = (int)ctl[1];
    const int chunks = (int)(gridDim.x >> 2);
    const int per = (cnt + chunks - 1) / chunks;
    const int lo = chunk_id * per, hi = min(cnt, lo + per);
    ntiles = max(0, hi - lo);
    k_end = a.keys;
    if (cnt == total_tiles) {              // nothing to skip: a contiguous run of tiles, no list needed
      use_list = false;
      k_begin = lo * X3_KT;
    } else {
      k_begin = 0;
    }
    for (int i = tid; use_list && i < words; i += X3_THREADS) {
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
  }
  const bool listed = use_list;
  const int k_first = k_begin, k_last = k_end;
  auto tile_key = [=](int t) -> int { return listed ? (int)tlist[t] * X3_KT : k_first + t * X3_KT; };

  // ---- one-time set-up: barriers, TMEM, constant operands, active-row words
  if (tid == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4);
      mbar_init(&p_full[s], 4); mbar_init(&p_empty[s], 1);
    }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == X3_WARP_PV) tmem_alloc(ctl, 512);
  // P^T buffers: zero once (query columns past N are never written but are part of the 128-row A operand)
  for (int i = tid; i < X3_NB * X3_P_BYTES / 16; i += X3_THREADS) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < (2 * 4096 + 8192 + 256) / 4; i += X3_THREADS) reinterpret_cast<uint32_t*>(sMref)[i] = 0u;
  if (tid < 4) {
    // rows that use their mask (flag != 0) as a 128-bit set; rows under the all-masked-row rule attend to every key
    uint32_t w = 0;
    for (int j = 0; j < 32; ++j) {
      const int q = qt * 128 + tid * 32 + j;
      if (q < a.Q && a.flags[(long long)g * a.q_stride + q] != 0) w |= 1u << j;
    }
    ctl[2 + tid] = w;
    ctl[6 + tid] = 0u;
  }
  __syncthreads();
  {
    // ones atom of the PV product's B operand: [128 keys][32 halfs] MN-major SWIZZLE_64B (8-key groups of 512 B, 16-byte
    // chunk c of key row k stored at chunk c ^ ((k >> 1) & 3)), element (k, 0) = 1
    if (tid < 128) *reinterpret_cast<__half*>(sE + (tid >> 3) * 512 + (tid & 7) * 64 + (((tid & 7) >> 1) & 3) * 16) = __float2half_rn(1.f);
    // ones operand: 8 rows x 32 B SWIZZLE_32B, element (r, 0) = 1: 16-byte chunk 0 of row r sits at ((0 ^ (r >> 2)) * 16)
    if (tid < 8) *reinterpret_cast<__half*>(sOnes + tid * 32 + ((tid >> 2) & 1) * 16) = __float2half_rn(1.f);
  }
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl[0];
  const uint32_t act0 = ctl[2], act1 = ctl[3], act2 = ctl[4], act3 = ctl[5];
  // query columns past nq: always blocked (their Q rows belong to the next group or are zero fill)
  uint32_t pad0, pad1, pad2, pad3;
  {
    auto padw = [&](int x) -> uint32_t { const int n = nq - 32 * x; return n >= 32 ? 0u : (n <= 0 ? 0xffffffffu : ~((1u << n) - 1u)); };
    pad0 = padw(0); pad1 = padw(1); pad2 = padw(2); pad3 = padw(3);
  }

  // running counters of every role (the barriers keep counting across passes)
  int kcount = 0;         // K tiles loaded / consumed so far (ring position)
  int vcount = 0;         // V tiles
  int ubase = 0;          // units of the passes so far (S buffers / s_full / s_empty rotate with the global unit number)
  int pbase = 0;          // units of the exponential passes so far (P buffers / p_full / p_empty)
  int passes_exp = 0;     // exponential passes finished (done parity)
  const int wg = warp >> 2, quarter = warp & 3;
  const int r = quarter * 32 + lane;                              // key row within the tile / TMEM lane
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;

  const int scan = min(ntiles, 1);           // tiles the max pass looks at: the first one
  bool retried = false;

  {
    // ================================================================= max pass over the first tile
    for (int i = tid; i < 2 * 2 * 128; i += X3_THREADS) smax[i] = (int)0x80000000;     // ordered_int minimum
    __syncthreads();
    if (warp == X3_WARP_K) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 16384);
        tma_load_2d(sQ, &tmQ, q_full, hp * 64, g * a.Q + qt * 128);
        for (int t = 0; t < scan; ++t) {
          const int st = (kcount + t) % X3_KS;
          mbar_wait(&k_empty[st], (uint32_t)((((kcount + t) / X3_KS) & 1) ^ 1));
          mbar_arrive_expect_tx(&k_full[st], X3_TILE_BYTES);
          tma_load_2d(sK + st * X3_TILE_BYTES, &tmK, &k_full[st], hp * 64, g * a.keys + tile_key(t));
        }
      }
    } else if (warp == X3_WARP_S) {
      if (lane == 0) {
        const uint32_t idesc_s = x3_idesc(128, N, 0, 0);
        mbar_wait(q_full, 0);
        for (int u = 0; u < 2 * scan; ++u) {                 // raw scores (no reference slab)
          const int t = u >> 1, w = u & 1, ug = ubase + u, b = ug % X3_NB;
          const int kc = kcount + t, st = kc % X3_KS;
          mbar_wait(&s_empty[b], (uint32_t)(((ug / X3_NB) & 1) ^ 1));
          mbar_wait(&k_full[st], (uint32_t)((kc / X3_KS) & 1));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * 128);
          const uint64_t adesc = x3_desc(smem_u32(sK) + st * X3_TILE_BYTES + w * 64, 16, 1024, 2);
          const uint64_t bdesc = x3_desc(smem_u32(sQ) + w * 64, 16, 1024, 2);
          umma_f16(d_tmem, adesc, bdesc, idesc_s, 0u);
          umma_f16(d_tmem, adesc + 2, bdesc + 2, idesc_s, 1u);
          umma_commit(&s_full[b]);
          if (w == 1) umma_commit(&k_empty[st]);
        }
      }
    } else if (warp < 4 * X3_NWG) {
      // ---- column maxima (masked and unmasked) of this warpgroup's units
      float mm[2][4], mu[2][4];
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int x = 0; x < 4; ++x) { mm[w][x] = -INFINITY; mu[w][x] = -INFINITY; }
      for (int u = wg; u < 2 * scan; u += X3_NWG) {
        const int t = u >> 1, w = u & 1, ug = ubase + u, b = ug % X3_NB;
        const int key = tile_key(t) + r;
        const bool kvalid = key < k_last;
        uint4 bw = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (kvalid) bw = __ldg(reinterpret_cast<const uint4*>(a.bits_t + ((long long)g * a.keys + key) * a.qw + qt * 4));
        const uint32_t m0 = kvalid ? ((bw.x & act0) | pad0) : 0xffffffffu, m1 = kvalid ? ((bw.y & act1) | pad1) : 0xffffffffu;
        const uint32_t m2 = kvalid ? ((bw.z & act2) | pad2) : 0xffffffffu, m3 = kvalid ? ((bw.w & act3) | pad3) : 0xffffffffu;
        mbar_wait(&s_full[b], (uint32_t)((ug / X3_NB) & 1));
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_off + (uint32_t)(b * 128);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
          uint32_t sv[16];
          tmem_ld_32x16_nowait(s_addr + c * 16, sv);
          tmem_ld_wait();
          reg_fence16(sv);
          const uint32_t word = (c >> 1) == 0 ? m0 : (c >> 1) == 1 ? m1 : (c >> 1) == 2 ? m2 : m3;
          const uint32_t hw = (c & 1) ? (word >> 16) : word;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float sc = __uint_as_float(sv[j]);
            const float scu = kvalid ? sc : -INFINITY;
            const float scm = (hw & (1u << j)) ? -INFINITY : sc;
            const float cu = warp_max_f32(scu), cm = warp_max_f32(scm);
            const int col = c * 16 + j;
            if (lane == (col & 31)) {
#pragma unroll
              for (int x = 0; x < 4; ++x)
                if ((col >> 5) == x) {
                  if (w == 0) { mm[0][x] = fmaxf(mm[0][x], cm); mu[0][x] = fmaxf(mu[0][x], cu); }
                  else { mm[1][x] = fmaxf(mm[1][x], cm); mu[1][x] = fmaxf(mu[1][x], cu); }
                }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[b]);
      }
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          atomicMax(&smax[(0 * 2 + w) * 128 + x * 32 + lane], ordered_int(mm[w][x]));
          atomicMax(&smax[(1 * 2 + w) * 128 + x * 32 + lane], ordered_int(mu[w][x]));
        }
    }
    kcount += scan;
    ubase += 2 * scan;
    __syncthreads();
    // ---- references: masked maximum of the first tile + 2 (else its unmasked maximum - 2, else 0), fp16-representable
    if (tid < 256) {
      const int w = tid >> 7, q = tid & 127;
      float m = ordered_float(smax[(0 * 2 + w) * 128 + q]) + 2.f;
      if (!(m > -INFINITY)) m = ordered_float(smax[(1 * 2 + w) * 128 + q]) - 2.f;
      if (!(m > -INFINITY)) m = 0.f;
      m = fminf(fmaxf(m, -60000.f), 60000.f);
      const __half hneg = __float2half_rd(-m);
      sref[w * 128 + q] = -__half2float(hneg);
      // Mref_w: [128 q][16 halfs] K-major SWIZZLE_32B, element (q, 0) = -m_ref
      *reinterpret_cast<__half*>(sMref + w * 4096 + (q >> 3) * 256 + (q & 7) * 32 + (((q & 7) >> 2) & 1) * 16) = hneg;
    }
    fence_async_proxy();
    __syncthreads();
  }

#pragma unroll 1
  for (int attempt = 0;; ++attempt) {
    // ================================================================= exponential pass over all tiles
    uint32_t anyun0 = 0, anyun1 = 0, anyun2 = 0, anyun3 = 0;     // queries that saw an unblocked key in this chunk
    const int units = 2 * ntiles;
    if (warp == X3_WARP_K) {
      if (lane == 0) {
        // (the ring holds ~1 tile in flight: not enough bytes for HBM latency, so tiles are pulled into L2 early)
        for (int t = 0; t < min(ntiles, X3_PF); ++t) tma_prefetch_2d(&tmK, hp * 64, g * a.keys + tile_key(t));
        for (int t = 0; t < ntiles; ++t) {
          if (t + X3_PF < ntiles) tma_prefetch_2d(&tmK, hp * 64, g * a.keys + tile_key(t + X3_PF));
          const int st = (kcount + t) % X3_KS;
          X3_TRACE(5, t, 0);
          mbar_wait(&k_empty[st], (uint32_t)((((kcount + t) / X3_KS) & 1) ^ 1));
          X3_TRACE(5, t, 1);
          mbar_arrive_expect_tx(&k_full[st], X3_TILE_BYTES);
          tma_load_2d(sK + st * X3_TILE_BYTES, &tmK, &k_full[st], hp * 64, g * a.keys + tile_key(t));
        }
      }
    } else if (warp == X3_WARP_V) {
      if (lane == 0) {
        for (int t = 0; t < min(ntiles, X3_PF); ++t) {
          tma_prefetch_2d(&tmV, hp * 64, g * a.keys + tile_key(t));
          tma_prefetch_2d(&tmV, hp * 64 + 32, g * a.keys + tile_key(t));
        }
        for (int t = 0; t < ntiles; ++t) {
          if (t + X3_PF < ntiles) {
            tma_prefetch_2d(&tmV, hp * 64, g * a.keys + tile_key(t + X3_PF));
            tma_prefetch_2d(&tmV, hp * 64 + 32, g * a.keys + tile_key(t + X3_PF));
          }
          const int st = (vcount + t) % X3_VS;
          X3_TRACE(6, t, 0);
          mbar_wait(&v_empty[st], (uint32_t)((((vcount + t) / X3_VS) & 1) ^ 1));
          X3_TRACE(6, t, 1);
          mbar_arrive_expect_tx(&v_full[st], X3_TILE_BYTES);
          tma_load_2d(sV + st * X3_TILE_BYTES, &tmV, &v_full[st], hp * 64, g * a.keys + tile_key(t));               // head 0: [128][32 ch]
          tma_load_2d(sV + st * X3_TILE_BYTES + 8192, &tmV, &v_full[st], hp * 64 + 32, g * a.keys + tile_key(t));   // head 1
        }
      }
    } else if (warp == X3_WARP_S) {
      if (lane == 0) {
        // S^T of unit u goes to buffer (global unit number) % 3: with two softmax warpgroups the third buffer always holds
        // the NEXT unit of whichever warpgroup finishes first, so a warpgroup never waits for its scores
        const uint32_t idesc_s = x3_idesc(128, N, 0, 0);
        const uint64_t ones_desc = x3_desc(smem_u32(sOnes), 16, 0, 6);          // SBO = 0: every 8-row atom is the same atom
        const uint64_t kdesc0 = x3_desc(smem_u32(sK), 16, 1024, 2);
        const uint64_t qdesc0 = x3_desc(smem_u32(sQ), 16, 1024, 2);
        const uint64_t mdesc0 = x3_desc(smem_u32(sMref), 16, 256, 6);
        for (int u = 0; u < units; ++u) {
          const int t = u >> 1, w = u & 1, ug = ubase + u, b = ug % X3_NB;
          const int kc = kcount + t, st = kc % X3_KS;
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * 128);
          const uint64_t adesc = kdesc0 + (uint64_t)((st * X3_TILE_BYTES + w * 64) >> 4);
          const uint64_t bdesc = qdesc0 + (uint64_t)((w * 64) >> 4);
          const uint64_t mdesc = mdesc0 + (uint64_t)((w * 4096) >> 4);
          mbar_wait(&s_empty[b], (uint32_t)(((ug / X3_NB) & 1) ^ 1));
          mbar_wait(&k_full[st], (uint32_t)((kc / X3_KS) & 1));
          tc_fence_after();
#ifndef X3_EXPERIMENT_NO_S
          umma_f16(d_tmem, adesc, bdesc, idesc_s, 0u);
          umma_f16_acc(d_tmem, adesc + 2, bdesc + 2, idesc_s);
          umma_f16_acc(d_tmem, ones_desc, mdesc, idesc_s);                       // - m_ref
#endif
          umma_commit(&s_full[b]);
          X3_TRACE(3, u, 0);
          if (w == 1) umma_commit(&k_empty[st]);
        }
      }
    } else if (warp == X3_WARP_PV) {
      if (lane == 0) {
        constexpr uint32_t idesc_o = x3_idesc(128, 48, 1, 1);      // A = P^T MN-major, B = [V_h | ones] MN-major
        // (descriptors are built once per unit and stepped by adding to the 14-bit address field: the issuing thread's own
        //  instruction stream was the limit -- ~20 integer instructions per UMMA cost ~450 cycles per unit)
        const uint64_t pdesc0 = x3_desc(smem_u32(sP), 16384, 1024, 2);
        for (int u = 0; u < units; ++u) {
          const int t = u >> 1, w = u & 1, pg = pbase + u, b = pg % X3_NB;
          const int vc = vcount + t, st = vc % X3_VS;
          const uint32_t v_addr = smem_u32(sV) + st * X3_TILE_BYTES + w * 8192;
          const uint64_t adesc = pdesc0 + (uint64_t)((b * X3_P_BYTES) >> 4);
          const uint64_t bdesc = x3_desc(v_addr, smem_u32(sE) - v_addr, 512, 4);   // second 32-wide atom of B: the ones atom
          const uint32_t o_tmem = tmem_base + 384u + (uint32_t)(w * 48);
          mbar_wait(&p_full[b], (uint32_t)((pg / X3_NB) & 1));
          mbar_wait(&v_full[st], (uint32_t)((vc / X3_VS) & 1));
          tc_fence_after();
#ifndef X3_EXPERIMENT_NO_PV
#ifndef X3_EXP_PV_STEPS
#define X3_EXP_PV_STEPS (X3_KT / 16)
#endif
#ifdef X3_EXP_PV_N32
          constexpr uint32_t idesc_x = x3_idesc(128, 32, 1, 1);
#else
          constexpr uint32_t idesc_x = idesc_o;
#endif
          umma_f16(o_tmem, adesc, bdesc, idesc_x, t > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 1; kk < X3_EXP_PV_STEPS; ++kk)
            umma_f16_acc(o_tmem, adesc + (uint64_t)(kk * (2048 >> 4)), bdesc + (uint64_t)(kk * (1024 >> 4)), idesc_x);
#endif
          umma_commit(&p_empty[b]);
          X3_TRACE(4, u, 0);
          if (w == 1) umma_commit(&v_empty[st]);
        }
        umma_commit(done);
      }
    } else {
      // ---- softmax warpgroups: thread = key row of the unit's tile; warpgroup g takes the units u = g (mod 2).
      // The chunks of consecutive units form ONE stream: the TMEM load of the next chunk -- also across a unit boundary --
      // is always in flight while the current chunk is computed, so the barrier round trips, the proxy fence and the
      // first-load latency of a unit hide behind arithmetic (they were ~650 of ~2900 cycles per unit).
      uint4 nb = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      bool nvalid = false;
      if (wg < units) {
        const int key = tile_key(wg >> 1) + r;
        nvalid = key < k_last;
        if (nvalid) nb = __ldg(reinterpret_cast<const uint4*>(a.bits_t + ((long long)g * a.keys + key) * a.qw + qt * 4));
      }
      uint32_t sv[2][16];
      const bool tr = (quarter == 0 && lane == 0);
      if (wg < units) {                                         // first chunk of the first unit
        const int ug = ubase + wg;
        mbar_wait(&s_full[ug % X3_NB], (uint32_t)((ug / X3_NB) & 1));
        tc_fence_after();
        tmem_ld_32x16_nowait(tmem_base + lane_off + (uint32_t)((ug % X3_NB) * 128), sv[0]);
      }
      // PH: which of the two register buffers holds chunk 0 of the unit (alternates when the chunk count is odd)
      auto unit_body = [&](auto phc, const int u) {
        constexpr int PH = decltype(phc)::value;
        const int ug = ubase + u, sb = ug % X3_NB, pg = pbase + u, pb = pg % X3_NB;
        const uint32_t prow = smem_u32(sP) + pb * X3_P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
        const uint32_t s_addr = tmem_base + lane_off + (uint32_t)(sb * 128);
        const uint4 bw = nb;
        const bool kvalid = nvalid;
        if (u + X3_NWG < units) {                               // mask words of the next unit
          const int key = tile_key((u + X3_NWG) >> 1) + r;
          nvalid = key < k_last;
          nb = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          if (nvalid) nb = __ldg(reinterpret_cast<const uint4*>(a.bits_t + ((long long)g * a.keys + key) * a.qw + qt * 4));
        }
        const uint32_t m0 = kvalid ? ((bw.x & act0) | pad0) : 0xffffffffu, m1 = kvalid ? ((bw.y & act1) | pad1) : 0xffffffffu;
        const uint32_t m2 = kvalid ? ((bw.z & act2) | pad2) : 0xffffffffu, m3 = kvalid ? ((bw.w & act3) | pad3) : 0xffffffffu;
        anyun0 |= ~m0; anyun1 |= ~m1; anyun2 |= ~m2; anyun3 |= ~m3;
        if (tr) X3_TRACE(wg, u / X3_NWG, 0);                  // unit begins
        if (X3_NB <= X3_NWG && u >= X3_NWG) {                   // (no spare S^T buffer: the first chunk is loaded here)
          mbar_wait(&s_full[sb], (uint32_t)((ug / X3_NB) & 1));
          tc_fence_after();
          tmem_ld_32x16_nowait(s_addr, sv[PH & 1]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nchunks) {
            uint32_t* cur = sv[(c + PH) & 1];
            uint32_t* nxt = sv[(c + PH + 1) & 1];
            tmem_ld_wait();
            reg_fence16(cur);
            if (c + 1 < nchunks) {
              tmem_ld_32x16_nowait(s_addr + (c + 1) * 16, nxt);
            } else {
              // the S buffer goes back to the issuer as soon as its last columns sit in registers ...
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s_empty[sb]);
              if (tr) X3_TRACE(wg, u / X3_NWG, 4);            // S drained
              // ... and the first chunk of this warpgroup's next unit starts loading (its S^T sits in the third buffer)
              if (X3_NB > X3_NWG && u + X3_NWG < units) {
                const int ug2 = ug + X3_NWG;
                mbar_wait(&s_full[ug2 % X3_NB], (uint32_t)((ug2 / X3_NB) & 1));
                tc_fence_after();
                tmem_ld_32x16_nowait(tmem_base + lane_off + (uint32_t)((ug2 % X3_NB) * 128), nxt);
              }
            }
            const uint32_t word = (c >> 1) == 0 ? m0 : (c >> 1) == 1 ? m1 : (c >> 1) == 2 ? m2 : m3;
            const uint32_t hw = (c & 1) ? (word >> 16) : word;
            uint32_t pk[8];
#ifdef X3_EXPERIMENT_SKELETON
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = cur[2 * j] ^ hw;
#else
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float s0 = __uint_as_float(cur[j]), s1 = __uint_as_float(cur[j + 1]);
              if (hw & (1u << j)) s0 = -INFINITY;
              if (hw & (1u << (j + 1))) s1 = -INFINITY;
              const bool poly1 = X3_POLY_MOD > 0 && ((j + 1) % (X3_POLY_MOD > 0 ? X3_POLY_MOD : 1)) == (X3_POLY_MOD > 0 ? X3_POLY_MOD : 1) - 1;
              const bool poly0 = X3_POLY_MOD > 0 && (j % (X3_POLY_MOD > 0 ? X3_POLY_MOD : 1)) == (X3_POLY_MOD > 0 ? X3_POLY_MOD : 1) - 1;
              pk[j >> 1] = pack_half2_sat(poly0 ? poly_ex2(s0) : fast_ex2(s0), poly1 ? poly_ex2(s1) : fast_ex2(s1));
            }
#endif
            if (c == 0) {
              // P^T buffer pg % 3 was last read by the products of unit pg - 3: retired long ago in steady state
              if (tr) X3_TRACE(wg, u / X3_NWG, 2);
              mbar_wait(&p_empty[pb], (uint32_t)(((pg / X3_NB) & 1) ^ 1));
              if (tr) X3_TRACE(wg, u / X3_NWG, 3);
            }
            // P^T[key r][queries 16c .. 16c+15]: two 16-byte chunks of the key's 128-byte row in query atom c / 4
            const uint32_t dst = prow + (c >> 2) * 16384;
            const int cc = (c & 3) * 2;
#ifdef X3_EXPERIMENT_NO_STORE
            if (pk[0] == 0x12345678u && pk[5] == 0x9abcdef0u)
#endif
            {
            st_shared_v4(dst + (((cc) ^ (r & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
            st_shared_v4(dst + (((cc + 1) ^ (r & 7)) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
            }
          }
        }
        fence_async_proxy();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
        if (tr) X3_TRACE(wg, u / X3_NWG, 5);                  // P handed over
      };
      if (NCH != 0 && (NCH & 1) == 0) {                         // even chunk count: chunk 0 always lands in buffer 0
#pragma unroll 1
        for (int u = wg; u < units; u += X3_NWG) unit_body(std::integral_constant<int, 0>{}, u);
      } else if (NCH != 0) {                                    // odd: the buffers swap roles from unit to unit
#pragma unroll 1
        for (int u = wg; u < units; u += 2 * X3_NWG) {
          unit_body(std::integral_constant<int, 0>{}, u);
          if (u + X3_NWG < units) unit_body(std::integral_constant<int, 1>{}, u + X3_NWG);
        }
      } else {                                                  // run-time chunk count: parity decided per unit
        int ph = 0;
#pragma unroll 1
        for (int u = wg; u < units; u += X3_NWG) {
          if (ph) unit_body(std::integral_constant<int, 1>{}, u); else unit_body(std::integral_constant<int, 0>{}, u);
          ph ^= nchunks & 1;
        }
      }
    }
    kcount += ntiles;
    vcount += ntiles;
    ubase += units;
    pbase += units;

    // ================================================================= partials + reference check
    int bad = 0;
    if (warp < 4 * X3_NWG) {
      // which queries saw an unblocked key anywhere in this chunk (needed to tell "all blocked" from "all underflowed")
      anyun0 = __reduce_or_sync(0xffffffffu, anyun0); anyun1 = __reduce_or_sync(0xffffffffu, anyun1);
      anyun2 = __reduce_or_sync(0xffffffffu, anyun2); anyun3 = __reduce_or_sync(0xffffffffu, anyun3);
      if (lane == 0) { atomicOr(&ctl[6], anyun0); atomicOr(&ctl[7], anyun1); atomicOr(&ctl[8], anyun2); atomicOr(&ctl[9], anyun3); }
    }
    __syncthreads();
    if (warp < 4) {
      const int q = qt * 128 + r;                           // r = query row here (TMEM lane of O / L)
      if (ntiles > 0) {
        mbar_wait(done, (uint32_t)(passes_exp & 1));
        tc_fence_after();
      }
      const bool had_unblocked = (ctl[6 + (r >> 5)] >> (r & 31)) & 1u;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        uint32_t ov[32], lv[16];
        if (ntiles > 0) {
          tmem_ld_32x32_nowait(tmem_base + lane_off + 384u + (uint32_t)(w * 48), ov);
          tmem_ld_32x16_nowait(tmem_base + lane_off + 384u + (uint32_t)(w * 48 + 32), lv);
          tmem_ld_wait();
          reg_fence32(ov);
          reg_fence16(lv);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) ov[c] = 0u;
          lv[0] = 0u;
        }
        const float l = __uint_as_float(lv[0]);
        const float mref_used = sref[w * 128 + r];             // the reference this attempt's O / l belong to
#ifdef X3_EXPERIMENT_SKELETON
        if (false) {
#else
        if (r < nq && had_unblocked && !(l >= 0.125f && l < 32768.f)) {
#endif
          // a probability may have saturated (l >= 2^15), or the row sat in / below fp16's subnormal range (l < 2^-3):
          // new reference from the sum itself, so that the row's sum lands near 2^3 next time
          bad = 1;
          float m = sref[w * 128 + r] + (l > 0.f ? log2f(l) - 3.f : -22.f);
          m = fminf(fmaxf(m, -60000.f), 60000.f);
          const __half hneg = __float2half_rd(-m);
          sref[w * 128 + r] = -__half2float(hneg);
          *reinterpret_cast<__half*>(sMref + w * 4096 + (r >> 3) * 256 + (r & 7) * 32 + (((r & 7) >> 2) & 1) * 16) = hneg;
        }
        const int h = 2 * hp + w;
        const long long pr = ((((long long)g * a.splits + chunk_id) * 8) + h) * a.q_pad + q;
        if (q < a.q_pad) {
          float* op = a.o_part + pr * 32;
#pragma unroll
          for (int c = 0; c < 32; c += 4)
            *reinterpret_cast<float4*>(op + c) = make_float4(__uint_as_float(ov[c]), __uint_as_float(ov[c + 1]),
                                                             __uint_as_float(ov[c + 2]), __uint_as_float(ov[c + 3]));
          const float mref = (l > 0.f) ? mref_used : -INFINITY;                // an empty partial is skipped by the combine
          *reinterpret_cast<float2*>(a.ml_part + pr * 2) = make_float2(mref, l);
        }
      }
      tc_fence_before();
      fence_async_proxy();                     // (new references -> visible to the S^T products of the next attempt)
    }
    if (ntiles > 0) ++passes_exp;
    const int any_bad = __syncthreads_or(bad);
    if (!any_bad || attempt >= 6) break;
    retried = true;
    if (tid < 4) ctl[6 + tid] = 0u;
  }

  if (a.stats && tid == 0) {
    atomicAdd(&a.stats[0], 1);
    if (retried) atomicAdd(&a.stats[1], 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == X3_WARP_PV) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ovis
