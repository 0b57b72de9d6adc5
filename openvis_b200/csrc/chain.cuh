// Query-side chain: ONE persistent CTA per group (a clip's or a frame's <= 128 queries) runs the whole post-attention part
// of a decoder layer as a list of phases -- out-projection + residual + LayerNorm, self-attention in-projections,
// self-attention, its out-projection + LayerNorm, FFN (2048 wide) + LayerNorm + decoder_norm, the mask-embed MLP and the
// next layer's query projection (reference: SelfAttentionLayer / CrossAttentionLayer / FFNLayer.forward_post and MLP,
// video_mask2former_transformer_decoder.py:52-62, 110-122, 175-179, 204-216).  Every phase is a 128-row x N GEMM of the
// tcgen05 pipeline of gemm_tn_kernel<256> (same TMA ring, same TMEM double buffer, same fused epilogues) restricted to the
// CTA's own rows, or the group's 8-head self-attention (one warp per head, mma.sync); a phase's output is read by the next
// phase of the SAME CTA only, so a CTA-wide barrier + proxy fence replaces the kernel boundary.  This turns ~14 dependent
// launches per layer (5-15 us each, most of them one busy tile per SM) into one.
#pragma once
#include "gemm_tn.cuh"
#include "prep.cuh"
#include "xattn.cuh"

namespace ovis {

struct alignas(128) ChainPhase {
  CUtensorMap tmA;        // [G * Q rows][K] fp16, box 128 rows x 64
  CUtensorMap tmB;        // [N][K] fp16 weights, box 256 rows x 64
  GemmArgs args;          // kind 0: rows_per_group = Q, a_group_stride = Q, num_groups = G
  SelfAttnArgs sa;        // kind 1
  LnReduceArgs lnr;       // kind 3 (wide chain only): row-parallel bias + residual + LayerNorm(s) over the GEMM's fp32 partials
  int kind;               // 0 GEMM phase, 1 self-attention, 3 split-K GEMM -> grid barrier -> LayerNorm reduction
  int par;                // wide chain: independent of the NEXT phase -> no barrier after this one (both run side by side)
  int cta_off;            // wide chain: CTA that takes this phase's item 0 (set by the launch: parallel phases get disjoint CTAs)
};

// The phases of one launch travel as a kernel parameter (constant bank): TMA descriptors read from plain global memory made
// every bulk load wait for a descriptor fetch (1560 cycles per 48 KB k-block measured, 3x the MMA time).
constexpr int CHAIN_MAX_PHASES = 12;
struct ChainLaunch {
  ChainPhase ph[CHAIN_MAX_PHASES];
};
static_assert(sizeof(ChainLaunch) <= 32000, "kernel parameter space");

constexpr int CHAIN_SMEM = GemmCfg<256>::SMEM_BYTES;

// event trace of group 0's CTA (tools/prof_chain.py --events; -DOVIS_CHAIN_EVTRACE builds only): trace[64] = counter,
// then (tag, cycle) pairs
#ifdef OVIS_CHAIN_EVTRACE
#define CHAIN_EV(tag)                                                                  \
  do {                                                                                 \
    if (trace && blockIdx.x == 0) trace[66 + (tag)] = clock64();                        \
  } while (0)
#else
#define CHAIN_EV(tag) do {} while (0)
#endif

// one warp = one head of one group: K / V of the head in the warp's own shared-memory slab, flash loop over 64-key blocks
// (fill: `nthr` threads, thread `tid`, copy K / V of (g, head) into the slab; tiles: the caller's warp takes the 16-row query
//  tiles mt0, mt0 + mtstep, ...)
__device__ __forceinline__ void chain_self_attn_fill(const SelfAttnArgs& a, int g, int head, __half* slab, int tid, int nthr) {
  const int Q = a.Q;
  const int Qp = (Q + SA_KB - 1) / SA_KB * SA_KB;
  __half* sk = slab;
  __half* sv = slab + (size_t)Qp * XA_LD;
  for (int i = tid; i < Qp * 4; i += nthr) {
    const int row = i >> 2, ch = i & 3;
    uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
    if (row < Q) {
      const long long r = (long long)g * Q + row;
      kv = __ldcg(reinterpret_cast<const uint4*>(a.qk + r * 512 + 256 + head * 32 + ch * 8));   // written earlier in this kernel
      vv = __ldcg(reinterpret_cast<const uint4*>(a.v + r * 256 + head * 32 + ch * 8));
    }
    *reinterpret_cast<uint4*>(sk + row * XA_LD + ch * 8) = kv;
    *reinterpret_cast<uint4*>(sv + row * XA_LD + ch * 8) = vv;
  }
}

__device__ __forceinline__ void chain_self_attn_tiles(const SelfAttnArgs& a, int g, int head, const __half* slab, int lane,
                                                      int mt0, int mtstep) {
  const int Q = a.Q;
  const int Qp = (Q + SA_KB - 1) / SA_KB * SA_KB;
  const __half* sk = slab;
  const __half* sv = slab + (size_t)Qp * XA_LD;
  const int quad = lane >> 2, tq = lane & 3;
  const int mtiles = (Q + 15) >> 4;
  for (int mt = mt0; mt < mtiles; mt += mtstep) {
    uint32_t qf[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = mt * 16 + quad + (i & 1) * 8;
        const int col = head * 32 + ks * 16 + tq * 2 + (i >> 1) * 8;
        qf[ks][i] = row < Q ? __ldcg(reinterpret_cast<const uint32_t*>(a.qk + ((long long)g * Q + row) * 512 + col)) : 0u;
      }
    float o[4][4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[dn][i] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    for (int kb = 0; kb < Q; kb += SA_KB) {
      float sc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sc[nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const __half* kr = sk + (kb + nt * 8 + quad) * XA_LD + ks * 16 + tq * 2;
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + 8);
          mma_16816(sc[nt], qf[ks], b0, b1);
        }
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int key = kb + nt * 8 + tq * 2 + (i & 1);
          const float v = key < Q ? sc[nt][i] * a.scale_log2 : -INFINITY;
          sc[nt][i] = v;
          mx[i >> 1] = fmaxf(mx[i >> 1], v);
        }
      float corr[2];
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 1));
        mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 2));
        const float mnew = fmaxf(mrow[hi], mx[hi]);
        corr[hi] = exp2f(mrow[hi] - mnew);
        mrow[hi] = mnew;
      }
      float ls[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = exp2f(sc[nt][i] - mrow[i >> 1]);
          sc[nt][i] = p;
          ls[i >> 1] += p;
        }
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) lrow[hi] = lrow[hi] * corr[hi] + ls[hi];
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        o[dn][0] *= corr[0]; o[dn][1] *= corr[0];
        o[dn][2] *= corr[1]; o[dn][3] *= corr[1];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t pa[4];
        pa[0] = pack_half2(sc[2 * j][0], sc[2 * j][1]);
        pa[1] = pack_half2(sc[2 * j][2], sc[2 * j][3]);
        pa[2] = pack_half2(sc[2 * j + 1][0], sc[2 * j + 1][1]);
        pa[3] = pack_half2(sc[2 * j + 1][2], sc[2 * j + 1][3]);
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(b0, b1, sv + (kb + j * 16 + (lane & 15)) * XA_LD + dn * 8);
          mma_16816(o[dn], pa, b0, b1);
        }
      }
    }
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      lrow[hi] += __shfl_xor_sync(0xffffffffu, lrow[hi], 1);
      lrow[hi] += __shfl_xor_sync(0xffffffffu, lrow[hi], 2);
    }
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int row = mt * 16 + quad + hi * 8;
      if (row < Q) {
        const float inv = 1.f / lrow[hi];
        __half* op = a.out + ((long long)g * Q + row) * 256 + head * 32 + tq * 2;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
          *reinterpret_cast<uint32_t*>(op + dn * 8) = pack_half2(o[dn][hi * 2] * inv, o[dn][hi * 2 + 1] * inv);
      }
    }
  }
}

__device__ __forceinline__ void chain_self_attn_head(const SelfAttnArgs& a, int g, int head, __half* slab, int lane) {
  chain_self_attn_fill(a, g, head, slab, lane, 32);
  __syncwarp();
  chain_self_attn_tiles(a, g, head, slab, lane, 0, 1);
}

__global__ void __launch_bounds__(320, 1)
gemm_chain_kernel(const __grid_constant__ ChainLaunch cl, int nphases, int G, long long* __restrict__ trace) {
  using Cfg = GemmCfg<256>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_smem + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_begin();
  const uint32_t tmem_base = *tmem_holder;

  // pipeline state of the three roles: persists across phases and groups
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  long long pf_frame = -1;
  uint32_t pf_set = 0;

  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    for (int p = 0; p < nphases; ++p) {
      const ChainPhase* ph = &cl.ph[p];
      const GemmArgs& args = ph->args;
      const int kind = ph->kind;
      if (trace && threadIdx.x == 0 && g == 0) trace[p] = clock64();   // (profiling builds of tools/prof_chain.py)
      if (kind == 0) {
        const int n_tiles = (args.N + BN - 1) / BN;
        const int k_blocks = args.K / Cfg::BK;
        if (warp == 0) {
          if (lane == 0) {
            tma_prefetch_desc(&ph->tmA);
            tma_prefetch_desc(&ph->tmB);
            const int a_row = g * args.a_group_stride;
            for (int nt = 0; nt < n_tiles; ++nt) {
              for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (kb < 8) CHAIN_EV(0 * 2000 + p * 160 + nt * 16 + kb);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sb = sa + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_2d(sa, &ph->tmA, &full_bar[stage], kb * Cfg::BK, a_row);
                tma_load_2d(sb, &ph->tmB, &full_bar[stage], kb * Cfg::BK, nt * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
            }
          }
        } else if (warp == 1) {
          if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN);
            for (int nt = 0; nt < n_tiles; ++nt) {
              mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + acc * BN;
              for (int kb = 0; kb < k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (kb < 8) CHAIN_EV(1 * 2000 + p * 160 + nt * 16 + kb);
                const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint64_t adesc = umma_desc_k_sw128(sa);
                const uint64_t bdesc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                for (int k = 0; k < Cfg::BK / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
              umma_commit(&tfull_bar[acc]);
              if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
          }
        } else {
          for (int nt = 0; nt < n_tiles; ++nt) {
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            #ifdef OVIS_CHAIN_EVTRACE
            if (lane == 0 && warp == 2) { CHAIN_EV(2 * 2000 + p * 160 + nt * 16); g_ev_base = trace ? trace + 66 + 2 * 2000 + p * 160 + nt * 16 : nullptr; }
#endif
            gemm_epilogue_tile<BN>(args, nullptr, nt, 0, g, warp, lane, tmem_base + acc * BN, stage_smem, pf_frame, pf_set);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (lane == 0 && warp == 2) CHAIN_EV(2 * 2000 + p * 160 + nt * 16 + 15);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
        // (every thread keeps its own copy of the ring / accumulator counters and advances only those its role uses)
      } else {
        if (warp >= 2) chain_self_attn_head(ph->sa, g, warp - 2, reinterpret_cast<__half*>(smem) + (size_t)(warp - 2) * (2 * 128 * XA_LD), lane);
      }
      // this phase's global writes -> visible to the next phase's TMA loads (async proxy) and loads of the other warps
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncthreads();
    }
    if (trace && threadIdx.x == 0 && g == 0) trace[nphases] = clock64();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// "Wide" chain: the same phase list, but every phase's tiles are spread over ALL CTAs of the launch and a grid-wide barrier
// replaces the CTA-wide one.  For calls with many groups (Frame decoders: one group per frame, thousands of query rows) the
// one-CTA-per-group chain serialises a layer's tiles on a few SMs, and the launch-per-op schedule pays ~14 dependent
// launches of 5-15 us per layer; here a layer is ONE cooperative launch whose phases cost their own latency chain plus a
// ~1.5 us barrier.  Rules that keep it correct without L1 invalidations:
//   * operands written by other CTAs are only ever read through TMA (L2) or ld.global.cg (self-attention);
//   * the residual stream (fp32, read with plain loads by the LayerNorm epilogue) is produced AND consumed by LayerNorm
//     phases only, and a LayerNorm phase has one column tile, so tile mt always belongs to CTA mt % gridDim.x.
__device__ __forceinline__ void chain_grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __threadfence();                                   // this thread's global writes of the phase
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int g0, g1;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g0) : "l"(bar + 1) : "memory");     // generation before arriving
    unsigned int prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
    if (prev == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");   // self-resetting: replayable from a graph
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar + 1) : "memory");
    } else {
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g1) : "l"(bar + 1) : "memory");
      } while (g1 == g0);
    }
  }
  __syncthreads();
  asm volatile("fence.proxy.async;" ::: "memory");   // the next phase's bulk loads see the other CTAs' writes
}

__global__ void __launch_bounds__(320, 1)
gemm_chain_wide_kernel(const __grid_constant__ ChainLaunch cl, int nphases, int G, unsigned int* __restrict__ bar) {
  using Cfg = GemmCfg<256>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_smem + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  long long pf_frame = -1;
  uint32_t pf_set = 0;

  for (int p = 0; p < nphases; ++p) {
    const ChainPhase* ph = &cl.ph[p];
    const GemmArgs& args = ph->args;
    if (ph->kind == 0 || ph->kind == 3) {
      // items = (group, row tile, column tile); a "group" is a K slice of the split linear + LayerNorm phases
      const int n_tiles = (args.N + BN - 1) / BN;
      const int m_tiles = (args.rows_per_group + Cfg::BM - 1) / Cfg::BM;
      const int per_group = m_tiles * n_tiles;
      const int items = per_group * args.num_groups;
      const int k_blocks = args.K / Cfg::BK;
      const int cta = (int)((blockIdx.x + gridDim.x - (unsigned)ph->cta_off) % gridDim.x);
      if (warp == 0) {
        if (lane == 0) {
          tma_prefetch_desc(&ph->tmA);
          tma_prefetch_desc(&ph->tmB);
          for (int it = cta; it < items; it += gridDim.x) {
            const int g = it / per_group, r = it - g * per_group;
            const int mt = r % m_tiles, nt = r / m_tiles;
            const int a_col = (g % args.a_k_mod) * args.a_k_offset_stride;
            const int b_col = (g % args.b_k_mod) * args.b_k_offset_stride;
            const int a_row = (g / args.a_row_div) * args.a_group_stride + mt * Cfg::BM;
            const int b_row = (g / args.b_row_div) * args.b_group_stride + nt * BN;
            for (int kb = 0; kb < k_blocks; ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              uint8_t* sb = sa + Cfg::A_BYTES;
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_2d(sa, &ph->tmA, &full_bar[stage], a_col + kb * Cfg::BK, a_row);
              tma_load_2d(sb, &ph->tmB, &full_bar[stage], b_col + kb * Cfg::BK, b_row);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN);
          for (int it = cta; it < items; it += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < k_blocks; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
              const uint64_t adesc = umma_desc_k_sw128(sa);
              const uint64_t bdesc = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
              for (int k = 0; k < Cfg::BK / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
      } else {
        for (int it = cta; it < items; it += gridDim.x) {
          const int g = it / per_group, r = it - g * per_group;
          const int mt = r % m_tiles, nt = r / m_tiles;
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          gemm_epilogue_tile<BN>(args, nullptr, nt, mt, g, warp, lane, tmem_base + acc * BN, stage_smem, pf_frame, pf_set);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (ph->kind == 3) {
        // the K slices' fp32 partials -> bias + residual + LayerNorm(s), one warp per row, every warp of the launch.  A row is
        // always finished by the same warp of the same CTA (its residual is that warp's own earlier write).
        chain_grid_barrier(bar, gridDim.x);
        const int nw = gridDim.x * 10;
        for (int row = blockIdx.x * 10 + warp; row < ph->lnr.rows; row += nw) ln_reduce_row<true>(ph->lnr, row, lane);
      }
    } else {
      // self-attention: a CTA takes one (group, head) at a time; K / V of the head sit in the idle TMA ring, the eight
      // non-issuing warps split the 16-row query tiles
      const int items = G * 8;
      __half* slab = reinterpret_cast<__half*>(smem);
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int g = it >> 3, head = it & 7;
        if (warp >= 2) chain_self_attn_fill(ph->sa, g, head, slab, threadIdx.x - 64, 256);
        __syncthreads();
        if (warp >= 2) chain_self_attn_tiles(ph->sa, g, head, slab, lane, warp - 2, 8);
        __syncthreads();
      }
    }
    if (p + 1 < nphases && !ph->par) chain_grid_barrier(bar, gridDim.x);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace ovis
