// Persistent tcgen05 GEMM for sm_100a:  D[rows, N] = A[rows, K] * B[N, K]^T   (fp16 operands, fp32 accumulate in TMEM)
//
//   * both operands K-major (row-major [rows][K] activations, PyTorch [out][in] weights), staged by TMA into
//     128B-swizzled shared memory, consumed by tcgen05.mma (UMMA 128 x BN x 16) issued by one thread;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 = epilogue (one TMEM lane quarter each);
//   * two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1;
//   * "groups": the M axis is a batch of independent row blocks (frames / clips); B may be a different row block
//     per group (b_group_stride) - this is how the per-frame mask-embed products run as one launch;
//   * fused epilogues (see GemmArgs::epi).
#pragma once
#include "ptx.cuh"

namespace ovis {

// QuickGELU x * sigmoid(1.702 x) (CLIP blocks, mask_adapted_clip/model.py:232-234) with ONE special-function op per element:
// sigmoid(y) = 0.5 * tanh(y / 2) + 0.5 (tanh.approx.f32; relative error ~2^-11, below the fp16 rounding of the stored
// activation).  The exp + reciprocal form costs two MUFU ops and made the fc1 epilogue the slowest of the tower's GEMMs.
__device__ __forceinline__ float quick_gelu(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  return x * fmaf(0.5f, t, 0.5f);
}


enum GemmEpi : int {
  EPI_STORE = 0,      // out[row][col] = act((acc + bias[col]) * scale); fp16 or fp32, smem-staged coalesced stores
  EPI_LN = 1,         // v = acc + bias + resid -> LayerNorm (-> optional 2nd LayerNorm); N == BN == 256
  EPI_SIGNBITS = 2,   // bits[g][r/32][col] = ballot(acc < 0); flags[g][col] = any(acc >= 0)
  EPI_STORE_T = 3,    // out[g*t_group_stride + col*ldt + r] = acc + bias[col]   (fp32, transposed / NCHW-style)
  EPI_SIGNBITS_T = 4, // key-major sign bits for xattn_tc3: bits_t[g][r][col/32] bit col%32 = (acc < 0), flags, 32-key block ANDs
};

#ifdef OVIS_CHAIN_EVTRACE
__device__ long long* g_ev_base = nullptr;    // profiling builds: event slots of the tile being finished by warp 2 (chain.cuh)
#define EPI_EV(sub) do { if (warp == 2 && lane == 0 && blockIdx.x == 0 && g_ev_base) g_ev_base[(sub)] = clock64(); } while (0)
#else
#define EPI_EV(sub) do {} while (0)
#endif

constexpr int GEMM_MAX_NTILES = 16;
constexpr int GEMM_MAX_OUT_MAPS = 16;

// output tensor maps for the TMA-store epilogue (box = 128 bytes x 32 rows, 128B swizzle), one per n-tile
struct GemmOutMaps {
  CUtensorMap m[GEMM_MAX_OUT_MAPS];
};

struct GemmArgs {
  int rows_per_group;     // valid A rows in each group
  int num_groups;
  int a_group_stride;     // A rows between consecutive groups
  int b_group_stride;     // B rows between consecutive groups (0 = one shared B)
  int a_k_offset_stride;  // extra A column offset per (group % a_k_mod) (SAN per-head slices), usually 0
  int a_k_mod;            // see above (>= 1)
  int b_k_offset_stride;  // extra B column (K) offset per (group % b_k_mod): split-K slices of one weight matrix
  int b_k_mod;            // see above (>= 1)
  int o_group_stride;     // EPI_STORE: output rows between consecutive groups (0: outputs follow the A row blocks)
  int a_row_div;          // A row block index = group / a_row_div (>= 1)
  int b_row_div;          // B row block index = group / b_row_div (>= 1)
  int N;                  // valid output columns
  int K;                  // reduction length, multiple of 64
  int epi;
  // ---- EPI_STORE
  void* out[GEMM_MAX_NTILES];          // per n-tile base pointer (column 0 of that tile)
  const float* bias[GEMM_MAX_NTILES];  // per n-tile bias (indexed by column within tile) or null
  int ldo;                // output row stride (elements)
  int out_f32;            // 0: fp16 output, 1: fp32 output
  int relu;               // activation: 0 none, 1 ReLU, 2 QuickGELU x * sigmoid(1.702 x) (CLIP blocks, model.py:232-234)
  float scale;
  const float* resid_st;  // optional fp32 [rows][ldo] added after the activation (residual connections; may alias out)
  int a_alt;              // 1: n-tiles with ((nt >> a_alt_shift) & 1) read their A operand from the second tensor map
  int a_alt_shift;        //    (key / value operands; shift 1 when a 256-wide output is split into two 128-wide tiles)
  int tma_store;          // 1: EPI_STORE writes through GemmOutMaps (single group, 16-byte aligned pitch)
  int a_prefetch;         // B-stationary kernel: A tiles pulled into L2 ahead of the ring (0 = off)
  int kps;                // B-stationary kernel: k-blocks requested together per ring barrier (1, 2 or 4; divides K/64)
  // fused L2-normalisation of the OV heads (kernel 3 of the path): normalize(f) @ text^T = (f @ text^T) / ||f|| row by row
  const float* row_ss_in; // optional [rows]: sum of squares of the A rows (or of the producing GEMM's output rows): the
                          //   store multiplies by scale * rsqrt(max(row_ss_in[row], 1e-24)) instead of scale
  float* row_ss_out;      // optional [rows], zeroed by the caller: += sum of squares of this GEMM's output row (atomicAdd)
  // ---- EPI_LN
  const float* resid;     // [rows][256] fp32
  const float* ln1_g; const float* ln1_b;
  const float* ln2_g; const float* ln2_b;   // null -> no second norm
  const float* pe;        // [pe_period][256] fp32 added for the "+pos" fp16 copy, or null
  int pe_period;
  float* y32; __half* y16; __half* ype16;   // LN1 outputs
  float* d32; __half* d16;                  // LN2 outputs
  // ---- EPI_SIGNBITS
  uint32_t* bits;         // [G][W][q_stride]
  unsigned char* flags;   // [G][q_stride]
  int words_per_group;    // W = ceil(rows_per_group / 32)
  int q_stride;
  // ---- EPI_SIGNBITS_T (flags / words_per_group / q_stride as above)
  uint32_t* bits_t;       // [G][rows_per_group][qw]: word w of a key = its blocked bits for queries 32w..32w+31 (1 past N)
  uint32_t* blockand;     // [G][W][qw]: AND of bits_t over the 32 keys of a block (rows past the group end count as blocked)
  int qw;                 // words per key (4 per 128-query tile)
  // ---- EPI_STORE_T
  float* out_t;
  long long t_group_stride;
  long long ldt;
  unsigned char* posflags;   // optional [frames][q_stride]: 1 when some logit of (frame, col) is > 0 (non-empty mask)
  int rows_per_frame;        // multiple of 32
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 8 * 4096;   // epilogue store staging: 8 warps x [32 rows][128 B]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;   // 256 or 512: power of two
  static constexpr int THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
};

// One 128 x BN accumulator tile: TMEM -> registers -> fused epilogue -> global.  Called by the four epilogue warps.
template <int BN>
// `pf_frame` / `pf_set`: per-thread memo of the posflags already raised by this thread (frame, bit per column chunk)
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& args, const CUtensorMap* omaps, int nt, int mt, int g,
                                                   int warp, int lane, uint32_t t_acc_base, uint8_t* stage_smem,
                                                   long long& pf_frame, uint32_t& pf_set) {
  using Cfg = GemmCfg<BN>;
  const int quarter = warp & 3;           // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;       // which half of the tile's columns this warp handles (warps 2-5: 0, 6-9: 1)
  const int r = mt * Cfg::BM + quarter * 32 + lane;        // row within group
  const bool row_ok = r < args.rows_per_group;
  const long long grow = (long long)(g / args.a_row_div) * args.a_group_stride + r;   // global A/D row
  const int col_base = nt * BN;
  const uint32_t t_acc = t_acc_base + ((uint32_t)(quarter * 32) << 16);
  uint32_t v[32];

  if (args.epi == EPI_STORE) {
    const float* bias = args.bias[nt];
    const int cols_per_unit = args.out_f32 ? 32 : 64;        // 128 bytes of output per row per unit
    const int esize = args.out_f32 ? 4 : 2;
    const uint32_t stg = smem_u32(stage_smem + (warp - 2) * 4096);   // [32 rows][8 x 16 B], XOR-swizzled
    const int r_warp0 = mt * Cfg::BM + quarter * 32;          // first row (within group) of this warp
    const long long grow0 = (args.o_group_stride ? (long long)g * args.o_group_stride
                                                 : (long long)(g / args.a_row_div) * args.a_group_stride) + r_warp0;
    const bool vec_ok = (((long long)args.ldo * esize) & 15) == 0;
    const bool bias_vec = bias && ((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    // Lean path for full column tiles (every GEMM of the decoder): vector bias or none, optional scale / ReLU, fp16 or fp32
    // output, bulk tensor store or row-coalesced stores.  The generic code below handles ragged tiles, residuals, row
    // scales and QuickGELU at run time and costs ~30 instructions per element: 8000 cycles per 128 x 256 tile, which made
    // the HBM-heavy projections epilogue-issue-bound and every small query-side GEMM a 4 us epilogue.
    // (fp32 residual: added in the row-coalesced store loop, where each lane holds 16 contiguous bytes of one row)
    const bool resid_ok = args.resid_st == nullptr ||
                          (args.out_f32 && !args.tma_store && ((reinterpret_cast<uintptr_t>(args.resid_st) & 15) == 0));
    if (resid_ok && !args.row_ss_in && !args.row_ss_out && (bias == nullptr || bias_vec) && col_base + BN <= args.N &&
        (args.tma_store || vec_ok)) {
      const bool do_scale = args.scale != 1.f;
      const float scale = args.scale;
      const bool do_relu = args.relu == 1;
      const bool do_gelu = args.relu == 2;
      const long long row_bytes = (long long)args.ldo * esize;
      const char* const rbase_w = args.resid_st == nullptr ? nullptr
          : reinterpret_cast<const char*>(args.resid_st) + (grow0 + (lane >> 3)) * row_bytes + (long long)nt * BN * esize + (lane & 7) * 16;
      char* const obase_w = reinterpret_cast<char*>(args.out[nt]) + (grow0 + (lane >> 3)) * row_bytes + (lane & 7) * 16;
      const int rows_left = args.rows_per_group - r_warp0 - (lane >> 3);      // this lane stores rows it*4 + (lane >> 3)
#pragma unroll 1
      for (int u0 = half * (BN / 2); u0 < (half + 1) * (BN / 2); u0 += cols_per_unit) {
        uint32_t va[32], vb[32];
        tmem_ld_32x32_raw(t_acc + u0, va);
        if (!args.out_f32) tmem_ld_32x32_raw(t_acc + u0 + 32, vb);
        tmem_ld_wait();
        reg_fence32(va);
        if (!args.out_f32) reg_fence32(vb);
        EPI_EV(1 + 4 * ((u0 / cols_per_unit) & 1));
        if (bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + u0 + 4 * j));
            va[4 * j] = __float_as_uint(__uint_as_float(va[4 * j]) + b0.x);
            va[4 * j + 1] = __float_as_uint(__uint_as_float(va[4 * j + 1]) + b0.y);
            va[4 * j + 2] = __float_as_uint(__uint_as_float(va[4 * j + 2]) + b0.z);
            va[4 * j + 3] = __float_as_uint(__uint_as_float(va[4 * j + 3]) + b0.w);
          }
          if (!args.out_f32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + u0 + 32 + 4 * j));
              vb[4 * j] = __float_as_uint(__uint_as_float(vb[4 * j]) + b0.x);
              vb[4 * j + 1] = __float_as_uint(__uint_as_float(vb[4 * j + 1]) + b0.y);
              vb[4 * j + 2] = __float_as_uint(__uint_as_float(vb[4 * j + 2]) + b0.z);
              vb[4 * j + 3] = __float_as_uint(__uint_as_float(vb[4 * j + 3]) + b0.w);
            }
          }
        }
        if (do_scale) {
#pragma unroll
          for (int j = 0; j < 32; ++j) va[j] = __float_as_uint(__uint_as_float(va[j]) * scale);
          if (!args.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) vb[j] = __float_as_uint(__uint_as_float(vb[j]) * scale);
          }
        }
        if (do_relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) va[j] = __float_as_uint(fmaxf(__uint_as_float(va[j]), 0.f));
          if (!args.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) vb[j] = __float_as_uint(fmaxf(__uint_as_float(vb[j]), 0.f));
          }
        }
        if (do_gelu) {   // QuickGELU x * sigmoid(1.702 x) (CLIP blocks, model.py:232-234)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(va[j]);
            va[j] = __float_as_uint(quick_gelu(x));
          }
          if (!args.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(vb[j]);
              vb[j] = __float_as_uint(quick_gelu(x));
            }
          }
        }
        uint4 pk[8];
        if (args.out_f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = make_uint4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pk[j] = make_uint4(pack_half2(__uint_as_float(va[8 * j]), __uint_as_float(va[8 * j + 1])),
                               pack_half2(__uint_as_float(va[8 * j + 2]), __uint_as_float(va[8 * j + 3])),
                               pack_half2(__uint_as_float(va[8 * j + 4]), __uint_as_float(va[8 * j + 5])),
                               pack_half2(__uint_as_float(va[8 * j + 6]), __uint_as_float(va[8 * j + 7])));
            pk[4 + j] = make_uint4(pack_half2(__uint_as_float(vb[8 * j]), __uint_as_float(vb[8 * j + 1])),
                                   pack_half2(__uint_as_float(vb[8 * j + 2]), __uint_as_float(vb[8 * j + 3])),
                                   pack_half2(__uint_as_float(vb[8 * j + 4]), __uint_as_float(vb[8 * j + 5])),
                                   pack_half2(__uint_as_float(vb[8 * j + 6]), __uint_as_float(vb[8 * j + 7])));
          }
        }
        EPI_EV(2 + 4 * ((u0 / cols_per_unit) & 1));
        if (args.tma_store) {
          if (lane == 0) tma_store_wait_read0();           // the previous unit's bulk store has read the staging buffer
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) st_shared_v4(stg + (uint32_t)((lane * 8 + (c ^ (lane & 7))) << 4), pk[c]);
        EPI_EV(3 + 4 * ((u0 / cols_per_unit) & 1));
        if (args.tma_store) {
          fence_async_proxy();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&omaps[nt], stage_smem + (warp - 2) * 4096, u0, (int)grow0);
            tma_store_commit();
          }
        } else {
          __syncwarp();
          char* o = obase_w + (long long)u0 * esize;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int row = it * 4 + (lane >> 3), c = lane & 7;
            uint4 val = ld_shared_v4(stg + (uint32_t)((row * 8 + (c ^ (row & 7))) << 4));
            if (it * 4 < rows_left) {
              if (rbase_w) {       // (may alias the output: each 16-byte chunk is read and written by the same lane)
                const float4 q4 = *reinterpret_cast<const float4*>(rbase_w + (long long)u0 * esize + it * 4 * row_bytes);
                val.x = __float_as_uint(__uint_as_float(val.x) + q4.x); val.y = __float_as_uint(__uint_as_float(val.y) + q4.y);
                val.z = __float_as_uint(__uint_as_float(val.z) + q4.z); val.w = __float_as_uint(__uint_as_float(val.w) + q4.w);
              }
              *reinterpret_cast<uint4*>(o + it * 4 * row_bytes) = val;
            }
          }
          __syncwarp();
        }
        EPI_EV(4 + 4 * ((u0 / cols_per_unit) & 1));
      }
      return;
    }
#pragma unroll 1
    for (int u0 = half * (BN / 2); u0 < (half + 1) * (BN / 2); u0 += cols_per_unit) {
      if (col_base + u0 >= args.N) break;                     // warp-uniform
      const int ncols = min(cols_per_unit, args.N - (col_base + u0));
      const int nh = args.out_f32 ? 1 : 2;
      uint32_t vv[2][32];
      // issue all TMEM loads of the unit, fetch the bias while they are in flight, wait once
      tmem_ld_32x32_raw(t_acc + u0, vv[0]);
      if (nh == 2) tmem_ld_32x32_raw(t_acc + u0 + 32, vv[1]);
      float4 bv[2][8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh >= nh) break;
        const int nleft = args.N - (col_base + u0 + hh * 32);          // columns past N must not touch bias[]
        if (bias_vec && nleft >= 32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[hh][j] = __ldg(reinterpret_cast<const float4*>(bias + u0 + hh * 32) + j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = 4 * j;
            bv[hh][j] = make_float4((bias && c < nleft) ? __ldg(bias + u0 + hh * 32 + c) : 0.f,
                                    (bias && c + 1 < nleft) ? __ldg(bias + u0 + hh * 32 + c + 1) : 0.f,
                                    (bias && c + 2 < nleft) ? __ldg(bias + u0 + hh * 32 + c + 2) : 0.f,
                                    (bias && c + 3 < nleft) ? __ldg(bias + u0 + hh * 32 + c + 3) : 0.f);
          }
        }
      }
      tmem_ld_wait();
      reg_fence32(vv[0]);
      if (nh == 2) reg_fence32(vv[1]);
      uint4 pk[8];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh >= nh) break;
        float f[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[4 * j] = __uint_as_float(vv[hh][4 * j]) + bv[hh][j].x;
          f[4 * j + 1] = __uint_as_float(vv[hh][4 * j + 1]) + bv[hh][j].y;
          f[4 * j + 2] = __uint_as_float(vv[hh][4 * j + 2]) + bv[hh][j].z;
          f[4 * j + 3] = __uint_as_float(vv[hh][4 * j + 3]) + bv[hh][j].w;
        }
        const bool row_in_range = r_warp0 + lane < args.rows_per_group;
        if (args.row_ss_in) {
          // x / max(||x||, 1e-12) like F.normalize (side_adapter.py:205); adapter.py:118-119 has no clamp: same for real rows
          const float rs = args.scale * rsqrtf(fmaxf(row_in_range ? __ldg(args.row_ss_in + grow0 + lane) : 1.f, 1e-24f));
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= rs;
        } else if (args.scale != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= args.scale;
        }
        if (args.row_ss_out && row_in_range) {
          const int nleft = args.N - (col_base + u0 + hh * 32);
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) ss += (j < nleft) ? f[j] * f[j] : 0.f;
          atomicAdd(args.row_ss_out + grow0 + lane, ss);
        }
        if (args.relu == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (args.relu == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
        }
        if (args.resid_st && r_warp0 + lane < args.rows_per_group) {
          const float* rp = args.resid_st + (grow0 + lane) * args.ldo + (long long)nt * BN + u0 + hh * 32;
          const int nleft = args.N - (col_base + u0 + hh * 32);
          if (nleft >= 32 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q4 = *reinterpret_cast<const float4*>(rp + 4 * j);
              f[4 * j] += q4.x; f[4 * j + 1] += q4.y; f[4 * j + 2] += q4.z; f[4 * j + 3] += q4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nleft) f[j] += rp[j];
          }
        }
        if (args.out_f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pk[j] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                               __float_as_uint(f[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            pk[hh * 4 + j] = make_uint4(pack_half2(f[8 * j], f[8 * j + 1]), pack_half2(f[8 * j + 2], f[8 * j + 3]),
                                        pack_half2(f[8 * j + 4], f[8 * j + 5]), pack_half2(f[8 * j + 6], f[8 * j + 7]));
        }
      }
      // stage through shared memory ([32 rows][128 B], 128B-swizzle pattern)
      if (args.tma_store) {                                   // the previous unit's bulk store must have read the buffer
        if (lane == 0) tma_store_wait_read0();
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) st_shared_v4(stg + (uint32_t)((lane * 8 + (c ^ (lane & 7))) << 4), pk[c]);
      if (args.tma_store) {
        // one bulk tensor store per unit: 32 rows x 128 B, rows / columns past the tensor are clipped by the TMA unit
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&omaps[nt], stage_smem + (warp - 2) * 4096, u0, (int)grow0);
          tma_store_commit();
        }
        continue;
      }
      __syncwarp();
      char* obase = reinterpret_cast<char*>(args.out[nt]) + (long long)u0 * esize;
      if (ncols == cols_per_unit && vec_ok) {
        uint4 val[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3), c = lane & 7;
          val[it] = ld_shared_v4(stg + (uint32_t)((row * 8 + (c ^ (row & 7))) << 4));
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3), c = lane & 7;
          if (r_warp0 + row < args.rows_per_group)
            *reinterpret_cast<uint4*>(obase + (grow0 + row) * args.ldo * esize + c * 16) = val[it];
        }
      } else {
        // ragged tail (N not a multiple of the unit, or unaligned pitch): element-wise, still row-coalesced
        const int epc = 16 / esize;                            // elements per 16 B chunk
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3), c = lane & 7;
          if (r_warp0 + row < args.rows_per_group) {
            const uint4 val = ld_shared_v4(stg + (uint32_t)((row * 8 + (c ^ (row & 7))) << 4));
            char* o = obase + (grow0 + row) * args.ldo * esize + c * 16;
            const uint32_t w[4] = {val.x, val.y, val.z, val.w};
            if (esize == 4) {
              for (int e = 0; e < 4; ++e)
                if (c * epc + e < ncols) reinterpret_cast<uint32_t*>(o)[e] = w[e];
            } else {
              for (int e = 0; e < 8; ++e)
                if (c * epc + e < ncols)
                  reinterpret_cast<unsigned short*>(o)[e] = (unsigned short)((w[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (args.epi == EPI_LN) {
    // BN == 256 == N.  Row-per-thread LayerNorm; the biased+residual row is parked back in TMEM between passes.
    // (row statistics span all 256 columns: only the first warp of each lane quarter works here)
    if (half != 0) return;
    const float* bias = args.bias[0];
    const float* res = args.resid + grow * 256;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      tmem_ld_32x32(t_acc + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        // (__ldcg: inside the query-side chain kernel the residual stream was written earlier by the same kernel)
        float4 rr = row_ok ? __ldcg(reinterpret_cast<const float4*>(res + c * 32 + j)) : make_float4(0, 0, 0, 0);
        float a0 = __uint_as_float(v[j]) + __ldg(bias + c * 32 + j) + rr.x;
        float a1 = __uint_as_float(v[j + 1]) + __ldg(bias + c * 32 + j + 1) + rr.y;
        float a2 = __uint_as_float(v[j + 2]) + __ldg(bias + c * 32 + j + 2) + rr.z;
        float a3 = __uint_as_float(v[j + 3]) + __ldg(bias + c * 32 + j + 3) + rr.w;
        sum += (a0 + a1) + (a2 + a3);
        v[j] = __float_as_uint(a0); v[j + 1] = __float_as_uint(a1);
        v[j + 2] = __float_as_uint(a2); v[j + 3] = __float_as_uint(a3);
      }
      tmem_st_32x32(t_acc + c * 32, v);
    }
    tmem_st_wait();
    const float mean = sum * (1.f / 256.f);
    float sq = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      tmem_ld_32x32(t_acc + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) { float d = __uint_as_float(v[j]) - mean; sq += d * d; }
    }
    const float rstd = rsqrtf(sq * (1.f / 256.f) + 1e-5f);
    const bool two = args.ln2_g != nullptr;
    const float* pe = args.pe ? args.pe + (long long)(r % args.pe_period) * 256 : nullptr;
    float sum2 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      tmem_ld_32x32(t_acc + c * 32, v);
      tmem_ld_wait();
      float y[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        y[j] = (__uint_as_float(v[j]) - mean) * rstd * __ldg(args.ln1_g + c * 32 + j) + __ldg(args.ln1_b + c * 32 + j);
        sum2 += y[j];
      }
      if (row_ok) {
        if (args.y32) {
          float* o = args.y32 + grow * 256 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
        }
        if (args.y16) {
          __half* o = args.y16 + grow * 256 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            u.x = pack_half2(y[j], y[j + 1]); u.y = pack_half2(y[j + 2], y[j + 3]);
            u.z = pack_half2(y[j + 4], y[j + 5]); u.w = pack_half2(y[j + 6], y[j + 7]);
            *reinterpret_cast<uint4*>(o + j) = u;
          }
        }
        if (args.ype16 && pe) {
          __half* o = args.ype16 + grow * 256 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float4 p0 = __ldg(reinterpret_cast<const float4*>(pe + c * 32 + j));
            float4 p1 = __ldg(reinterpret_cast<const float4*>(pe + c * 32 + j + 4));
            uint4 u;
            u.x = pack_half2(y[j] + p0.x, y[j + 1] + p0.y); u.y = pack_half2(y[j + 2] + p0.z, y[j + 3] + p0.w);
            u.z = pack_half2(y[j + 4] + p1.x, y[j + 5] + p1.y); u.w = pack_half2(y[j + 6] + p1.z, y[j + 7] + p1.w);
            *reinterpret_cast<uint4*>(o + j) = u;
          }
        }
      }
      if (two) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(y[j]);
        tmem_st_32x32(t_acc + c * 32, v);
      }
    }
    if (two) {
      tmem_st_wait();
      const float mean2 = sum2 * (1.f / 256.f);
      float sq2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(t_acc + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { float d = __uint_as_float(v[j]) - mean2; sq2 += d * d; }
      }
      const float rstd2 = rsqrtf(sq2 * (1.f / 256.f) + 1e-5f);
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        tmem_ld_32x32(t_acc + c * 32, v);
        tmem_ld_wait();
        if (row_ok) {
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            y[j] = (__uint_as_float(v[j]) - mean2) * rstd2 * __ldg(args.ln2_g + c * 32 + j) + __ldg(args.ln2_b + c * 32 + j);
          if (args.d32) {
            float* o = args.d32 + grow * 256 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          }
          if (args.d16) {
            __half* o = args.d16 + grow * 256 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u;
              u.x = pack_half2(y[j], y[j + 1]); u.y = pack_half2(y[j + 2], y[j + 3]);
              u.z = pack_half2(y[j + 4], y[j + 5]); u.w = pack_half2(y[j + 6], y[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = u;
            }
          }
        }
      }
    }
  } else if (args.epi == EPI_SIGNBITS) {
    // rows = keys / cells, columns = queries.  One ballot per column gives the 32-key word of this warp.
    const int r0 = mt * Cfg::BM + quarter * 32;
    const int word = r0 >> 5;
    const int nvalid = max(0, min(32, args.rows_per_group - r0));
    const uint32_t validmask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
#pragma unroll 1
    for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
      if (col_base + c * 32 >= args.N) break;
      tmem_ld_32x32(t_acc + c * 32, v);
      tmem_ld_wait();
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // blocked  <=>  sigmoid(x) < 0.5  <=>  x < 0 ; rows past the group end count as blocked
        const uint32_t w = __ballot_sync(0xffffffffu, !row_ok || __uint_as_float(v[j]) < 0.f);
        if (lane == j) mine = w;
      }
      const int col = col_base + c * 32 + lane;
      if (col < args.N && word < args.words_per_group) {
        args.bits[((long long)g * args.words_per_group + word) * args.q_stride + col] = mine;
        if ((~mine & validmask) != 0u) {                   // (memo: see posflags below)
          const long long key = (long long)g * GEMM_MAX_NTILES + nt;
          if (key != pf_frame) { pf_frame = key; pf_set = 0u; }
          if (!((pf_set >> c) & 1u)) {
            args.flags[(long long)g * args.q_stride + col] = 1;
            pf_set |= 1u << c;
          }
        }
      }
    }
  } else if (args.epi == EPI_SIGNBITS_T) {
    // rows = keys / cells, columns = queries.  A thread owns one key: its 32 sign bits of a column chunk ARE the word the
    // transposed attention kernel wants (thread = key there as well), so no ballots; the two warp-wide reductions give the
    // per-query "has an unblocked key" flags and the 32-key block ANDs the tile-skip map is built from.
    const int r0 = mt * Cfg::BM + quarter * 32;
#pragma unroll 1
    for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
      const int col0 = col_base + c * 32;
      const int widx = col0 >> 5;
      if (widx >= args.qw) break;
      uint32_t mine = 0xffffffffu;
      if (col0 < args.N) {
        tmem_ld_32x32(t_acc + c * 32, v);
        tmem_ld_wait();
        mine = 0u;
#pragma unroll
        for (int j = 0; j < 32; ++j)   // blocked  <=>  sigmoid(x) < 0.5  <=>  x < 0
          mine |= (__uint_as_float(v[j]) < 0.f ? 1u : 0u) << j;
        const int nv = args.N - col0;
        if (nv < 32) mine |= ~((1u << nv) - 1u);              // queries past N: blocked
      }
      if (row_ok) args.bits_t[((long long)g * args.rows_per_group + r) * args.qw + widx] = mine;
      const uint32_t un = __reduce_or_sync(0xffffffffu, row_ok ? ~mine : 0u);
      const uint32_t an = __reduce_and_sync(0xffffffffu, row_ok ? mine : 0xffffffffu);
      if (lane == 0 && r0 < args.rows_per_group)
        args.blockand[((long long)g * args.words_per_group + (r0 >> 5)) * args.qw + widx] = an;
      const int col = col0 + lane;
      if (col < args.N && ((un >> lane) & 1u)) {
        const long long key = (long long)g * GEMM_MAX_NTILES + nt;
        if (key != pf_frame) { pf_frame = key; pf_set = 0u; }
        if (!((pf_set >> c) & 1u)) {
          args.flags[(long long)g * args.q_stride + col] = 1;
          pf_set |= 1u << c;
        }
      }
    }
  } else {  // EPI_STORE_T
    const float* bias = args.bias[nt];
    const uint32_t stg = smem_u32(stage_smem + (warp - 2) * 4096);     // [32 columns][32 rows] fp32, one chunk
    const int r_warp0 = mt * Cfg::BM + quarter * 32;
#pragma unroll 1
    for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
      if (col_base + c * 32 >= args.N) break;
      tmem_ld_32x32(t_acc + c * 32, v);
      tmem_ld_wait();
      if (bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col_base + c * 32 + j < args.N) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(bias + c * 32 + j));
      }
      if (args.tma_store) {
        // transposed through shared memory (lane = row: conflict-free 4-byte stores), then ONE bulk tensor store of
        // 32 columns x 128 bytes; columns >= N and rows past the group are clipped by the TMA unit
        if (lane == 0) tma_store_wait_read0();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + (uint32_t)(j * 128 + lane * 4)), "r"(v[j]) : "memory");
        fence_async_proxy();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&omaps[0], stage_smem + (warp - 2) * 4096, r_warp0, col_base + c * 32, g);
          tma_store_commit();
        }
      } else if (row_ok) {
        const int ncols = min(32, args.N - (col_base + c * 32));
        float* op = args.out_t + (long long)g * args.t_group_stride + (long long)(col_base + c * 32) * args.ldt + r;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < ncols) __stcs(op, __uint_as_float(v[j]));
          op += args.ldt;
        }
      }
      if (args.posflags) {
        // all 32 rows of this warp belong to one frame (rows_per_frame % 32 == 0)
        // sign bits of the row's 32 logits shifted into one word (one funnel shift per logit: bit 31-j = logit j is
        // negative), then AND-reduced over the 32 rows of the warp: a clear bit = some row has a non-negative logit.
        // (+0.0 counts as positive here; an exactly-zero fp32 accumulation does not occur on real features.)
        uint32_t neg = 0xffffffffu;
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) neg = __funnelshift_l(v[j], neg, 1);
        }
        const uint32_t allneg = __reduce_and_sync(0xffffffffu, neg);
        const int col = col_base + c * 32 + lane;
        const int r0 = mt * Cfg::BM + quarter * 32;
        if (col < args.N && !((allneg >> (31 - lane)) & 1u) && r0 < args.rows_per_group) {
          // plain store, no read-back (a load here would sit on the epilogue's critical path); the per-thread memo
          // keeps the thousands of tiles of one frame from re-writing the same byte
          const long long frame = (long long)g * (args.rows_per_group / args.rows_per_frame) + r0 / args.rows_per_frame;
          const long long key = frame * GEMM_MAX_NTILES + nt;
          if (key != pf_frame) { pf_frame = key; pf_set = 0u; }
          if (!((pf_set >> c) & 1u)) {
            args.posflags[frame * args.q_stride + col] = 1;
            pf_set |= 1u << c;
          }
        }
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(320, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmOutMaps om, const GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
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

  const int m_tiles = (args.rows_per_group + Cfg::BM - 1) / Cfg::BM;
  const int n_tiles = (args.N + BN - 1) / BN;
  const int total_tiles = args.num_groups * m_tiles * n_tiles;
  const int k_blocks = args.K / Cfg::BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
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
  pdl_begin();   // (the prologue above overlapped the previous kernel's tail; no global read before this point)
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        const int mt = (tile / n_tiles) % m_tiles;
        const int g = tile / (n_tiles * m_tiles);
        const int a_row = (g / args.a_row_div) * args.a_group_stride + mt * Cfg::BM;
        const int a_col = (g % args.a_k_mod) * args.a_k_offset_stride;
        const int b_row = (g / args.b_row_div) * args.b_group_stride + nt * BN;
        const int b_col = (g % args.b_k_mod) * args.b_k_offset_stride;
        const CUtensorMap* ta = (args.a_alt && ((nt >> args.a_alt_shift) & 1)) ? &tmA2 : &tmA;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, ta, &full_bar[stage], a_col + kb * Cfg::BK, a_row);
          tma_load_2d(sb, &tmB, &full_bar[stage], b_col + kb * Cfg::BK, b_row);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);       // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    int acc = 0;
    uint32_t acc_phase = 0;
    long long pf_frame = -1;
    uint32_t pf_set = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % n_tiles;
      const int mt = (tile / n_tiles) % m_tiles;
      const int g = tile / (n_tiles * m_tiles);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<BN>(args, om.m, nt, mt, g, warp, lane, tmem_base + acc * BN, stage_smem, pf_frame, pf_set);
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (args.tma_store && lane == 0) tma_store_wait_read0();  // bulk stores read shared memory until then
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// B-stationary variant for K <= 256 (every HBM-heavy GEMM of this path: K/V projection, mask bits, full-resolution
// mask logits, SAN bias branch).  A CTA keeps one [BN x K] B tile (weights / mask embeddings of one group) resident in
// shared memory and streams only A tiles through the TMA ring, so the ring is twice as deep for the same shared
// memory and B is read from L2 once per CTA instead of once per tile.  Work decomposition: "columns" = (group, n-tile);
// the grid is laid out as ncols x R CTAs, CTA (col, r) walks m-tiles r, r+R, ... of its column(s), so the CTAs that
// share an A tile (same m-tile, different n-tile) run at the same time and the tile is fetched from HBM once.
template <int BN>
struct GemmBsCfg {
  static constexpr int BM = 128, BK = 64, KB_MAX = 4;           // K <= 256
  static constexpr int B_BYTES = BN * BK * 2 * KB_MAX;           // 128 KB (BN 256) / 64 KB (BN 128)
  static constexpr int A_BYTES = BM * BK * 2;                    // 16 KB per ring stage
  static constexpr int STAGES = (BN == 256) ? 4 : 8;
  static constexpr int STAGING_BYTES = 8 * 4096;
  static constexpr int SMEM_BYTES = B_BYTES + STAGES * A_BYTES + STAGING_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
};

template <int BN>
__global__ void __launch_bounds__(320, 1)
gemm_tn_bs_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmOutMaps om, const GemmArgs args) {
  using Cfg = GemmBsCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = smem + Cfg::B_BYTES;
  uint8_t* stage_smem = sA + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_smem + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;        // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;   // [2]
  uint64_t* bfull_bar = bars + 2 * STAGES + 4;    // [1]
  uint64_t* bempty_bar = bars + 2 * STAGES + 5;   // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (args.rows_per_group + Cfg::BM - 1) / Cfg::BM;
  const int n_tiles = (args.N + BN - 1) / BN;
  const int cols = args.num_groups * n_tiles;
  const int k_blocks = args.K / Cfg::BK;
  const int ncols_cta = min(cols, (int)gridDim.x);      // columns processed concurrently
  const int R = (int)gridDim.x / ncols_cta;             // CTAs per column
  const int col0 = (int)blockIdx.x % ncols_cta;
  const int r0 = (int)blockIdx.x / ncols_cta;
  const bool active = r0 < R;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8); }
    mbar_init(bfull_bar, 1);
    mbar_init(bempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (active) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0, bphase = 0;
        for (int col = col0; col < cols; col += ncols_cta) {
          const int nt = col % n_tiles;
          const int g = col / n_tiles;
          const int b_row = (g / args.b_row_div) * args.b_group_stride + nt * BN;
          const int a_col = (g % args.a_k_mod) * args.a_k_offset_stride;
          const int a_row0 = (g / args.a_row_div) * args.a_group_stride;
          const CUtensorMap* ta = (args.a_alt && ((nt >> args.a_alt_shift) & 1)) ? &tmA2 : &tmA;
          mbar_wait(bempty_bar, bphase ^ 1);               // previous column's MMAs have retired
          mbar_arrive_expect_tx(bfull_bar, (uint32_t)(k_blocks * BN * Cfg::BK * 2));
          const int b_col = (g % args.b_k_mod) * args.b_k_offset_stride;
          for (int kb = 0; kb < k_blocks; ++kb)
            tma_load_2d(sB + kb * (BN * Cfg::BK * 2), &tmB, bfull_bar, b_col + kb * Cfg::BK, b_row);
          bphase ^= 1;
          // the ring's barriers work on groups of `kps` k-block slots: the boxes of a group (the 128-byte column slices of
          // the same 128 rows) are requested back to back, so DRAM sees whole 512-byte rows instead of four visits
          const int kps = args.kps, ngroups = STAGES / kps;
          // Optional (A/B, OVIS_GEMM_PF): the A tiles this CTA will need are pulled into L2 `pf` tiles ahead of the ring.
          // Measured: no gain for kv_proj, a loss for mask_logits (profiles/experiments/gemm_bs_r2.md), so the default is 0.
          const int pf = args.a_prefetch;
          for (int d = 0; d < pf; ++d)
            if (r0 + d * R < m_tiles)
              for (int kb = 0; kb < k_blocks; ++kb) tma_prefetch_2d(ta, a_col + kb * Cfg::BK, a_row0 + (r0 + d * R) * Cfg::BM);
          for (int mt = r0; mt < m_tiles; mt += R) {
            if (pf > 0 && mt + pf * R < m_tiles)
              for (int kb = 0; kb < k_blocks; ++kb) tma_prefetch_2d(ta, a_col + kb * Cfg::BK, a_row0 + (mt + pf * R) * Cfg::BM);
            for (int kb = 0; kb < k_blocks; kb += kps) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(kps * Cfg::A_BYTES));
              for (int i = 0; i < kps; ++i)
                tma_load_2d(sA + (stage * kps + i) * Cfg::A_BYTES, ta, &full_bar[stage], a_col + (kb + i) * Cfg::BK,
                            a_row0 + mt * Cfg::BM);
              if (++stage == ngroups) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0, bphase = 0;
        const uint32_t b_addr = smem_u32(sB);
        for (int col = col0; col < cols; col += ncols_cta) {
          mbar_wait(bfull_bar, bphase);
          bphase ^= 1;
          tc_fence_after();
          for (int mt = r0; mt < m_tiles; mt += R) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            const int kps = args.kps, ngroups = STAGES / kps;
            for (int kb = 0; kb < k_blocks; kb += kps) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              for (int i = 0; i < kps; ++i) {
                const uint64_t adesc = umma_desc_k_sw128(smem_u32(sA + (stage * kps + i) * Cfg::A_BYTES));
                const uint64_t bdesc = umma_desc_k_sw128(b_addr + (kb + i) * (BN * Cfg::BK * 2));
#pragma unroll
                for (int k = 0; k < Cfg::BK / 16; ++k)
                  umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb + i) | k) != 0 ? 1u : 0u);
              }
              umma_commit(&empty_bar[stage]);
              if (++stage == ngroups) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
          umma_commit(bempty_bar);                       // B tile may be overwritten once everything above retired
        }
      }
    } else {
      // ===================== epilogue warps (2..5) =====================
      int acc = 0;
      uint32_t acc_phase = 0;
      long long pf_frame = -1;
      uint32_t pf_set = 0;
      for (int col = col0; col < cols; col += ncols_cta) {
        const int nt = col % n_tiles;
        const int g = col / n_tiles;
        for (int mt = r0; mt < m_tiles; mt += R) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          gemm_epilogue_tile<BN>(args, om.m, nt, mt, g, warp, lane, tmem_base + acc * BN, stage_smem, pf_frame, pf_set);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (args.tma_store && lane == 0) tma_store_wait_read0();  // bulk stores read shared memory until then
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace ovis
