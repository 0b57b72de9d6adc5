// Masked cross-attention of the Mask2Former-style decoder (reference: CrossAttentionLayer.forward_post,
// video_mask2former_transformer_decoder.py:110-122 -> nn.MultiheadAttention with a bool attn_mask).
//
// Flash-style, split over the key axis so that a single clip with T*HW keys (Video decoder) still fills
// the GPU: grid = (splits, heads, groups).  The mask predicate is generated in-kernel from the packed
// sign bits written by the mask-head GEMM epilogue (one bit per (query, key), shared by all heads) and
// from the per-row "has an unblocked key" flag that implements the reference's
// `attn_mask[where(attn_mask.sum(-1) == N)] = False` rule (frame_..._decoder.py:87) without a host sync.
//
// Round-1 kernel: mma.sync.m16n8k16 (fp16 in, fp32 accumulate), cp.async double-buffered K/V tiles.
#pragma once
#include "ptx.cuh"

namespace ovis {

struct XattnArgs {
  const __half* q;        // [G*Q][256]  already multiplied by d^-1/2 * log2(e)
  const __half* k;        // [G*keys][256]
  const __half* v;        // [G*keys][256]
  const uint32_t* bits;   // [G][W][q_stride]   bit = 1 -> blocked
  const unsigned char* flags;  // [G][q_stride]  1 -> row has at least one unblocked key
  float* o_part;          // [G][S][8][q_pad][32]  un-normalised
  float* ml_part;         // [G][S][8][q_pad][2]   (running max in log2 domain, running sum)
  int Q, q_pad, q_stride;
  int keys;               // keys per group
  int W;                  // words per group = ceil(keys/32)
  int splits;
  int chunk;              // keys per split, multiple of 64
};

constexpr int XA_KT = 64;        // keys per tile
constexpr int XA_LD = 40;        // smem row stride in halves (80 B: conflict-free for ldmatrix and 32-bit frag loads)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(smem_row)));
}

// block = ceil(Q/32) warps; each warp owns 32 query rows (two m16 tiles) of one head.
__global__ void __launch_bounds__(256)
xattn_split_kernel(const XattnArgs a) {
  __shared__ __align__(16) __half sK[2][XA_KT][XA_LD];
  __shared__ __align__(16) __half sV[2][XA_KT][XA_LD];

  const int split = blockIdx.x, head = blockIdx.y, g = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane >> 2, tq = lane & 3;
  const int k_begin = split * a.chunk;
  const int k_end = min(k_begin + a.chunk, a.keys);
  const int ntiles = (k_end - k_begin + XA_KT - 1) / XA_KT;

  // ---- Q fragments (rows >= Q read as zero)
  uint32_t qf[2][2][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = warp * 32 + mt * 16 + quad + (i & 1) * 8;
        const int col = head * 32 + ks * 16 + tq * 2 + (i >> 1) * 8;
        qf[mt][ks][i] = row < a.Q ? *reinterpret_cast<const uint32_t*>(a.q + ((long long)g * a.Q + row) * 256 + col) : 0u;
      }
    }
  }
  // per-thread rows: (mt, hi) -> row = warp*32 + mt*16 + quad + hi*8
  bool rflag[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int row = warp * 32 + mt * 16 + quad + hi * 8;
      rflag[mt][hi] = row < a.Q ? (a.flags[(long long)g * a.q_stride + row] != 0) : false;
    }

  float o[2][4][4];
  float mrow[2][2], lrow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int dn = 0; dn < 4; ++dn)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][dn][i] = 0.f;
    mrow[mt][0] = mrow[mt][1] = -INFINITY;
    lrow[mt][0] = lrow[mt][1] = 0.f;
  }

  const __half* kbase = a.k + (long long)g * a.keys * 256 + head * 32;
  const __half* vbase = a.v + (long long)g * a.keys * 256 + head * 32;
  auto load_tile = [&](int t, int buf) {
    const int kb = k_begin + t * XA_KT;
    for (int c = threadIdx.x; c < 2 * XA_KT * 4; c += blockDim.x) {
      const int which = c / (XA_KT * 4);
      const int rc = c % (XA_KT * 4);
      const int row = rc >> 2, ch = rc & 3;
      const int key = kb + row;
      const bool ok = key < k_end;
      const __half* src = (which ? vbase : kbase) + (long long)(ok ? key : 0) * 256 + ch * 8;
      __half* dst = which ? &sV[buf][row][ch * 8] : &sK[buf][row][ch * 8];
      cp_async16(dst, src, ok);
    }
    cp_async_commit();
  };

  if (ntiles > 0) load_tile(0, 0);
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_tile(t + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    const int kb = k_begin + t * XA_KT;
    // mask words of this 64-key tile for the thread's four rows
    uint32_t mw[2][2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int row = warp * 32 + mt * 16 + quad + hi * 8;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int wi = (kb >> 5) + w;
          uint32_t bitsw = 0u;
          if (rflag[mt][hi] && wi < a.W) bitsw = __ldg(a.bits + ((long long)g * a.W + wi) * a.q_stride + row);
          // keys past the end of the split are always blocked
          const int nvalid = k_end - (kb + w * 32);
          const uint32_t inval = nvalid >= 32 ? 0u : (nvalid <= 0 ? 0xffffffffu : ~((1u << nvalid) - 1u));
          mw[mt][hi][w] = bitsw | inval;
        }
      }

#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sK[buf][nt * 8 + quad][ks * 16 + tq * 2]);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sK[buf][nt * 8 + quad][ks * 16 + tq * 2 + 8]);
          mma_16816(s[nt], qf[mt][ks], b0, b1);
        }
      }
      // mask + row max
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int hi = i >> 1;
          const int kk = nt * 8 + tq * 2 + (i & 1);
          const bool blocked = (mw[mt][hi][kk >> 5] >> (kk & 31)) & 1u;
          if (blocked) s[nt][i] = -INFINITY;
          mx[hi] = fmaxf(mx[hi], s[nt][i]);
        }
      }
      float corr[2], muse[2];
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 1));
        mx[hi] = fmaxf(mx[hi], __shfl_xor_sync(0xffffffffu, mx[hi], 2));
        const float mnew = fmaxf(mrow[mt][hi], mx[hi]);
        muse[hi] = (mnew == -INFINITY) ? 0.f : mnew;
        corr[hi] = exp2f(mrow[mt][hi] - muse[hi]);     // 0 when the old max was -inf
        mrow[mt][hi] = mnew;
      }
      float ls[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = exp2f(s[nt][i] - muse[i >> 1]);
          s[nt][i] = p;
          ls[i >> 1] += p;
        }
      }
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        ls[hi] += __shfl_xor_sync(0xffffffffu, ls[hi], 1);
        ls[hi] += __shfl_xor_sync(0xffffffffu, ls[hi], 2);
        lrow[mt][hi] = lrow[mt][hi] * corr[hi] + ls[hi];
      }
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        o[mt][dn][0] *= corr[0]; o[mt][dn][1] *= corr[0];
        o[mt][dn][2] *= corr[1]; o[mt][dn][3] *= corr[1];
      }
      // O += P V
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t pa[4];
        pa[0] = pack_half2(s[2 * j][0], s[2 * j][1]);
        pa[1] = pack_half2(s[2 * j][2], s[2 * j][3]);
        pa[2] = pack_half2(s[2 * j + 1][0], s[2 * j + 1][1]);
        pa[3] = pack_half2(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(b0, b1, &sV[buf][j * 16 + (lane & 15)][dn * 8]);
          mma_16816(o[mt][dn], pa, b0, b1);
        }
      }
    }
    __syncthreads();
  }

  // ---- write the split's partial result
  const long long pbase = (((long long)g * a.splits + split) * 8 + head) * a.q_pad;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int row = warp * 32 + mt * 16 + quad + hi * 8;
      if (row < a.q_pad) {
        float* op = a.o_part + (pbase + row) * 32;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
          *reinterpret_cast<float2*>(op + dn * 8 + tq * 2) = make_float2(o[mt][dn][hi * 2], o[mt][dn][hi * 2 + 1]);
        if (tq == 0) *reinterpret_cast<float2*>(a.ml_part + (pbase + row) * 2) = make_float2(mrow[mt][hi], lrow[mt][hi]);
      }
    }
}

// Merge the key splits: out[g*Q+q][h*32+d] = sum_s 2^(m_s-M) O_s / sum_s 2^(m_s-M) l_s   (fp16 for the out-proj GEMM)
// grid (Q, 8 heads, G), 256 threads = 8 split lanes x 32 channels; each lane folds its splits online, then a smem merge.
__global__ void __launch_bounds__(256)
xattn_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part, __half* __restrict__ out,
                     int Q, int q_pad, int splits) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  __shared__ float sm_m[8], sm_den[8], sm_num[8][32];
  const int q = blockIdx.x, h = blockIdx.y, g = blockIdx.z;
  const int d = threadIdx.x & 31, sl = threadIdx.x >> 5;
  float M = -INFINITY, num = 0.f, den = 0.f;
  for (int s = sl; s < splits; s += 8) {
    const long long r = (((long long)g * splits + s) * 8 + h) * q_pad + q;
    const float2 ml = __ldg(reinterpret_cast<const float2*>(ml_part + r * 2));
    const float o = __ldg(o_part + r * 32 + d);
    if (ml.x == -INFINITY) continue;
    const float Mn = fmaxf(M, ml.x);
    const float c0 = exp2f(M - Mn), c1 = exp2f(ml.x - Mn);      // exp2f(-inf) = 0 on the first contribution
    num = num * c0 + o * c1;
    den = den * c0 + ml.y * c1;
    M = Mn;
  }
  sm_num[sl][d] = num;
  if (d == 0) { sm_m[sl] = M; sm_den[sl] = den; }
  __syncthreads();
  if (sl == 0) {
    float Mt = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) Mt = fmaxf(Mt, sm_m[i]);
    float n = 0.f, dn = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float w = (sm_m[i] == -INFINITY) ? 0.f : exp2f(sm_m[i] - Mt);
      n += w * sm_num[i][d];
      dn += w * sm_den[i];
    }
    out[((long long)g * Q + q) * 256 + h * 32 + d] = __float2half_rn(n / dn);
  }
}

// The same merge when there are only a few partials per row (Frame decoders: `splits` = 2-8): one thread per pair of
// output channels walks the partials serially, so no thread idles and nothing goes through shared memory
// (xattn_combine_kernel's 8 split lanes x 32 channels layout leaves 6 of 8 warps without work at splits = 2).
__global__ void __launch_bounds__(256)
xattn_combine_few_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part, __half* __restrict__ out,
                         int Q, int q_pad, int splits, long long total) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // ((g * Q + q) * 8 + h) * 16 + d2
  if (e >= total) return;
  const int d2 = (int)(e & 15), h = (int)((e >> 4) & 7);
  const long long row = e >> 7;
  const int q = (int)(row % Q);
  const long long g = row / Q;
  float M = -INFINITY, n0 = 0.f, n1 = 0.f, den = 0.f;
  for (int s = 0; s < splits; ++s) {
    const long long r = ((g * splits + s) * 8 + h) * q_pad + q;
    const float2 ml = __ldg(reinterpret_cast<const float2*>(ml_part + r * 2));
    const float2 o = __ldg(reinterpret_cast<const float2*>(o_part + r * 32 + d2 * 2));
    if (ml.x == -INFINITY) continue;
    const float Mn = fmaxf(M, ml.x);
    const float c0 = exp2f(M - Mn), c1 = exp2f(ml.x - Mn);
    n0 = n0 * c0 + o.x * c1;
    n1 = n1 * c0 + o.y * c1;
    den = den * c0 + ml.y * c1;
    M = Mn;
  }
  *reinterpret_cast<__half2*>(out + row * 256 + h * 32 + d2 * 2) = __floats2half2_rn(n0 / den, n1 / den);
}

// Self-attention over the Q object queries (SelfAttentionLayer.forward_post, video_..._decoder.py:52-62).
// One CTA per (head, group); thread = query row; K/V of the head staged in shared memory.
struct SelfAttnArgs {
  const __half* qk;   // [G*Q][512]: q = cols [0,256), k = cols [256,512)  (biases included, unscaled)
  const __half* v;    // [G*Q][256]
  __half* out;        // [G*Q][256]
  int Q;
  float scale_log2;   // d^-1/2 * log2(e)
};

// PARTS lanes per query row (keys j = part, part + PARTS, ...), each with its own online softmax; the partial
// (max, sum, acc[32]) are merged with shuffles.  512 threads: PARTS = 4 for Q <= 128, PARTS = 2 for Q <= 256.
template <int PARTS>
__global__ void __launch_bounds__(512)
self_attn_kernel(const SelfAttnArgs a) {
  extern __shared__ __half sm[];       // K [Q][32], V [Q][32]
  __half* sk = sm;
  __half* sv = sm + a.Q * 32;
  const int head = blockIdx.x, g = blockIdx.y;
  for (int i = threadIdx.x; i < a.Q * 4; i += blockDim.x) {
    const int row = i >> 2, ch = i & 3;
    const long long r = (long long)g * a.Q + row;
    *reinterpret_cast<uint4*>(sk + row * 32 + ch * 8) = *reinterpret_cast<const uint4*>(a.qk + r * 512 + 256 + head * 32 + ch * 8);
    *reinterpret_cast<uint4*>(sv + row * 32 + ch * 8) = *reinterpret_cast<const uint4*>(a.v + r * 256 + head * 32 + ch * 8);
  }
  __syncthreads();
  const int part = threadIdx.x % PARTS;
  // rows in passes of 512 / PARTS (one pass for the decoders' Q <= 256; the resampler's frame axis can be longer);
  // every thread runs every pass so that the shuffles below stay uniform
  for (int row = threadIdx.x / PARTS; row < ((a.Q + 512 / PARTS - 1) / (512 / PARTS)) * (512 / PARTS); row += 512 / PARTS) {
  const bool row_ok = row < a.Q;                 // (whole lane groups are in or out)
  float q[32], acc[32];
  {
    const __half* qp = a.qk + ((long long)g * a.Q + (row_ok ? row : 0)) * 512 + head * 32;
#pragma unroll
    for (int d = 0; d < 32; d += 2) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(qp + d));
      q[d] = f.x * a.scale_log2; q[d + 1] = f.y * a.scale_log2;
      acc[d] = acc[d + 1] = 0.f;
    }
  }
  float m = -INFINITY, l = 0.f;
  for (int j = part; j < a.Q; j += PARTS) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int d = 0; d < 32; d += 2) {
      const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(sk + j * 32 + d));
      s0 = fmaf(q[d], kf.x, s0);
      s1 = fmaf(q[d + 1], kf.y, s1);
    }
    const float s = s0 + s1;
    const float mn = fmaxf(m, s);
    const float c = exp2f(m - mn);
    const float p = exp2f(s - mn);
    l = l * c + p;
#pragma unroll
    for (int d = 0; d < 32; d += 2) {
      const float2 vf = __half22float2(*reinterpret_cast<const __half2*>(sv + j * 32 + d));
      acc[d] = fmaf(p, vf.x, acc[d] * c);
      acc[d + 1] = fmaf(p, vf.y, acc[d + 1] * c);
    }
    m = mn;
  }
  // merge the key partitions of the row (adjacent lanes of the warp)
  float mt = m;
#pragma unroll
  for (int o = 1; o < PARTS; o <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
  const float sc = (m == -INFINITY) ? 0.f : exp2f(m - mt);
  l *= sc;
#pragma unroll
  for (int o = 1; o < PARTS; o <<= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float inv = 1.f / l;
#pragma unroll
  for (int d = 0; d < 32; ++d) {
    float v = acc[d] * sc;
#pragma unroll
    for (int o = 1; o < PARTS; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[d] = v * inv;
  }
  if (row_ok) {      // each lane of the group writes its share of the 32 channels
    constexpr int CH = 32 / PARTS;       // 8 or 16
    __half* op = a.out + ((long long)g * a.Q + row) * 256 + head * 32 + part * CH;
    // (acc[] is indexed with compile-time constants only: pick the lane's values with selects)
#pragma unroll
    for (int e = 0; e < CH; e += 8) {
      float o[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        float v = acc[e + x];
#pragma unroll
        for (int pp = 1; pp < PARTS; ++pp) v = (part == pp) ? acc[pp * CH + e + x] : v;
        o[x] = v;
      }
      uint4 u;
      u.x = pack_half2(o[0], o[1]); u.y = pack_half2(o[2], o[3]); u.z = pack_half2(o[4], o[5]); u.w = pack_half2(o[6], o[7]);
      *reinterpret_cast<uint4*>(op + e) = u;
    }
  }
  }   // row passes
}

// Tensor-core version (mma.sync m16n8k16, fp16 operands / fp32 accumulation): one CTA per (head, group), K and V of
// the head in shared memory, each of the 8 warps takes 16-row query tiles round-robin and runs a flash-style loop over
// 64-key blocks.  The SIMT kernel above needed 35 us for 36 frames x 100 queries (10 TFLOP/s) and 98 us at 200 queries:
// nine of those per decoder call were 12 % of the Frame decoders' time.  Same interface and scaling as self_attn_kernel
// (which stays as the checker: OVIS_SELF_ATTN=simt).
constexpr int SA_KB = 64;
__global__ void __launch_bounds__(256)
self_attn_mma_kernel(const SelfAttnArgs a) {
  pdl_begin();   // programmatic dependent launch: scheduled while the previous kernel drains, reads nothing before this
  extern __shared__ __align__(16) __half sm_sa[];          // K [Qp][XA_LD], V [Qp][XA_LD]; Qp = Q rounded up to 64
  const int Q = a.Q;
  const int Qp = (Q + SA_KB - 1) / SA_KB * SA_KB;
  __half* sk = sm_sa;
  __half* sv = sm_sa + (size_t)Qp * XA_LD;
  const int head = blockIdx.x, g = blockIdx.y;
  for (int i = threadIdx.x; i < Qp * 4; i += blockDim.x) {
    const int row = i >> 2, ch = i & 3;
    uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
    if (row < Q) {
      const long long r = (long long)g * Q + row;
      kv = *reinterpret_cast<const uint4*>(a.qk + r * 512 + 256 + head * 32 + ch * 8);
      vv = *reinterpret_cast<const uint4*>(a.v + r * 256 + head * 32 + ch * 8);
    }
    *reinterpret_cast<uint4*>(sk + row * XA_LD + ch * 8) = kv;
    *reinterpret_cast<uint4*>(sv + row * XA_LD + ch * 8) = vv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = lane >> 2, tq = lane & 3;
  const int mtiles = (Q + 15) >> 4;
  for (int mt = warp; mt < mtiles; mt += 8) {
    uint32_t qf[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = mt * 16 + quad + (i & 1) * 8;
        const int col = head * 32 + ks * 16 + tq * 2 + (i >> 1) * 8;
        qf[ks][i] = row < Q ? *reinterpret_cast<const uint32_t*>(a.qk + ((long long)g * Q + row) * 512 + col) : 0u;
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
      // scale to base-2 logits, drop the padding keys, row max (rows quad and quad + 8; a quad of lanes shares a row)
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
        const float mnew = fmaxf(mrow[hi], mx[hi]);          // finite: every block holds at least one real key
        corr[hi] = exp2f(mrow[hi] - mnew);                   // 0 for the first block
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
      for (int hi = 0; hi < 2; ++hi) lrow[hi] = lrow[hi] * corr[hi] + ls[hi];      // per-lane partial sums; merged at the end
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

}  // namespace ovis
