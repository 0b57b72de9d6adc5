// Temporal association kernels: cosine-cost query matching of consecutive frames
// (openvis/modeling/minvis.py:28-72) and the replicate-padded unfold that turns the resampler's temporal Conv1d
// (openvis/modeling/resampler.py:205-213) into a GEMM.
//
// Matching.  The reference walks a clip frame by frame: frame i is matched (scipy linear_sum_assignment) against frame
// i-1 *re-ordered by the previous match*, so its T solves are serial and each costs a device->host copy.  Re-ordering
// the rows of a cost matrix only re-labels them, so the optimum of frame i in the slots of the re-ordered frame i-1 is
// sigma_i[j] = pi_i[sigma_{i-1}[j]], where pi_i solves the problem between the *raw* frames i-1 and i.  All B*T raw
// problems are therefore independent: one CTA each (match_assign_kernel), followed by a trivially parallel composition
// of the permutations (match_compose_kernel).  Nothing leaves the device.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ovis {

// out[g][t][kk*C + c] = in[g][clamp(t + kk - taps/2, 0, T-1)][c]   (Conv1d padding='same', padding_mode='replicate')
__global__ void __launch_bounds__(256)
temporal_unfold_kernel(const __half* __restrict__ in, __half* __restrict__ out, int T, int C8, int taps, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk of the output
  if (e >= total) return;
  const int c = (int)(e % C8);
  long long r = e / C8;
  const int kk = (int)(r % taps);
  r /= taps;                                                                   // r = g * T + t
  const int t = (int)(r % T);
  int ts = t + kk - taps / 2;
  ts = ts < 0 ? 0 : (ts >= T ? T - 1 : ts);
  const uint4 v = reinterpret_cast<const uint4*>(in)[(r - t + ts) * C8 + c];
  reinterpret_cast<uint4*>(out)[e] = v;
}

struct MatchArgs {
  const float* en;     // [B][T][n][C] L2-normalised embeddings
  float* cost;         // [B][T][n][n] or null: cost[a][k] = 1 - <en[i-1][a], en[i][k]> (frame 0 against itself)
  int* pi;             // [B][T][n]: query of frame i assigned to query a of frame i-1
  int T, n, C, cost_in_smem;
};

constexpr int MATCH_THREADS = 256;
constexpr int MATCH_TILE = 128;       // cost tile per pass, 8 x 8 entries per thread
constexpr int MATCH_KC = 32;          // channels staged per step

__host__ __device__ inline size_t match_smem_bytes(int n, int cost_in_smem) {
  size_t b = (size_t)2 * MATCH_KC * (MATCH_TILE + 1) * sizeof(float);          // staging
  b += (size_t)(n + 1) * (3 * sizeof(double) + 2 * sizeof(int)) + 8;           // u, v, minv, p, way
  b += ((size_t)(n + 1) + 7) / 8 * 8;                                          // used
  if (cost_in_smem) b += (size_t)n * n * sizeof(float);
  return b;
}

__global__ void __launch_bounds__(MATCH_THREADS)
match_assign_kernel(const MatchArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int n = a.n, C = a.C;
  const int i = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  double* u = reinterpret_cast<double*>(smraw);
  double* v = u + (n + 1);
  double* minv = v + (n + 1);
  int* p = reinterpret_cast<int*>(minv + (n + 1));
  int* way = p + (n + 1);
  unsigned char* used = reinterpret_cast<unsigned char*>(way + (n + 1));
  float* sA = reinterpret_cast<float*>(used + ((n + 1 + 7) / 8) * 8);
  float* sB = sA + MATCH_KC * (MATCH_TILE + 1);
  float* scost = sB + MATCH_KC * (MATCH_TILE + 1);
  const size_t prob = (size_t)b * a.T + i;
  float* gcost = a.cost ? a.cost + prob * n * n : nullptr;
  const float* tgt = a.en + ((size_t)b * a.T + (i > 0 ? i - 1 : 0)) * n * C;
  const float* cur = a.en + prob * n * C;

  // ---- cost matrix: cost[t][k] = 1 - tgt[t] . cur[k]; 128 x 128 tiles, thread (ty, tx) owns rows ty + 16 r, cols tx + 16 c
  const int tx = tid & 15, ty = tid >> 4;
  for (int a0 = 0; a0 < n; a0 += MATCH_TILE)
    for (int k0 = 0; k0 < n; k0 += MATCH_TILE) {
      float acc[8][8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
      for (int c0 = 0; c0 < C; c0 += MATCH_KC) {
        __syncthreads();
        for (int e = tid; e < MATCH_TILE * MATCH_KC; e += MATCH_THREADS) {
          const int row = e / MATCH_KC, cc = e % MATCH_KC;
          const bool ch_ok = c0 + cc < C;
          sA[cc * (MATCH_TILE + 1) + row] = (ch_ok && a0 + row < n) ? tgt[(size_t)(a0 + row) * C + c0 + cc] : 0.f;
          sB[cc * (MATCH_TILE + 1) + row] = (ch_ok && k0 + row < n) ? cur[(size_t)(k0 + row) * C + c0 + cc] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int cc = 0; cc < MATCH_KC; ++cc) {
          float ra[8], rb[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) ra[r] = sA[cc * (MATCH_TILE + 1) + ty + 16 * r];
#pragma unroll
          for (int c = 0; c < 8; ++c) rb[c] = sB[cc * (MATCH_TILE + 1) + tx + 16 * c];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(ra[r], rb[c], acc[r][c]);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int row = a0 + ty + 16 * r, col = k0 + tx + 16 * c;
          if (row < n && col < n) {
            const float cv = 1.f - acc[r][c];
            if (a.cost_in_smem) scost[(size_t)row * n + col] = cv;
            if (gcost) gcost[(size_t)row * n + col] = cv;
          }
        }
    }
  __syncthreads();
  if (tid >= 32) return;

  // ---- assignment: shortest augmenting paths with dual potentials (exact optimum, like the reference's
  // scipy.optimize.linear_sum_assignment, minvis.py:38), one warp, lane owns columns j = lane + 1 + 32 m.
  // Rows = target queries (frame i-1), columns = current queries; 1-based, column 0 is the virtual root.
  const float* cbase = a.cost_in_smem ? scost : gcost;
  const double INF = 1e300;
  for (int j = lane; j <= n; j += 32) { u[j] = 0.0; v[j] = 0.0; p[j] = 0; way[j] = 0; }
  __syncwarp();
  for (int row = 1; row <= n; ++row) {
    for (int j = lane + 1; j <= n; j += 32) { minv[j] = INF; used[j] = 0; }
    if (lane == 0) { p[0] = row; used[0] = 0; }
    __syncwarp();
    int j0 = 0;
    while (true) {
      if (lane == 0) used[j0] = 1;
      __syncwarp();
      const int i0 = p[j0];
      const double ui0 = u[i0];
      const float* crow = cbase + (size_t)(i0 - 1) * n;
      double best = INF;
      int bj = 0x7fffffff;
      for (int j = lane + 1; j <= n; j += 32) {
        if (!used[j]) {
          float cf = crow[j - 1];
          if (!(cf == cf)) cf = 2.f;                 // NaN (zero-norm embedding): the largest cosine cost
          const double cur_c = (double)cf - ui0 - v[j];
          double mv = minv[j];
          if (cur_c < mv) { mv = cur_c; minv[j] = cur_c; way[j] = j0; }
          if (mv < best) { best = mv; bj = j; }
        }
      }
      // warp arg-min (smallest column among equal values) with three redux operations on an order-preserving integer
      // image of the double instead of five rounds of 64-bit shuffles + compares: this loop is the kernel's critical path
      {
        unsigned long long key = (unsigned long long)__double_as_longlong(best);
        key = (key >> 63) ? ~key : (key | 0x8000000000000000ull);
        const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
        const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
        const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
        bj = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? bj : 0x7fffffff);
        unsigned long long mk = ((unsigned long long)mh << 32) | ml;
        mk = (mk >> 63) ? (mk & 0x7fffffffffffffffull) : ~mk;
        best = __longlong_as_double((long long)mk);
      }
      // dual update; the same lane owns column j here and in the scan above (shuffles are no memory fence).  Column 0
      // (the root, always in the tree) carries the row being inserted.
      if (lane == 0) u[row] += best;
      for (int j = lane + 1; j <= n; j += 32) {
        if (used[j]) { u[p[j]] += best; v[j] -= best; }
        else minv[j] -= best;
      }
      __syncwarp();
      j0 = bj;
      if (p[j0] == 0) break;
    }
    if (lane == 0) {
      do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
    }
    __syncwarp();
  }
  int* pi = a.pi + prob * n;
  for (int j = lane + 1; j <= n; j += 32) pi[p[j] - 1] = j - 1;
}

// indices[b][i][j] = pi_i[indices[b][i-1][j]], indices[b][-1] = identity (minvis.py:49-62 in raw coordinates)
__global__ void match_compose_kernel(const int* __restrict__ pi, long long* __restrict__ indices, int T, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (j >= n) return;
  int idx = j;
  for (int i = 0; i < T; ++i) {
    idx = pi[((size_t)b * T + i) * n + idx];
    indices[((size_t)b * T + i) * n + j] = idx;
  }
}

// out[b][t][q][:] = in[b][t][idx[b][t][q]][:] with explicit element strides for (b, t, q) and `inner` contiguous floats:
// batch_index (openvis/utils/index.py:4-11) as used by batch_video_match_via_embeds (minvis.py:57) and
// BriVIS.reset_image_output_order (brivis.py:231-240).
__global__ void __launch_bounds__(128)
reorder_queries_kernel(const float* __restrict__ in, const long long* __restrict__ idx, float* __restrict__ out, int T, int n,
                       long long inner, long long sb, long long st, long long sq) {
  const int q = blockIdx.x, t = blockIdx.y, b = blockIdx.z;
  const long long src_q = idx[((size_t)b * T + t) * n + q];
  const float* s = in + b * sb + t * st + src_q * sq;
  float* d = out + b * sb + t * st + q * sq;
  if (((inner | sb | st | sq) & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    for (long long e = threadIdx.x; e < inner / 4; e += blockDim.x)
      reinterpret_cast<float4*>(d)[e] = reinterpret_cast<const float4*>(s)[e];
  } else {
    for (long long e = threadIdx.x; e < inner; e += blockDim.x) d[e] = s[e];
  }
}

}  // namespace ovis
