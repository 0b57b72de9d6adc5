// C-ABI implementation (include/openvis_b200.h): argument checks, TMA tensor-map encoding, launches.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/openvis_b200.h"
#include "gemm_tn.cuh"
#include "prep.cuh"
#include "prep_tma.cuh"
#include "xattn.cuh"
#include "xattn_tc.cuh"
#include "xattn_tc2.cuh"
#include "xattn_tc3.cuh"
#include "chain.cuh"
#include "crop.cuh"
#include <vector>
#include "san_attn.cuh"
#include "san_attn_tc.cuh"
#include "postproc.cuh"
#include "msda.cuh"
#include "temporal.cuh"
#include "pixdec.cuh"

using namespace ovis;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}
#define CHECK_ARG(cond, msg)                                      \
  do {                                                            \
    if (!(cond)) return fail(OVIS_ERR_ARG, "%s: " msg, __func__); \
  } while (0)

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: launch failed: %s", what, cudaGetErrorString(e));
    return OVIS_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OVIS_OK;
}

struct DeviceInfo {
  int ok = -1;      // -1 unknown, 0 good, else error code
  int sms = 0;
};
DeviceInfo g_dev[64];
std::mutex g_mu;

int device_info(int* sms) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return fail(OVIS_ERR_ARCH, "%s: no CUDA device", "ovis");
  }
  if (dev < 0 || dev >= 64) return fail(OVIS_ERR_ARCH, "%s: device index out of range", "ovis");
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceInfo& d = g_dev[dev];
  if (d.ok < 0) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
      cudaGetLastError();
      return fail(OVIS_ERR_ARCH, "%s: cannot query device", "ovis");
    }
    d.sms = p.multiProcessorCount;
    d.ok = (p.major == 10) ? 0 : OVIS_ERR_ARCH;
  }
  if (d.ok != 0) return fail(OVIS_ERR_ARCH, "%s: kernels are built for sm_100a (B200) only; no fallback path", "ovis");
  if (sms) {
    // OVIS_SM_BUDGET=n: persistent kernels use at most n SMs, leaving the rest to concurrently running clips' small
    // (latency-bound) kernels on other streams
    static const int budget = getenv("OVIS_SM_BUDGET") ? atoi(getenv("OVIS_SM_BUDGET")) : 0;
    *sms = (budget > 0 && budget < d.sms) ? budget : d.sms;
  }
  return OVIS_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int get_encoder() {
  if (g_encode) return OVIS_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return fail(OVIS_ERR_CUDA, "%s: cuTensorMapEncodeTiled not available from the driver", "ovis");
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return OVIS_OK;
}

// fp16 row-major [rows][cols] (row stride ld elements) -> box {64 cols, box_rows}, 128B swizzle, zero OOB fill
int make_map_f16(CUtensorMap* m, const void* base, unsigned long long rows, unsigned long long cols, unsigned long long ld,
                 unsigned box_rows) {
  int rc = get_encoder();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) return fail(OVIS_ERR_ARG, "%s: operand must be 16-byte aligned with a 16-byte-multiple row pitch", "tensor map");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OVIS_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%lld)", "tensor map", (long long)r);
  return OVIS_OK;
}

// fp16 [rows][cols] -> load box {32 columns = 64 bytes, box_rows rows}, 64B swizzle (per-head V boxes of xattn_tc3)
int make_map_f16_sw64(CUtensorMap* m, const void* base, unsigned long long rows, unsigned long long cols, unsigned long long ld,
                      unsigned box_rows) {
  int rc = get_encoder();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) return fail(OVIS_ERR_ARG, "%s: operand must be 16-byte aligned with a 16-byte-multiple row pitch", "tensor map");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OVIS_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%lld)", "tensor map (64B swizzle)", (long long)r);
  return OVIS_OK;
}

// output [rows][cols] (fp16 / fp32, row pitch ld elements) -> store box {128 bytes, 32 rows}, 128B swizzle
int make_store_map(CUtensorMap* m, const void* base, unsigned long long rows, unsigned long long cols, unsigned long long ld,
                   int f32) {
  int rc = get_encoder();
  if (rc) return rc;
  const unsigned es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * es};
  cuuint32_t box[2] = {128 / es, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OVIS_ERR_CUDA, "%s: cuTensorMapEncodeTiled (store) failed (%lld)", "tensor map", (long long)r);
  return OVIS_OK;
}

// generic 4-D tiled map (dims / box innermost first; strides in bytes for dims 1..3)
int make_map_4d(CUtensorMap* m, const void* base, int f32, const unsigned long long dims[4], const unsigned long long strides[3],
                const unsigned box[4], CUtensorMapSwizzle sw) {
  int rc = get_encoder();
  if (rc) return rc;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides[0], strides[1], strides[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), d,
                        st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OVIS_ERR_CUDA, "%s: cuTensorMapEncodeTiled (4-D) failed (%lld)", "tensor map", (long long)r);
  return OVIS_OK;
}

// Launch helper of the kernels that execute griddepcontrol.wait before their first global access (ptx.cuh: pdl_begin).
// OVIS_PDL=1 adds the programmatic-stream-serialization attribute.  Off by default: measured on the five bench workloads
// it changes nothing (profiles/experiments/chain_r2.md) -- the dependent query-side kernels are bound by their own
// load -> MMA -> epilogue latency chain (8 us for one 128 x 256 x 256 tile), not by the launch gap.
static bool pdl_enabled() {
  static const bool on = getenv("OVIS_PDL") && !strcmp(getenv("OVIS_PDL"), "1");
  return on;
}
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);     // (errors surface in check_launch)
}

void init_args(GemmArgs& a) {
  memset(&a, 0, sizeof(a));
  a.num_groups = 1;
  a.a_k_mod = 1;
  a.a_row_div = 1;
  a.b_k_mod = 1;
  a.kps = 1;
  a.a_prefetch = 0;
  a.b_row_div = 1;
  a.scale = 1.f;
  a.pe_period = 1;
}

template <int BN>
int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const GemmOutMaps& om,
                   const GemmArgs& a, int sms, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "gemm: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return OVIS_ERR_CUDA;
    }
    attr_done[dev] = true;
  }
  const int m_tiles = (a.rows_per_group + 127) / 128;
  const int n_tiles = (a.N + BN - 1) / BN;
  const long long total = (long long)a.num_groups * m_tiles * n_tiles;
  if (total <= 0) return OVIS_OK;
  const int grid = (int)(total < sms ? total : sms);
  launch_k(gemm_tn_kernel<BN>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, ta, ta2, tb, om, a);
  return check_launch("gemm_tn_kernel");
}

template <int BN>
int launch_gemm_bs(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const GemmOutMaps& om,
                   const GemmArgs& a, int sms, cudaStream_t st) {
  using Cfg = GemmBsCfg<BN>;
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_bs_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "gemm(bs): cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return OVIS_ERR_CUDA;
    }
    attr_done[dev] = true;
  }
  const int m_tiles = (a.rows_per_group + 127) / 128;
  const int n_tiles = (a.N + BN - 1) / BN;
  const long long cols = (long long)a.num_groups * n_tiles;
  if (cols <= 0 || m_tiles <= 0) return OVIS_OK;
  // ncols x R CTAs: as many columns side by side as fit, R CTAs walking the m-tiles of each
  const int ncols = (int)(cols < sms ? cols : sms);
  int R = sms / ncols;
  if (R > m_tiles) R = m_tiles;
  if (R < 1) R = 1;
  gemm_tn_bs_kernel<BN><<<ncols * R, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ta, ta2, tb, om, a);
  return check_launch("gemm_tn_bs_kernel");
}

// A: [a_rows][a_cols] fp16 pitch lda; B: [b_rows][K] fp16 pitch ldb.
int launch_gemm(const void* A, long long a_rows, long long a_cols, long long lda, const void* B, long long b_rows,
                long long ldb, const GemmArgs& a_in, int bn, cudaStream_t st, const void* A2 = nullptr) {
  GemmArgs a = a_in;
  int sms = 0;
  int rc = device_info(&sms);
  if (rc) return rc;
  if (a.K <= 0 || a.K % 64) return fail(OVIS_ERR_ARG, "%s: K must be a positive multiple of 64", "gemm");
  if ((a.N + bn - 1) / bn > GEMM_MAX_NTILES) return fail(OVIS_ERR_ARG, "%s: too many column tiles", "gemm");
  CUtensorMap ta, ta2, tb;
  rc = make_map_f16(&ta, A, a_rows, a_cols, lda, 128);
  if (rc) return rc;
  rc = make_map_f16(&ta2, A2 ? A2 : A, a_rows, a_cols, lda, 128);
  if (rc) return rc;
  rc = make_map_f16(&tb, B, b_rows, (unsigned long long)a.K * a.b_k_mod, ldb, bn);
  if (rc) return rc;
  // TMA-store epilogue for plain stores: one group, 16-byte aligned rows, few enough n-tiles for the map array
  static GemmOutMaps om;     // (contents only read when a.tma_store is set; copied into the launch by value)
  GemmOutMaps om_local;
  const GemmOutMaps* omp = &om;
  {
    static const bool no_tma_store = getenv("OVIS_GEMM_NO_TMA_STORE") != nullptr;    // A/B testing only
    const int nt = (a.N + bn - 1) / bn;
    const int es = a.out_f32 ? 4 : 2;
    bool ok = a.epi == EPI_STORE && (a.num_groups == 1 || (a.o_group_stride > 0 && a.o_group_stride % 32 == 0)) &&
              nt <= GEMM_MAX_OUT_MAPS && (((long long)a.ldo * es) % 16) == 0 && !no_tma_store &&
              !(a.resid_st && a.out_f32);   // fp32 residual: the lean epilogue adds it in its row-coalesced store loop
    for (int t = 0; ok && t < nt; ++t) ok = (reinterpret_cast<uintptr_t>(a.out[t]) & 15) == 0;
    if (ok) {
      for (int t = 0; t < nt; ++t) {
        const int ncols = a.N - t * bn < bn ? a.N - t * bn : bn;
        const unsigned long long out_rows = a.num_groups == 1 ? (unsigned long long)a.rows_per_group
                                                               : (unsigned long long)a.num_groups * a.o_group_stride;
        rc = make_store_map(&om_local.m[t], a.out[t], out_rows, (unsigned long long)ncols, (unsigned long long)a.ldo, a.out_f32);
        if (rc) return rc;
      }
      a.tma_store = 1;
      omp = &om_local;
    }
    // transposed fp32 store: one 3-D map {row in group, column, group}, box {32 rows (128 B), 32 columns, 1}
    if (a.epi == EPI_STORE_T && !no_tma_store && (reinterpret_cast<uintptr_t>(a.out_t) & 15) == 0 && (a.ldt % 4) == 0 &&
        (a.num_groups == 1 || (a.t_group_stride % 4) == 0)) {
      rc = get_encoder();
      if (rc) return rc;
      cuuint64_t dims[3] = {(cuuint64_t)a.rows_per_group, (cuuint64_t)a.N, (cuuint64_t)a.num_groups};
      cuuint64_t strides[2] = {(cuuint64_t)a.ldt * 4,
                               a.num_groups == 1 ? (cuuint64_t)a.ldt * 4 * (cuuint64_t)a.N : (cuuint64_t)a.t_group_stride * 4};
      cuuint32_t box[3] = {32, 32, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = g_encode(&om_local.m[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a.out_t, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(OVIS_ERR_CUDA, "%s: cuTensorMapEncodeTiled (transposed store) failed (%lld)", "tensor map", (long long)r);
      a.tma_store = 1;
      omp = &om_local;
    }
  }
  // HBM-heavy shapes (K <= 256, many row tiles): B-stationary kernel; small / long-K shapes: streaming kernel
  const long long m_tiles_total = (long long)a.num_groups * ((a.rows_per_group + 127) / 128);
  static const bool no_bs = getenv("OVIS_GEMM_NO_BS") != nullptr;    // A/B testing only
  {
    // OVIS_GEMM_KPS (A/B): k-blocks per ring barrier of the B-stationary kernel; default = all k-blocks of a row tile for
    // the 128-wide kernel (ring of 8 slots = 2 tiles; mask_bits 69 -> 63.5 us), 1 for the 256-wide one (ring of 4 slots:
    // grouping halves the look-ahead there, kv_proj 476 -> 487 us)
    static const int kps_env = getenv("OVIS_GEMM_KPS") ? atoi(getenv("OVIS_GEMM_KPS")) : 0;
    const int kblocks = a.K / 64;
    int kps = kps_env > 0 ? kps_env : (bn == 256 ? 1 : 4);
    while (kps > 1 && (kblocks % kps != 0 || (bn == 256 ? 4 : 8) % kps != 0)) kps >>= 1;
    a.kps = kps < 1 ? 1 : kps;
    static const int pf_env = getenv("OVIS_GEMM_PF") ? atoi(getenv("OVIS_GEMM_PF")) : -1;    // A/B testing
    a.a_prefetch = pf_env >= 0 ? pf_env : 0;    // (measured: no gain for kv_proj, -12 % for mask_logits: profiles/experiments/gemm_bs_r2.md)
  }
  if (a.K <= 256 && m_tiles_total >= 64 && !no_bs)
    return bn == 256 ? launch_gemm_bs<256>(ta, ta2, tb, *omp, a, sms, st) : launch_gemm_bs<128>(ta, ta2, tb, *omp, a, sms, st);
  return bn == 256 ? launch_gemm_bn<256>(ta, ta2, tb, *omp, a, sms, st) : launch_gemm_bn<128>(ta, ta2, tb, *omp, a, sms, st);
}

}  // namespace

extern "C" {

int ovis_version(void) { return 100; }
const char* ovis_last_error(void) { return g_err; }
int ovis_device_check(void) { return device_info(nullptr); }
long long ovis_launch_count(void) { return g_launches.load(); }
void ovis_add_launch_count(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ovis_nchw_to_tokens_f16(const float* in, void* out, void* out_pos, const float* pos, const float* pos_t, int B, int C,
                            int N, void* stream) {
  CHECK_ARG(in && out && B > 0 && C > 0 && N > 0 && C % 2 == 0, "bad arguments");
  CHECK_ARG((out_pos == nullptr) == (pos == nullptr), "out_pos and pos go together");
  int rc = device_info(nullptr);
  if (rc) return rc;
  dim3 grid((N + 63) / 64, (C + 63) / 64, B);
  nchw_to_tokens_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (__half*)out, (__half*)out_pos, pos, pos_t, C, N);
  return check_launch("nchw_to_tokens_f16_kernel");
}

int ovis_nchw_to_tokens_hw_f16(const float* in, void* out, void* out_pos, const float* pos_cn, const float* pos_t, int B, int C,
                               int h, int w, void* stream) {
  CHECK_ARG(in && out && B > 0 && C > 0 && h > 0 && w > 0, "bad arguments");
  CHECK_ARG(C % 32 == 0 && w % 4 == 0, "needs C % 32 == 0 and w % 4 == 0 (16-byte row pitch)");
  CHECK_ARG((out_pos == nullptr) == (pos_cn == nullptr), "out_pos and pos_cn go together");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_pos)) & 15) == 0,
            "pointers must be 16-byte aligned");
  int sms = 0;
  int rc = device_info(&sms);
  if (rc) return rc;
  TokTmaMaps maps;
  const unsigned long long uW = w, uH = h, uC = C, uB = B;
  {
    const unsigned long long d[4] = {uW, uH, uC, uB}, st[3] = {uW * 4, uH * uW * 4, uC * uH * uW * 4};
    const unsigned bx[4] = {32, 8, 32, 1};
    rc = make_map_4d(&maps.in, in, 1, d, st, bx, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    const unsigned long long d[4] = {uC, uW, uH, uB}, st[3] = {uC * 2, uW * uC * 2, uH * uW * uC * 2};
    const unsigned bx[4] = {32, 32, 8, 1};
    rc = make_map_4d(&maps.xt, out, 0, d, st, bx, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = make_map_4d(&maps.xp, out_pos ? out_pos : out, 0, d, st, bx, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tokens_prep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_SMEM);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "nchw_to_tokens: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return OVIS_ERR_CUDA;
    }
    attr_done[dev] = true;
  }
  // work item = (spatial tile, run of frames): enough runs for >= 4 items per SM, runs of at most TT_MAX_RUN frames
  const long long spatial = (long long)((h + 7) / 8) * ((w + 31) / 32) * (C / 32);
  long long runs = (4ll * sms + spatial - 1) / spatial;
  if (runs < (B + TT_MAX_RUN - 1) / TT_MAX_RUN) runs = (B + TT_MAX_RUN - 1) / TT_MAX_RUN;
  if (runs > B) runs = B;
  const int run_len = (int)((B + runs - 1) / runs);
  runs = (B + run_len - 1) / run_len;
  const long long items = spatial * runs;
  const int grid = (int)(items < sms ? items : sms);
  tokens_prep_tma_kernel<<<grid, 256, TT_SMEM, (cudaStream_t)stream>>>(maps, pos_cn, pos_t, B, C, h, w, (int)runs, run_len);
  return check_launch("tokens_prep_tma_kernel");
}

int ovis_maskfeat_prep(const float* F, void* ft, void* g0, void* g1, void* g2, int B, int C, int H, int W, void* stream) {
  CHECK_ARG(F && ft && g0 && g1 && g2, "null pointer");
  CHECK_ARG(B > 0 && C % 32 == 0 && H % 8 == 0 && W % 8 == 0 && H > 0 && W > 0, "needs C % 32 == 0, H % 8 == 0, W % 8 == 0");
  int rc = device_info(nullptr);
  if (rc) return rc;
  static const bool no_tma = getenv("OVIS_PREP_NO_TMA") != nullptr;     // A/B testing only
  const bool aligned = ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(ft) | reinterpret_cast<uintptr_t>(g0) |
                         reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(g2)) & 15) == 0;
  if (!no_tma && aligned) {
    int sms = 0;
    rc = device_info(&sms);
    if (rc) return rc;
    PrepTmaMaps maps;
    const unsigned long long uW = W, uH = H, uC = C, uB = B;
    {
      const unsigned long long d[4] = {uW, uH, uC, uB}, st[3] = {uW * 4, uH * uW * 4, uC * uH * uW * 4};
      const unsigned bx[4] = {32, 8, 32, 1};
      rc = make_map_4d(&maps.in, F, 1, d, st, bx, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
    void* outs[4] = {ft, g2, g1, g0};
    CUtensorMap* om[4] = {&maps.ft, &maps.g2, &maps.g1, &maps.g0};
    for (int l = 0; l < 4; ++l) {
      const unsigned s_ = 1u << l;                       // block size 1, 2, 4, 8
      const unsigned long long w = uW / s_, h = uH / s_;
      const unsigned long long d[4] = {uC, w, h, uB}, st[3] = {uC * 2, w * uC * 2, h * w * uC * 2};
      const unsigned bx[4] = {32, 32 / s_, 8 / s_, 1};
      rc = make_map_4d(om[l], outs[l], 0, d, st, bx, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(maskfeat_prep_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "maskfeat_prep: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
        return OVIS_ERR_CUDA;
      }
      attr_done[dev] = true;
    }
    const long long total = (long long)B * (H / 8) * ((W + 31) / 32) * (C / 32);
    const int grid = (int)(total < sms ? total : sms);
    maskfeat_prep_tma_kernel<<<grid, 256, PT_SMEM, (cudaStream_t)stream>>>(maps, B, C, H, W);
    return check_launch("maskfeat_prep_tma_kernel");
  }
  dim3 grid(((W + 31) / 32) * (H / 8), C / 32, B);
  maskfeat_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, (__half*)ft, (__half*)g0, (__half*)g1, (__half*)g2, C, H, W);
  return check_launch("maskfeat_prep_kernel");
}

int ovis_cast_f16(const float* in, void* out, long long n, void* stream) {
  CHECK_ARG(in && out && n >= 0, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  if (n == 0) return OVIS_OK;
  cast_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (__half*)out, n);
  return check_launch("cast_f16_kernel");
}

int ovis_init_queries(const float* qf, const float* qe, const float* g, const float* b, float* z32, void* z16, void* ze16,
                      float* d32, void* d16, int Q, int rows, void* stream) {
  CHECK_ARG(qf && qe && g && b && z32 && z16 && ze16 && d32 && d16 && Q > 0 && rows > 0, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  launch_k(init_queries_kernel, dim3((rows + 7) / 8), dim3(256), 0, (cudaStream_t)stream, qf, qe, g, b, z32, (__half*)z16, (__half*)ze16, d32,
                                                                        (__half*)d16, Q, rows);
  return check_launch("init_queries_kernel");
}

int ovis_rownorm(const float* in, const float* g, const float* b, float* out32, void* out16, int rows, int D, int mode,
                 void* stream) {
  CHECK_ARG(in && rows > 0 && D > 0 && (out32 || out16), "bad arguments");
  CHECK_ARG(!(mode & 1) || (g && b), "LayerNorm needs weight and bias");
  int rc = device_info(nullptr);
  if (rc) return rc;
  CHECK_ARG(!(mode & 4), "mode bit 2 (row sums of squares) needs ovis_rowstats");
  rownorm_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(in, g, b, out32, (__half*)out16, rows, D, mode, nullptr, nullptr, 0, 0);
  return check_launch("rownorm_kernel");
}

int ovis_rowstats(const float* in, const float* g, const float* b, void* out16, float* ss, float* ss_zero, int rows, int D,
                  int layer_norm, int group_rows, int group_stride, void* stream) {
  CHECK_ARG(in && rows > 0 && D > 0 && out16 && (ss || ss_zero) && group_rows >= 0 && (group_rows == 0 || group_stride >= group_rows),
            "bad arguments");
  CHECK_ARG(!layer_norm || (g && b), "LayerNorm needs weight and bias");
  int rc = device_info(nullptr);
  if (rc) return rc;
  rownorm_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(in, g, b, nullptr, (__half*)out16, rows, D,
                                                                    (layer_norm ? 1 : 0) | (ss ? 4 : 0), ss, ss_zero, group_rows,
                                                                    group_stride);
  return check_launch("rownorm_kernel");
}

int ovis_linear_rowscale_f16(const void* x, long long rows, int K, int ldx, const void* w, int N, const float* bias, float scale,
                             const float* row_ss_in, float* row_ss_out, void* out, int ldo, int out_f32, void* stream) {
  CHECK_ARG(x && w && out && rows > 0 && N > 0 && ldx >= K && ldo >= N && (row_ss_in || row_ss_out), "bad arguments");
  CHECK_ARG(rows < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = (int)rows;
  a.a_group_stride = (int)rows;
  a.N = N;
  a.K = K;
  a.epi = EPI_STORE;
  const int bn = (N <= 128) ? 128 : 256;
  const int nt = (N + bn - 1) / bn;
  CHECK_ARG(nt <= GEMM_MAX_NTILES, "N too large");
  for (int t = 0; t < nt; ++t) {
    a.out[t] = out_f32 ? (void*)((float*)out + (long long)t * bn) : (void*)((__half*)out + (long long)t * bn);
    a.bias[t] = bias ? bias + (long long)t * bn : nullptr;
  }
  a.ldo = ldo;
  a.out_f32 = out_f32;
  a.scale = scale;
  a.row_ss_in = row_ss_in;
  a.row_ss_out = row_ss_out;
  return launch_gemm(x, rows, K, ldx, w, N, K, a, bn, (cudaStream_t)stream);
}

int ovis_linear_f16(const void* x, long long rows, int K, int ldx, const void* w, int N, const float* bias, float scale,
                    int relu, void* out, int ldo, int out_f32, void* stream) {
  CHECK_ARG(x && w && out && rows > 0 && N > 0 && ldx >= K && ldo >= N, "bad arguments");
  CHECK_ARG(rows < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = (int)rows;
  a.a_group_stride = (int)rows;
  a.N = N;
  a.K = K;
  a.epi = EPI_STORE;
  const int bn = (N <= 128) ? 128 : 256;
  const int nt = (N + bn - 1) / bn;
  CHECK_ARG(nt <= GEMM_MAX_NTILES, "N too large");
  for (int t = 0; t < nt; ++t) {
    a.out[t] = out_f32 ? (void*)((float*)out + (long long)t * bn) : (void*)((__half*)out + (long long)t * bn);
    a.bias[t] = bias ? bias + (long long)t * bn : nullptr;
  }
  a.ldo = ldo;
  a.out_f32 = out_f32;
  a.relu = relu;
  a.scale = scale;
  return launch_gemm(x, rows, K, ldx, w, N, K, a, bn, (cudaStream_t)stream);
}

int ovis_linear_act_f16(const void* x, long long rows, int K, int ldx, const void* w, int N, const float* bias, float scale,
                        int act, const float* resid, void* out, int ldo, int out_f32, void* stream) {
  CHECK_ARG(x && w && out && rows > 0 && N > 0 && ldx >= K && ldo >= N && act >= 0 && act <= 2, "bad arguments");
  CHECK_ARG(rows < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = (int)rows;
  a.a_group_stride = (int)rows;
  a.N = N;
  a.K = K;
  a.epi = EPI_STORE;
  const int bn = (N <= 128) ? 128 : 256;
  const int nt = (N + bn - 1) / bn;
  CHECK_ARG(nt <= GEMM_MAX_NTILES, "N too large");
  for (int t = 0; t < nt; ++t) {
    a.out[t] = out_f32 ? (void*)((float*)out + (long long)t * bn) : (void*)((__half*)out + (long long)t * bn);
    a.bias[t] = bias ? bias + (long long)t * bn : nullptr;
  }
  a.ldo = ldo;
  a.out_f32 = out_f32;
  a.relu = act;
  a.scale = scale;
  a.resid_st = resid;
  return launch_gemm(x, rows, K, ldx, w, N, K, a, bn, (cudaStream_t)stream);
}

int ovis_linear_ln_f16(const void* x, long long rows, int K, const void* w, const float* bias, const float* resid,
                       const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b, const float* pe,
                       int pe_period, float* y32, void* y16, void* ype16, float* d32, void* d16, float* split_ws,
                       long long split_ws_floats, void* stream) {
  CHECK_ARG(x && w && bias && resid && ln1_g && ln1_b && rows > 0, "bad arguments");
  CHECK_ARG((ln2_g == nullptr) == (ln2_b == nullptr), "second LayerNorm needs both weight and bias");
  CHECK_ARG(!ype16 || (pe && pe_period > 0), "ype16 needs pe");
  CHECK_ARG(rows < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  // Up to 16384 rows (the Video decoders' 100..400 queries, the Frame decoders' frames x queries): one CTA per 128-row
  // tile with the whole K loop and a row-serial LayerNorm epilogue leaves most of the GPU idle (37-57 us per call), so
  // the product is split over K slices of 256 and two 128-column tiles (fp32 partials in split_ws) and a row-parallel
  // kernel finishes bias + residual + LayerNorm(s).  Beyond that the fused epilogue has enough tiles to fill the GPU.
  const long long rows_pad = ((rows + 127) / 128) * 128;
  const int S = K / 256;
  static const bool no_split = getenv("OVIS_LN_NO_SPLIT") != nullptr;     // A/B testing only
  // Many rows (the pixel decoder's encoder: frames x 19 320 positions): the row-serial epilogue's thread-per-row stores are
  // uncoalesced, so with a workspace the product is stored plainly (TMA, fp32) and the same row-parallel kernel finishes.
  const bool many = !no_split && split_ws && rows > 16384 && split_ws_floats >= rows_pad * 256;
  if (many) {
    a.rows_per_group = (int)rows;
    a.a_group_stride = (int)rows;
    a.N = 256;
    a.K = K;
    a.epi = EPI_STORE;
    a.out[0] = split_ws;
    a.ldo = 256;
    a.out_f32 = 1;
    int rc = launch_gemm(x, rows, K, K, w, 256, K, a, 256, (cudaStream_t)stream);
    if (rc) return rc;
  }
  if (many || (!no_split && split_ws && K % 256 == 0 && rows <= 16384 && split_ws_floats >= (long long)S * rows_pad * 256)) {
    if (!many) {
      a.rows_per_group = (int)rows;
      a.num_groups = S;
      a.a_group_stride = 0;
      a.a_k_mod = S;
      a.a_k_offset_stride = 256;
      a.b_k_mod = S;
      a.b_k_offset_stride = 256;
      a.o_group_stride = (int)rows_pad;
      a.N = 256;
      a.K = 256;
      a.epi = EPI_STORE;
      a.out[0] = split_ws;
      a.out[1] = split_ws + 128;
      a.ldo = 256;
      a.out_f32 = 1;
      int rc = launch_gemm(x, rows, K, K, w, 256, K, a, 128, (cudaStream_t)stream);
      if (rc) return rc;
    }
    LnReduceArgs r;
    r.part = split_ws; r.S = many ? 1 : S; r.part_stride = rows_pad;
    r.bias = bias; r.resid = resid;
    r.ln1_g = ln1_g; r.ln1_b = ln1_b; r.ln2_g = ln2_g; r.ln2_b = ln2_b;
    r.pe = pe; r.pe_period = pe_period > 0 ? pe_period : 1;
    r.y32 = y32; r.y16 = (__half*)y16; r.ype16 = (__half*)ype16; r.d32 = d32; r.d16 = (__half*)d16;
    r.rows = (int)rows;
    launch_k(ln_reduce_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, r);
    return check_launch("ln_reduce_kernel");
  }
  a.rows_per_group = (int)rows;
  a.a_group_stride = (int)rows;
  a.N = 256;
  a.K = K;
  a.epi = EPI_LN;
  a.bias[0] = bias;
  a.resid = resid;
  a.ln1_g = ln1_g; a.ln1_b = ln1_b; a.ln2_g = ln2_g; a.ln2_b = ln2_b;
  a.pe = pe; a.pe_period = pe_period > 0 ? pe_period : 1;
  a.y32 = y32; a.y16 = (__half*)y16; a.ype16 = (__half*)ype16; a.d32 = d32; a.d16 = (__half*)d16;
  return launch_gemm(x, rows, K, K, w, 256, K, a, 256, (cudaStream_t)stream);
}

int ovis_kv_proj_f16(const void* xk, const void* xv, long long rows, const void* w, int n_tiles, void* const* out,
                     const float* const* bias, void* stream) {
  CHECK_ARG(xk && xv && w && out && rows > 0 && n_tiles > 0 && n_tiles <= GEMM_MAX_NTILES, "bad arguments");
  CHECK_ARG(rows < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = (int)rows;
  a.a_group_stride = (int)rows;
  a.N = n_tiles * 256;
  a.K = 256;
  a.epi = EPI_STORE;
  a.a_alt = 1;
  a.ldo = 256;
  // OVIS_KV_BN=128: two 128-wide column tiles per output (deeper A ring, smaller accumulators); default 256
  static const int kv_bn = getenv("OVIS_KV_BN") ? atoi(getenv("OVIS_KV_BN")) : 256;
  const int bn = (kv_bn == 128 && 2 * n_tiles <= GEMM_MAX_NTILES && 2 * n_tiles <= GEMM_MAX_OUT_MAPS) ? 128 : 256;
  const int split = 256 / bn;
  a.a_alt_shift = split == 2 ? 1 : 0;
  for (int t = 0; t < n_tiles; ++t) {
    CHECK_ARG(out[t] != nullptr, "null output tile");
    for (int h = 0; h < split; ++h) {
      a.out[t * split + h] = (__half*)out[t] + h * bn;
      a.bias[t * split + h] = (bias && bias[t]) ? bias[t] + h * bn : nullptr;
    }
  }
  return launch_gemm(xk, rows, 256, 256, w, (long long)n_tiles * 256, 256, a, bn, (cudaStream_t)stream, xv);
}

int ovis_mask_bits(const void* gt, int groups, int rows_per_group, const void* me, int Q, unsigned int* bits,
                   unsigned char* flags, int q_stride, void* stream) {
  CHECK_ARG(gt && me && bits && flags && groups > 0 && rows_per_group > 0 && Q > 0 && Q <= 256 && q_stride >= Q, "bad arguments");
  CHECK_ARG((long long)groups * rows_per_group < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = rows_per_group;
  a.num_groups = groups;
  a.a_group_stride = rows_per_group;
  a.b_group_stride = Q;
  a.N = Q;
  a.K = 256;
  a.epi = EPI_SIGNBITS;
  a.bits = bits;
  a.flags = flags;
  a.words_per_group = (rows_per_group + 31) / 32;
  a.q_stride = q_stride;
  return launch_gemm(gt, (long long)groups * rows_per_group, 256, 256, me, (long long)groups * Q, 256, a, Q <= 128 ? 128 : 256,
                     (cudaStream_t)stream);
}

int ovis_mask_bits_t(const void* gt, int groups, int rows_per_group, const void* me, int Q, unsigned int* bits_t,
                     unsigned int* blockand, unsigned char* flags, int q_stride, void* stream) {
  CHECK_ARG(gt && me && bits_t && blockand && flags && groups > 0 && rows_per_group > 0 && Q > 0 && Q <= 256 && q_stride >= Q,
            "bad arguments");
  CHECK_ARG((long long)groups * rows_per_group < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = rows_per_group;
  a.num_groups = groups;
  a.a_group_stride = rows_per_group;
  a.b_group_stride = Q;
  a.N = Q;
  a.K = 256;
  a.epi = EPI_SIGNBITS_T;
  a.bits_t = bits_t;
  a.blockand = blockand;
  a.flags = flags;
  a.words_per_group = (rows_per_group + 31) / 32;
  a.q_stride = q_stride;
  a.qw = 4 * ((Q + 127) / 128);
  return launch_gemm(gt, (long long)groups * rows_per_group, 256, 256, me, (long long)groups * Q, 256, a, Q <= 128 ? 128 : 256,
                     (cudaStream_t)stream);
}

int ovis_mask_logits(const void* ft, int groups, int rows_per_group, const void* me, int me_group_stride, int Q,
                     const float* bias, float* out, long long t_group_stride, long long ldt,
                     unsigned char* posflags, int rows_per_frame, void* stream) {
  CHECK_ARG(ft && me && out && groups > 0 && rows_per_group > 0 && Q > 0 && me_group_stride >= 0, "bad arguments");
  CHECK_ARG(!posflags || (rows_per_frame > 0 && rows_per_frame % 32 == 0 && rows_per_group % rows_per_frame == 0),
            "posflags needs rows_per_frame % 32 == 0 dividing rows_per_group");
  CHECK_ARG((long long)groups * rows_per_group < (1ll << 31), "too many rows");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = rows_per_group;
  a.num_groups = groups;
  a.a_group_stride = rows_per_group;
  a.b_group_stride = me_group_stride;
  a.N = Q;
  a.K = 256;
  a.epi = EPI_STORE_T;
  const int bn = Q <= 128 ? 128 : 256;
  const int nt = (Q + bn - 1) / bn;
  CHECK_ARG(nt <= GEMM_MAX_NTILES, "too many output channels");
  for (int t = 0; t < nt; ++t) a.bias[t] = bias ? bias + (long long)t * bn : nullptr;
  a.out_t = out;
  a.t_group_stride = t_group_stride;
  a.ldt = ldt;
  a.posflags = posflags;
  a.rows_per_frame = rows_per_frame > 0 ? rows_per_frame : 32;
  a.q_stride = Q;
  const long long b_rows = me_group_stride ? (long long)(groups - 1) * me_group_stride + Q : Q;
  return launch_gemm(ft, (long long)groups * rows_per_group, 256, 256, me, b_rows, 256, a, bn, (cudaStream_t)stream);
}

int ovis_san_bias_logits(const void* af, int B, int P, int heads, const void* ae, int Q, float* out, void* stream) {
  CHECK_ARG(af && ae && out && B > 0 && P > 0 && heads > 0 && Q > 0 && Q <= 256, "bad arguments");
  GemmArgs a;
  init_args(a);
  a.rows_per_group = P;
  a.num_groups = B * heads;
  a.a_group_stride = P;
  a.a_row_div = heads;
  a.a_k_mod = heads;
  a.a_k_offset_stride = 256;
  a.b_group_stride = Q;
  a.b_row_div = heads;
  a.N = Q;
  a.K = 256;
  a.epi = EPI_STORE_T;
  a.out_t = out;
  a.t_group_stride = (long long)Q * P;
  a.ldt = P;
  return launch_gemm(af, (long long)B * P, (long long)heads * 256, (long long)heads * 256, ae, (long long)B * Q, 256, a,
                     Q <= 128 ? 128 : 256, (cudaStream_t)stream);
}

// which masked cross-attention kernel runs: 2 = head-pair CTAs, four softmax warpgroups (default);
// OVIS_XATTN=1 -> first tcgen05 layout, OVIS_XATTN=0 -> round-1 mma.sync kernel (A/B testing only)
static int xattn_variant() {
  static const int v = getenv("OVIS_XATTN") ? atoi(getenv("OVIS_XATTN")) : 2;
  return v;
}

int ovis_xattn_plan(int G, int Q, int keys, int* splits, int* q_pad, long long* o_floats, long long* ml_floats) {
  CHECK_ARG(G > 0 && Q > 0 && Q <= 256 && keys > 0 && splits && q_pad && o_floats && ml_floats, "bad arguments");
  int sms = 148;
  device_info(&sms);   // sizing only: fall back to the B200 SM count when no device is visible
  const int qtiles = (Q + 127) / 128;
  const int tiles = (keys + XT_KT - 1) / XT_KT;
  int s;
  if (xattn_variant() == 2) {
    // one CTA per SM per (key chunk, head pair, query tile, group); every chunk yields two partials per head (the
    // even and the odd key tiles), so the partial count is 2 * chunks
    int c = sms / (4 * G * qtiles);
    if (c > tiles / 4) c = tiles / 4;
    if (c < 1) c = 1;
    const int chunk = ((tiles + c - 1) / c) * XT_KT;
    c = (keys + chunk - 1) / chunk;
    s = 2 * c;
  } else {
    // one CTA per SM (the kernel owns all 512 TMEM columns and ~194 KB of shared memory): one wave of
    // splits * qtiles * G CTAs, at least two key tiles per split
    s = sms / (G * qtiles);
    if (s > tiles / 2) s = tiles / 2;
    if (s < 1) s = 1;
    const int chunk = ((tiles + s - 1) / s) * XT_KT;
    s = (keys + chunk - 1) / chunk;
  }
  *splits = s;
  *q_pad = qtiles * 128;
  *o_floats = (long long)G * s * 8 * (*q_pad) * 32;
  // (max, sum) partials, followed by the tile-skip bitmap of the tcgen05 kernel: X2_MAP_WORDS words per (group, query tile)
  *ml_floats = (long long)G * s * 8 * (*q_pad) * 2 + (long long)G * qtiles * X2_MAP_WORDS;
  return OVIS_OK;
}

int ovis_xattn(const void* q, const void* k, const void* v, const unsigned int* bits, const unsigned char* flags, int G,
               int Q, int q_stride, int keys, int splits, float* o_part, float* ml_part, void* out, void* stream) {
  CHECK_ARG(q && k && v && bits && flags && o_part && ml_part && out, "null pointer");
  CHECK_ARG(G > 0 && Q > 0 && Q <= 256 && keys > 0 && splits > 0 && q_stride >= Q, "bad arguments");
  CHECK_ARG((long long)G * keys < (1ll << 31), "too many keys");
  int rc = device_info(nullptr);
  if (rc) return rc;
  const int qtiles = (Q + 127) / 128;
  const int q_pad = qtiles * 128;
  const int tiles = (keys + XT_KT - 1) / XT_KT;
  const int variant = xattn_variant();
  const int chunks = variant == 2 ? splits / 2 : splits;
  CHECK_ARG(variant != 2 || (splits % 2 == 0), "the split count must come from ovis_xattn_plan (even)");
  const int chunk = ((tiles + chunks - 1) / chunks) * XT_KT;
  CHECK_ARG((long long)chunk * (chunks - 1) < keys, "splits too large for the key count (use ovis_xattn_plan)");
  if (variant == 0) {
    XattnArgs a;
    a.q = (const __half*)q; a.k = (const __half*)k; a.v = (const __half*)v;
    a.bits = bits; a.flags = flags; a.o_part = o_part; a.ml_part = ml_part;
    a.Q = Q; a.q_pad = q_pad; a.q_stride = q_stride;
    a.keys = keys; a.W = (keys + 31) / 32; a.splits = splits; a.chunk = chunk;
    xattn_split_kernel<<<dim3(splits, 8, G), ((Q + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(a);
    rc = check_launch("xattn_split_kernel");
    if (rc) return rc;
  } else {
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(xattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X2_SMEM);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "xattn: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
        return OVIS_ERR_CUDA;
      }
      attr_done[dev] = true;
    }
    CUtensorMap tq, tk, tv;
    rc = make_map_f16(&tq, q, (unsigned long long)G * Q, 256, 256, 128);
    if (rc) return rc;
    rc = make_map_f16(&tk, k, (unsigned long long)G * keys, 256, 256, XT_KT);
    if (rc) return rc;
    rc = make_map_f16(&tv, v, (unsigned long long)G * keys, 256, 256, XT_KT);
    if (rc) return rc;
    XattnTcArgs a;
    a.bits = bits; a.flags = flags; a.o_part = o_part; a.ml_part = ml_part;
    a.Q = Q; a.q_pad = q_pad; a.q_stride = q_stride;
    a.keys = keys; a.W = (keys + 31) / 32; a.splits = splits; a.chunk = chunk;
    static const char* tr = getenv("OVIS_XATTN_TRACE");      // device pointer (decimal) of the trace buffer
    a.trace = tr ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 10)) : nullptr;
    a.skipmap = nullptr;
    a.map_words = 0;
    if (variant == 2) {
      // tile skipping (default on; OVIS_XATTN_SKIP=0 for A/B): needs the bitmap and the per-CTA list to fit
      static const bool skip_on = !(getenv("OVIS_XATTN_SKIP") && atoi(getenv("OVIS_XATTN_SKIP")) == 0);
      const int map_words = (tiles + 31) / 32;
      if (skip_on && map_words <= X2_MAP_WORDS && (tiles + chunks - 1) / chunks <= X2_LIST_MAX) {
        uint32_t* map = reinterpret_cast<uint32_t*>(ml_part + (long long)G * splits * 8 * q_pad * 2);
        launch_k(xattn_skipmap_kernel, dim3(dim3(map_words, qtiles, G)), dim3(128), 0, (cudaStream_t)stream, bits, flags, map, Q, q_stride, keys,
                                                                                           a.W, X2_MAP_WORDS);
        rc = check_launch("xattn_skipmap_kernel");
        if (rc) return rc;
        a.skipmap = map;
        a.map_words = X2_MAP_WORDS;
      }
      launch_k(xattn_tc2_kernel, dim3(dim3(chunks * 4, qtiles, G)), dim3(X2_THREADS), X2_SMEM, (cudaStream_t)stream, tq, tk, tv, a);
      rc = check_launch("xattn_tc2_kernel");
    } else {
      xattn_tc_kernel<<<dim3(splits, qtiles, G), XT_THREADS, XT_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, a);
      rc = check_launch("xattn_tc_kernel");
    }
    if (rc) return rc;
  }
  static const bool no_combine = getenv("OVIS_XATTN_NO_COMBINE") != nullptr;     // timing experiments only
  if (no_combine) return OVIS_OK;
  if (splits <= 8) {
    // Frame decoders (a few hundred keys per group): 2-8 partials per row, one thread per output element pair
    const long long total = (long long)G * Q * 128;
    launch_k(xattn_combine_few_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, o_part, ml_part, (__half*)out, Q,
                                                                                                q_pad, splits, total);
    return check_launch("xattn_combine_few_kernel");
  }
  launch_k(xattn_combine_kernel, dim3(dim3(Q, 8, G)), dim3(256), 0, (cudaStream_t)stream, o_part, ml_part, (__half*)out, Q, q_pad, splits);
  return check_launch("xattn_combine_kernel");
}

// transposed-score kernel (xattn_tc3): chunks of 128-key tiles, one partial per chunk and head
static void xattn_t_chunks(int G, int Q, int keys, int sms, int* chunks, int* chunk) {
  const int qtiles = (Q + 127) / 128;
  const int tiles = (keys + X3_KT - 1) / X3_KT;
  int c = sms / (4 * G * qtiles);
  if (c > tiles / 4) c = tiles / 4;
  if (c < 1) c = 1;
  const int ck = ((tiles + c - 1) / c) * X3_KT;
  *chunk = ck;
  *chunks = (keys + ck - 1) / ck;
}

int ovis_xattn_plan_t(int G, int Q, int keys, int* use_t, int* splits, int* q_pad, long long* o_floats, long long* ml_floats) {
  CHECK_ARG(G > 0 && Q > 0 && Q <= 256 && keys > 0 && use_t && splits && q_pad && o_floats && ml_floats, "bad arguments");
  int sms = 148;
  device_info(&sms);
  const int qtiles = (Q + 127) / 128;
  int chunks, chunk;
  xattn_t_chunks(G, Q, keys, sms, &chunks, &chunk);
  // The transposed kernel pays one extra key tile per CTA (the reference pass) and has its chunk count as a template
  // parameter only for a single query tile: measured (profiles/xattn_t_ab_r2.txt) 1.3-1.4x over xattn_tc2 from ~100 key tiles
  // per CTA, 1.0x at 30, 0.7x at 8, 0.96x for Q = 200 (two query tiles).  OVIS_XATTN_T=0 disables it, =1 forces it (tests).
  static const int force = getenv("OVIS_XATTN_T") ? atoi(getenv("OVIS_XATTN_T")) : -1;
  const int tiles_per_cta = chunk / X3_KT;
  int use = tiles_per_cta >= 24 && qtiles == 1 && xattn_variant() == 2;
  if (force == 0) use = 0;
  if (force == 1) use = 1;
  *use_t = use;
  *splits = chunks;
  *q_pad = qtiles * 128;
  *o_floats = (long long)G * chunks * 8 * (*q_pad) * 32;
  *ml_floats = (long long)G * chunks * 8 * (*q_pad) * 2 + (long long)G * qtiles * X3_MAP_WORDS;
  return OVIS_OK;
}

int ovis_xattn_t(const void* q, const void* k, const void* v, const unsigned int* bits_t, const unsigned int* blockand,
                 const unsigned char* flags, int G, int Q, int q_stride, int keys, int splits, float* o_part, float* ml_part,
                 void* out, int* stats, void* stream) {
  CHECK_ARG(q && k && v && bits_t && flags && o_part && ml_part && out, "null pointer");
  CHECK_ARG(G > 0 && Q > 0 && Q <= 256 && keys > 0 && splits > 0 && q_stride >= Q, "bad arguments");
  CHECK_ARG((long long)G * keys < (1ll << 31), "too many keys");
  int sms = 148;
  int rc = device_info(&sms);
  if (rc) return rc;
  const int qtiles = (Q + 127) / 128;
  const int q_pad = qtiles * 128;
  const int tiles = (keys + X3_KT - 1) / X3_KT;
  const int chunk = ((tiles + splits - 1) / splits) * X3_KT;
  CHECK_ARG((long long)chunk * (splits - 1) < keys, "splits too large for the key count (use ovis_xattn_plan_t)");
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(xattn_tc3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_tc3_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_tc3_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(xattn_tc3_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_SMEM);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "xattn_t: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return OVIS_ERR_CUDA;
    }
    attr_done[dev] = true;
  }
  CUtensorMap tq, tk, tv;
  rc = make_map_f16(&tq, q, (unsigned long long)G * Q, 256, 256, 128);
  if (rc) return rc;
  rc = make_map_f16(&tk, k, (unsigned long long)G * keys, 256, 256, X3_KT);
  if (rc) return rc;
  rc = make_map_f16_sw64(&tv, v, (unsigned long long)G * keys, 256, 256, X3_KT);
  if (rc) return rc;
  XattnT3Args a;
  a.bits_t = bits_t; a.flags = flags; a.o_part = o_part; a.ml_part = ml_part;
  a.Q = Q; a.q_pad = q_pad; a.q_stride = q_stride; a.qw = 4 * qtiles;
  a.keys = keys; a.splits = splits; a.chunk = chunk;
  a.skipmap = nullptr; a.map_words = 0; a.stats = stats;
  static const char* tr3 = getenv("OVIS_XATTN_TRACE");      // device pointer (decimal) of the trace buffer
  a.trace = tr3 ? reinterpret_cast<long long*>(strtoull(tr3, nullptr, 10)) : nullptr;
  static const bool skip_on = !(getenv("OVIS_XATTN_SKIP") && atoi(getenv("OVIS_XATTN_SKIP")) == 0);
  const int map_words = (tiles + 31) / 32;
  if (skip_on && blockand && map_words <= X3_MAP_WORDS && (tiles + splits - 1) / splits <= X3_LIST_MAX) {
    uint32_t* map = reinterpret_cast<uint32_t*>(ml_part + (long long)G * splits * 8 * q_pad * 2);
    launch_k(xattn_t3_skipmap_kernel, dim3(dim3(map_words, qtiles, G)), dim3(32), 0, (cudaStream_t)stream, blockand, flags, map, Q, q_stride, a.qw, keys,
                                                                                         (keys + 31) / 32, X3_MAP_WORDS);
    rc = check_launch("xattn_t3_skipmap_kernel");
    if (rc) return rc;
    a.skipmap = map;
    a.map_words = X3_MAP_WORDS;
  }
  // chunk count of the query tile as a template parameter where every tile of the launch has the same one
  const int nq_last = Q - (qtiles - 1) * 128;
  const int nch = qtiles == 1 ? ((nq_last + 15) >> 4) : 0;
  const dim3 grid(splits * 4, qtiles, G);
  if (nch == 7) launch_k(xattn_tc3_kernel<7>, dim3(grid), dim3(X3_THREADS), X3_SMEM, (cudaStream_t)stream, tq, tk, tv, a);
  else if (nch == 8) launch_k(xattn_tc3_kernel<8>, dim3(grid), dim3(X3_THREADS), X3_SMEM, (cudaStream_t)stream, tq, tk, tv, a);
  else if (nch == 5) launch_k(xattn_tc3_kernel<5>, dim3(grid), dim3(X3_THREADS), X3_SMEM, (cudaStream_t)stream, tq, tk, tv, a);
  else launch_k(xattn_tc3_kernel<0>, dim3(grid), dim3(X3_THREADS), X3_SMEM, (cudaStream_t)stream, tq, tk, tv, a);
  rc = check_launch("xattn_tc3_kernel");
  if (rc) return rc;
  if (splits <= 8) {
    const long long total = (long long)G * Q * 128;
    launch_k(xattn_combine_few_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, o_part, ml_part, (__half*)out, Q,
                                                                                                q_pad, splits, total);
    return check_launch("xattn_combine_few_kernel");
  }
  launch_k(xattn_combine_kernel, dim3(dim3(Q, 8, G)), dim3(256), 0, (cudaStream_t)stream, o_part, ml_part, (__half*)out, Q, q_pad, splits);
  return check_launch("xattn_combine_kernel");
}

// ---- query-side chain (chain.cuh): a list of phases executed by one persistent CTA per group
struct OvisChain {
  std::vector<ChainPhase> host;
  int Q = 0, G = 0;
  bool checked = false;
  bool wide = false;              // phases spread over all CTAs, grid barrier between them (gemm_chain_wide_kernel)
  unsigned int* bar = nullptr;    // wide: {arrival count, generation}
  float* split_ws = nullptr;      // wide: fp32 scratch of the split linear + LayerNorm phases
  long long split_ws_floats = 0;
};

int ovis_chain_create(int nphases, int G, int Q, void** handle) {
  CHECK_ARG(nphases > 0 && G > 0 && Q > 0 && Q <= 256 && handle, "bad arguments (at most 256 queries per group)");
  int rc = device_info(nullptr);
  if (rc) return rc;
  OvisChain* c = new OvisChain();
  c->host.resize(nphases);
  memset(static_cast<void*>(c->host.data()), 0, sizeof(ChainPhase) * nphases);
  for (auto& p : c->host) p.kind = -1;
  c->Q = Q; c->G = G;
  *handle = c;
  return OVIS_OK;
}

int ovis_chain_set_wide(void* handle, int wide) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c, "bad arguments");
  for (auto& p : c->host) CHECK_ARG(p.kind < 0, "the mode must be chosen before the phases are set");
  if (wide && !c->bar) {
    if (cudaMalloc(&c->bar, 2 * sizeof(unsigned int)) != cudaSuccess || cudaMemset(c->bar, 0, 2 * sizeof(unsigned int)) != cudaSuccess) {
      cudaGetLastError();
      return fail(OVIS_ERR_CUDA, "%s: cannot allocate the grid barrier", "ovis_chain_set_wide");
    }
  }
  c->wide = wide != 0;
  return OVIS_OK;
}

int ovis_chain_set_parallel(void* handle, int idx, int parallel) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c && idx >= 0 && idx + 1 < (int)c->host.size(), "bad arguments");
  CHECK_ARG(c->host[idx].kind == 0, "only a plain GEMM phase (already set) can run beside its successor");
  c->host[idx].par = parallel ? 1 : 0;
  return OVIS_OK;
}

int ovis_chain_set_scratch(void* handle, float* ws, long long floats) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c && ws && floats > 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, "bad arguments");
  c->split_ws = ws;
  c->split_ws_floats = floats;
  return OVIS_OK;
}

static int chain_maps(OvisChain* c, ChainPhase& p, const void* x, int K, int ldx, const void* w, int N) {
  int rc = make_map_f16(&p.tmA, x, (unsigned long long)c->G * c->Q, (unsigned long long)K, (unsigned long long)ldx, 128);
  if (rc) return rc;
  return make_map_f16(&p.tmB, w, (unsigned long long)N, (unsigned long long)K, (unsigned long long)K, 256);
}

int ovis_chain_set_linear(void* handle, int idx, const void* x, int K, int ldx, const void* w, int N, const float* bias, float scale,
                          int relu, void* out, int ldo, int out_f32) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c && idx >= 0 && idx < (int)c->host.size() && x && w && out && K > 0 && K % 64 == 0 && N > 0 && ldx >= K && ldo >= N,
            "bad arguments");
  CHECK_ARG((N + 255) / 256 <= GEMM_MAX_NTILES, "N too large");
  ChainPhase& p = c->host[idx];
  GemmArgs& a = p.args;
  init_args(a);
  a.rows_per_group = c->Q; a.num_groups = c->G; a.a_group_stride = c->Q;
  if (c->wide) { a.rows_per_group = c->G * c->Q; a.num_groups = 1; a.a_group_stride = c->G * c->Q; }
  a.N = N; a.K = K; a.epi = EPI_STORE;
  for (int t = 0; t < (N + 255) / 256; ++t) {
    a.out[t] = out_f32 ? (void*)((float*)out + (long long)t * 256) : (void*)((__half*)out + (long long)t * 256);
    a.bias[t] = bias ? bias + (long long)t * 256 : nullptr;
  }
  a.ldo = ldo; a.out_f32 = out_f32; a.relu = relu; a.scale = scale;
  p.kind = 0;
  c->checked = false;
  return chain_maps(c, p, x, K, ldx, w, N);
}

int ovis_chain_set_linear_ln(void* handle, int idx, const void* x, int K, const void* w, const float* bias, const float* resid,
                             const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b, const float* pe,
                             int pe_period, float* y32, void* y16, void* ype16, float* d32, void* d16) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c && idx >= 0 && idx < (int)c->host.size() && x && w && bias && resid && ln1_g && ln1_b && K > 0 && K % 64 == 0,
            "bad arguments");
  CHECK_ARG((ln2_g == nullptr) == (ln2_b == nullptr) && (!ype16 || (pe && pe_period > 0)), "bad LayerNorm arguments");
  ChainPhase& p = c->host[idx];
  GemmArgs& a = p.args;
  init_args(a);
  a.rows_per_group = c->Q; a.num_groups = c->G; a.a_group_stride = c->Q;
  if (c->wide) { a.rows_per_group = c->G * c->Q; a.num_groups = 1; a.a_group_stride = c->G * c->Q; }
  a.N = 256; a.K = K; a.epi = EPI_LN;
  a.bias[0] = bias; a.resid = resid;
  a.ln1_g = ln1_g; a.ln1_b = ln1_b; a.ln2_g = ln2_g; a.ln2_b = ln2_b;
  a.pe = pe; a.pe_period = pe_period > 0 ? pe_period : 1;
  a.y32 = y32; a.y16 = (__half*)y16; a.ype16 = (__half*)ype16; a.d32 = d32; a.d16 = (__half*)d16;
  p.kind = 0;
  c->checked = false;
  if (c->wide) {
    // the wide chain never takes the row-serial LayerNorm epilogue (25-40 us per phase): K slices of 256 -> fp32 partials
    // (plain stores), grid barrier, row-parallel reduction + LayerNorm by every warp of the launch (kind 3)
    const long long rows = (long long)c->G * c->Q, rows_pad = (rows + 127) / 128 * 128;
    const int S = K / 256;
    CHECK_ARG(K % 256 == 0, "the wide chain's linear + LayerNorm needs K % 256 == 0");
    CHECK_ARG(c->split_ws && c->split_ws_floats >= (long long)S * rows_pad * 256, "ovis_chain_set_scratch: workspace missing or too small");
    init_args(a);
    a.rows_per_group = (int)rows; a.num_groups = S; a.a_group_stride = 0;
    a.a_k_mod = S; a.a_k_offset_stride = 256; a.b_k_mod = S; a.b_k_offset_stride = 256;
    a.o_group_stride = (int)rows_pad;
    a.N = 256; a.K = 256; a.epi = EPI_STORE;
    a.out[0] = c->split_ws; a.ldo = 256; a.out_f32 = 1;
    LnReduceArgs& r = p.lnr;
    r.part = c->split_ws; r.S = S; r.part_stride = rows_pad;
    r.bias = bias; r.resid = resid;
    r.ln1_g = ln1_g; r.ln1_b = ln1_b; r.ln2_g = ln2_g; r.ln2_b = ln2_b;
    r.pe = pe; r.pe_period = pe_period > 0 ? pe_period : 1;
    r.y32 = y32; r.y16 = (__half*)y16; r.ype16 = (__half*)ype16; r.d32 = d32; r.d16 = (__half*)d16;
    r.rows = (int)rows;
    p.kind = 3;
  }
  return chain_maps(c, p, x, K, K, w, 256);
}

int ovis_chain_set_self_attn(void* handle, int idx, const void* qk, const void* v, void* out) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c && idx >= 0 && idx < (int)c->host.size() && qk && v && out, "bad arguments");
  ChainPhase& p = c->host[idx];
  p.sa.qk = (const __half*)qk; p.sa.v = (const __half*)v; p.sa.out = (__half*)out; p.sa.Q = c->Q;
  p.sa.scale_log2 = 0.17677669529663687f * 1.4426950408889634f;   // 32^-1/2 * log2(e)
  p.kind = 1;
  c->checked = false;
  return OVIS_OK;
}

static int chain_run(void* handle, int first, int count, long long* trace, void* stream);

int ovis_chain_run(void* handle, int first, int count, void* stream) { return chain_run(handle, first, count, nullptr, stream); }

int ovis_chain_run_traced(void* handle, int first, int count, long long* trace, void* stream) {
  CHECK_ARG(trace, "bad arguments");
  return chain_run(handle, first, count, trace, stream);
}

static int chain_run(void* handle, int first, int count, long long* trace, void* stream) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c && first >= 0 && count > 0 && first + count <= (int)c->host.size(), "bad arguments");
  int sms = 148;
  int rc = device_info(&sms);
  if (rc) return rc;
  if (!c->checked) {
    for (auto& p : c->host) CHECK_ARG(p.kind >= 0, "a phase of the chain has not been set");
    c->checked = true;
  }
  static bool attr_done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM);
    if (e != cudaSuccess) return fail(OVIS_ERR_CUDA, "%s: cannot raise the shared-memory limit", "gemm_chain_kernel");
    attr_done[dev] = true;
  }
  CHECK_ARG(count <= CHAIN_MAX_PHASES, "too many phases in one launch");
  CHECK_ARG(c->wide || c->Q <= 128, "the one-CTA-per-group chain takes at most 128 queries per group");
  ChainLaunch cl;                                  // (copied into the launch by cudaLaunchKernel before it returns)
  memcpy(static_cast<void*>(cl.ph), c->host.data() + first, sizeof(ChainPhase) * count);
  if (c->wide) {
    CHECK_ARG(!trace, "the wide chain has no trace mode");
    static bool wattr_done[64] = {false};
    if (!wattr_done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(gemm_chain_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHAIN_SMEM);
      if (e != cudaSuccess) return fail(OVIS_ERR_CUDA, "%s: cannot raise the shared-memory limit", "gemm_chain_wide_kernel");
      wattr_done[dev] = true;
    }
    // as many CTAs as the busiest phase has tiles (or groups), within the SM budget; a COOPERATIVE launch, so that all of
    // them are resident together -- the grid barrier must never wait for a CTA that cannot be scheduled
    long long most = c->G;
    for (int i = 0; i < count; ++i) {
      const ChainPhase& p = cl.ph[i];
      if (p.kind == 0 || p.kind == 3) {
        const long long it = (long long)((p.args.rows_per_group + 127) / 128) * ((p.args.N + 255) / 256) * p.args.num_groups;
        if (it > most) most = it;
      } else if ((long long)c->G * 8 > most) {
        most = (long long)c->G * 8;
      }
    }
    // (cap: half of the SMs by default, so that the wide launches of TWO calls in flight fit side by side whatever the
    //  driver's scheduling of cooperative grids from different streams is; OVIS_CHAIN_CTAS overrides)
    static const int wide_cap = getenv("OVIS_CHAIN_CTAS") ? atoi(getenv("OVIS_CHAIN_CTAS")) : 0;
    int cap = wide_cap > 0 ? (wide_cap < sms ? wide_cap : sms) : (sms / 2 > 0 ? sms / 2 : 1);
    const int grid = (int)(most < cap ? most : cap);
    // phases that run beside their successor: the successor's items start on the CTAs after this phase's
    cl.ph[count - 1].par = 0;
    for (int i = 0, off = 0; i < count; ++i) {
      ChainPhase& p = cl.ph[i];
      p.cta_off = off % grid;
      if (p.par) {
        off += (int)((long long)((p.args.rows_per_group + 127) / 128) * ((p.args.N + 255) / 256) * p.args.num_groups % grid);
      } else {
        off = 0;
      }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = CHAIN_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, gemm_chain_wide_kernel, cl, count, c->G, c->bar);
    return check_launch("gemm_chain_wide_kernel");
  }
  const int grid = c->G < sms ? c->G : sms;
  launch_k(gemm_chain_kernel, dim3(grid), dim3(320), CHAIN_SMEM, (cudaStream_t)stream, cl, count, c->G, trace);
  return check_launch("gemm_chain_kernel");
}

int ovis_chain_upload(void* handle) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  CHECK_ARG(c, "bad arguments");
  for (auto& p : c->host) CHECK_ARG(p.kind >= 0, "a phase of the chain has not been set");
  c->checked = true;
  return OVIS_OK;
}

int ovis_chain_destroy(void* handle) {
  OvisChain* c = static_cast<OvisChain*>(handle);
  if (!c) return OVIS_OK;
  if (c->bar) cudaFree(c->bar);
  delete c;
  return OVIS_OK;
}

int ovis_self_attn(const void* qk, const void* v, void* out, int G, int Q, void* stream) {
  CHECK_ARG(qk && v && out && G > 0 && Q > 0 && Q <= 1536 && G <= 65535, "bad arguments (at most 1536 rows per group)");
  int rc = device_info(nullptr);
  if (rc) return rc;
  SelfAttnArgs a;
  a.qk = (const __half*)qk; a.v = (const __half*)v; a.out = (__half*)out; a.Q = Q;
  a.scale_log2 = 0.17677669529663687f * 1.4426950408889634f;   // 32^-1/2 * log2(e)
  static const bool simt = getenv("OVIS_SELF_ATTN") && !strcmp(getenv("OVIS_SELF_ATTN"), "simt");      // A/B and checking only
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t smem_mma = (size_t)2 * ((Q + SA_KB - 1) / SA_KB * SA_KB) * XA_LD * sizeof(__half);   // K, V of one head, 80 B per row
  if (!simt && smem_mma <= (size_t)227 * 1024) {            // (up to 1408 rows; beyond that the SIMT kernel's denser layout)
    const size_t smem = smem_mma;
    if (smem > 48 * 1024) {
      static bool attr_done[64] = {false};          // function attributes are per device
      if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(self_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return fail(OVIS_ERR_CUDA, "%s: cannot raise the shared-memory limit", "self_attn_mma_kernel");
        attr_done[dev] = true;
      }
    }
    launch_k(self_attn_mma_kernel, dim3(dim3(8, G)), dim3(256), smem, (cudaStream_t)stream, a);
    return check_launch("self_attn_mma_kernel");
  }
  const size_t smem = (size_t)Q * 32 * 2 * sizeof(__half);      // K and V of one head: 128 B per row
  if (smem > 48 * 1024) {
    static bool attr_done[64] = {false};          // function attributes are per device
    if (!attr_done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(self_attn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1536 * 128);
      if (e != cudaSuccess) return fail(OVIS_ERR_CUDA, "%s: cannot raise the shared-memory limit", "self_attn_kernel");
      attr_done[dev] = true;
    }
  }
  if (Q <= 128) self_attn_kernel<4><<<dim3(8, G), 512, smem, (cudaStream_t)stream>>>(a);
  else self_attn_kernel<2><<<dim3(8, G), 512, smem, (cudaStream_t)stream>>>(a);
  return check_launch("self_attn_kernel");
}

int ovis_clip_aggregate(const float* logits, const unsigned char* valid, float* probs, unsigned char* qvalid, int T, int Q,
                        int K, void* stream) {
  CHECK_ARG(logits && valid && probs && qvalid && T > 0 && Q > 0 && K > 0 && K <= 12000, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  clip_aggregate_kernel<<<Q, 256, (size_t)K * sizeof(float), (cudaStream_t)stream>>>(logits, valid, probs, qvalid, T, Q, K);
  return check_launch("clip_aggregate_kernel");
}

// ---- OpenVIS crop classifier front end (crop.cuh)
int ovis_mask_boxes(const float* masks, int T, int N, long long stride_t, long long stride_n, int H, int W, int logits,
                    float thresh, int* boxes, unsigned char* valid, void* stream) {
  const long long count = (long long)T * N;
  CHECK_ARG(masks && boxes && valid && T > 0 && N > 0 && count < (1ll << 31) && H > 0 && W > 0, "bad arguments");
  CHECK_ARG(thresh > 0.f && thresh < 1.f, "threshold must be inside (0, 1)");
  if (logits) thresh = logf(thresh / (1.f - thresh));       // sigmoid(x) > thresh  <=>  x > logit(thresh)
  CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  int rc = device_info(nullptr);
  if (rc) return rc;
  mask_boxes_kernel<<<(unsigned)count, 256, 0, (cudaStream_t)stream>>>(masks, N, stride_t, stride_n, H, W, thresh, boxes, valid);
  return check_launch("mask_boxes_kernel");
}

int ovis_crop_blend(const float* frames, const float* masks, long long stride_t, long long stride_n, int logits, const int* ids,
                    const int* boxes, int M, int T, int N, int H, int W, int R, void* regions_f16, void* stream) {
  CHECK_ARG(frames && masks && ids && boxes && regions_f16 && M > 0 && M <= 65535 && T > 0 && N > 0 && H > 0 && W > 0 && R > 0,
            "bad arguments (at most 65535 regions per call)");
  CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "boxes must be 16-byte aligned");
  int rc = device_info(nullptr);
  if (rc) return rc;
  CropArgs a;
  a.frames = frames; a.masks = masks; a.ids = ids; a.boxes = boxes; a.regions = (__half*)regions_f16;
  a.N = N; a.H = H; a.W = W; a.R = R;
  a.stride_t = stride_t; a.stride_n = stride_n; a.logits = logits;
  crop_blend_kernel<<<dim3((R * R + 255) / 256, M), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("crop_blend_kernel");
}

int ovis_clip_patchify(const void* regions_f16, long long M, int R, int P, const float* mean, const float* std, void* out_f16,
                       void* stream) {
  CHECK_ARG(regions_f16 && out_f16 && mean && std && M > 0 && R > 0 && P > 0 && R % P == 0, "bad arguments");
  const long long rows = M * (R / P) * (R / P);
  CHECK_ARG(rows < (1ll << 31), "too many patches");
  int rc = device_info(nullptr);
  if (rc) return rc;
  patchify_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const __half*)regions_f16, (__half*)out_f16, R, P, mean[0],
                                                                   mean[1], mean[2], 1.f / std[0], 1.f / std[1], 1.f / std[2], rows);
  return check_launch("patchify_kernel");
}

int ovis_clip_embed(const float* patch_tokens, const float* class_embedding, const float* positional_embedding,
                    const float* ln_g, const float* ln_b, float* x, long long M, int Lp, int width, void* stream) {
  CHECK_ARG(patch_tokens && class_embedding && positional_embedding && ln_g && ln_b && x && M > 0 && Lp > 0, "bad arguments");
  CHECK_ARG(width > 0 && width % 32 == 0 && width <= 1024, "width must be a multiple of 32, at most 1024");
  const long long rows = M * (1 + Lp);
  CHECK_ARG(rows < (1ll << 33), "too many tokens");
  int rc = device_info(nullptr);
  if (rc) return rc;
  clip_embed_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(patch_tokens, class_embedding, positional_embedding,
                                                                                  ln_g, ln_b, x, Lp, width, rows);
  return check_launch("clip_embed_kernel");
}

int ovis_ms_deform_attn_forward(const float* value, const long long* spatial_shapes, const long long* level_start_index,
                                const float* sampling_loc, const float* attn_weight, float* out, int N, int S, int M, int D,
                                int Lq, int L, int P, void* stream) {
  CHECK_ARG(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out, "null pointer");
  CHECK_ARG(N > 0 && S > 0 && M > 0 && D > 0 && Lq > 0 && L > 0 && P > 0, "bad sizes");
  int rc = device_info(nullptr);
  if (rc) return rc;
  MsdaArgs a;
  a.value = value; a.shapes = spatial_shapes; a.start = level_start_index; a.loc = sampling_loc; a.weight = attn_weight;
  a.out = out; a.N = N; a.S = S; a.M = M; a.D = D; a.Lq = Lq; a.L = L; a.P = P;
  const long long groups = (long long)N * Lq * M;
  const bool aligned = ((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(sampling_loc) & 7) == 0;
  const int lpg = D / 4;
  const bool vec = aligned && D % 4 == 0 && (lpg == 1 || lpg == 2 || lpg == 4 || lpg == 8 || lpg == 16 || lpg == 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) {
    const long long threads = groups * lpg;
    const long long blocks = (threads + 255) / 256;
    CHECK_ARG(blocks < (1ll << 31), "too many queries for one launch");
    switch (lpg) {
      case 1: msda_forward_vec_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(a); break;
      case 2: msda_forward_vec_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(a); break;
      case 4: msda_forward_vec_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(a); break;
      case 8: msda_forward_vec_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(a); break;
      case 16: msda_forward_vec_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(a); break;
      default: msda_forward_vec_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(a); break;
    }
    return check_launch("msda_forward_vec_kernel");
  }
  const long long total = groups * D;
  const long long blocks = (total + 255) / 256;
  CHECK_ARG(blocks < (1ll << 31), "too many outputs for one launch");
  msda_forward_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  return check_launch("msda_forward_scalar_kernel");
}

int ovis_msda_prepare(const float* proj, const float* reference_points, const long long* spatial_shapes, long long rows, int M,
                      int L, int P, int ref_dim, float* sampling_loc, float* attn_weight, void* stream) {
  CHECK_ARG(proj && reference_points && spatial_shapes && sampling_loc && attn_weight, "null pointer");
  CHECK_ARG(rows > 0 && M > 0 && L > 0 && P > 0 && (ref_dim == 2 || ref_dim == 4), "bad sizes (reference points are 2- or 4-d)");
  CHECK_ARG(rows * M < (1ll << 38), "too many queries");
  int rc = device_info(nullptr);
  if (rc) return rc;
  MsdaPrepArgs a;
  a.proj = proj; a.ref = reference_points; a.shapes = spatial_shapes; a.loc = sampling_loc; a.weight = attn_weight;
  a.rows = rows; a.M = M; a.L = L; a.P = P; a.ref_dim = ref_dim;
  msda_prepare_kernel<<<(unsigned)((rows * M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("msda_prepare_kernel");
}

int ovis_msda_fused_f16(const void* value, const float* proj, const float* reference_points, long long ref_bs,
                        const long long* spatial_shapes, const long long* level_start_index, void* out, int N, int S, int M,
                        int Lq, int L, int P, int ref_dim, void* stream) {
  CHECK_ARG(value && proj && reference_points && spatial_shapes && level_start_index && out, "null pointer");
  CHECK_ARG(N > 0 && S > 0 && M > 0 && Lq > 0 && L > 0 && P > 0 && (ref_dim == 2 || ref_dim == 4) && ref_bs >= 0, "bad sizes");
  CHECK_ARG(L * P <= 16, "at most 16 sampling points per head (L * P)");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(proj) & 7) == 0 && (M * L * P) % 2 == 0, "alignment");
  int rc = device_info(nullptr);
  if (rc) return rc;
  MsdaFusedArgs a;
  a.value = (const __half*)value; a.proj = proj; a.ref = reference_points; a.ref_bs = ref_bs; a.shapes = spatial_shapes;
  a.start = level_start_index; a.out = (__half*)out; a.N = N; a.S = S; a.M = M; a.Lq = Lq; a.L = L; a.P = P; a.ref_dim = ref_dim;
  const long long blocks = ((long long)N * Lq * M * 4 + 255) / 256;
  CHECK_ARG(blocks < (1ll << 31), "too many queries for one launch");
  msda_fused_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("msda_fused_kernel");
}

int ovis_topk_scores(const float* scores, int Q, int K, int k, float* out_scores, int* out_query, int* out_label,
                     float* out_entropy, void* stream) {
  CHECK_ARG(scores && out_scores && out_query && out_label && out_entropy, "null pointer");
  CHECK_ARG(Q > 0 && K > 0 && k > 0 && k <= 32 && (long long)Q * K >= k && (long long)Q * K < (1ll << 31), "bad sizes (k <= 32)");
  int rc = device_info(nullptr);
  if (rc) return rc;
  if (k <= 10)
    topk_scores_kernel<10, 1024><<<1, 1024, 0, (cudaStream_t)stream>>>(scores, Q, K, k, out_scores, out_query, out_label, out_entropy);
  else
    topk_scores_kernel<32, 256><<<1, 256, 0, (cudaStream_t)stream>>>(scores, Q, K, k, out_scores, out_query, out_label, out_entropy);
  return check_launch("topk_scores_kernel");
}

int ovis_mask_postprocess(const float* masks, long long q_stride, const int* query, int n_sel, int T, int h4, int w4,
                          int pad_h, int pad_w, int img_h, int img_w, int out_h, int out_w, unsigned int* bits, void* stream) {
  CHECK_ARG(masks && query && bits, "null pointer");
  CHECK_ARG(n_sel > 0 && T > 0 && h4 > 0 && w4 > 0 && pad_h > 0 && pad_w > 0 && out_h > 0 && out_w > 0, "bad sizes");
  CHECK_ARG(img_h > 0 && img_w > 0 && img_h <= pad_h && img_w <= pad_w, "the image must fit into the padded size");
  int rc = device_info(nullptr);
  if (rc) return rc;
  MaskPostArgs a;
  a.masks = masks; a.q_stride = q_stride > 0 ? q_stride : (long long)T * h4 * w4; a.query = query; a.bits = bits;
  a.n_sel = n_sel; a.T = T; a.h4 = h4; a.w4 = w4;
  a.pad_h = pad_h; a.pad_w = pad_w; a.img_h = img_h; a.img_w = img_w; a.out_h = out_h; a.out_w = out_w;
  a.words = (out_w + 31) / 32;
  CHECK_ARG((long long)n_sel * T <= 65535 && (out_h + MP_ROWS - 1) / MP_ROWS <= 65535, "too many planes / rows for one launch");
  // output = image size and an exact x4 first interpolation (the evaluation default): fixed-fraction fast path
  static const bool no_x4 = getenv("OVIS_POST_X4") && atoi(getenv("OVIS_POST_X4")) == 0;     // A/B testing
  if (!no_x4 && out_h == img_h && out_w == img_w && pad_h == 4 * h4 && pad_w == 4 * w4) {
    dim3 gridx((a.words * 32 + 255) / 256, (h4 + 1 + MPX_PAIRS - 1) / MPX_PAIRS, n_sel * T);
    mask_postprocess_x4_kernel<<<gridx, 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("mask_postprocess_x4_kernel");
  }
  dim3 grid((a.words * 32 + 255) / 256, (out_h + MP_ROWS - 1) / MP_ROWS, n_sel * T);
  mask_postprocess_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("mask_postprocess_kernel");
}

int ovis_san_pool_bias(const float* bias, float* pooled, int BN, int Q, int h, int w, int gh, int gw, void* stream) {
  CHECK_ARG(bias && pooled && BN > 0 && Q > 0 && h > 0 && w > 0 && gh > 0 && gw > 0, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  const long long total = (long long)BN * Q * gh * gw;
  san_pool_bias_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bias, pooled, total, h, w, gh, gw);
  return check_launch("san_pool_bias_kernel");
}

int ovis_san_attn(const void* qkv, const float* pooled, void* out, int B, int Q, int L, int heads, void* stream) {
  CHECK_ARG(qkv && out && B > 0 && Q >= 0 && L > 0 && heads > 0, "bad arguments");
  CHECK_ARG((size_t)(1 + L) * 64 * 2 * sizeof(__half) <= 200 * 1024, "too many patch tokens for the shared-memory K/V stage");
  int rc = device_info(nullptr);
  if (rc) return rc;
  static const bool simt = getenv("OVIS_SAN_ATTN_SIMT") != nullptr;     // checker / A-B: the SIMT kernel
  int dev = 0;
  cudaGetDevice(&dev);
  SanAttnArgs a;
  a.qkv = (const __half*)qkv; a.pooled = pooled; a.out = (__half*)out;
  a.Q = Q; a.L = L; a.heads = heads; a.B = B;
  a.scale_log2 = 0.125f * 1.4426950408889634f;   // 64^-1/2 * log2(e)
  // OVIS_SAN_ATTN=mma: the mma.sync kernel (A/B and checking); default: tcgen05 kernel for up to 256 keys (CLS + patches)
  static const bool force_mma = getenv("OVIS_SAN_ATTN") && !strcmp(getenv("OVIS_SAN_ATTN"), "mma");
  if (!simt && !force_mma && 1 + L <= ST_MAX_KEYS && heads <= 65535 && B <= 65535 &&
      (long long)B * (Q + 1 + L) < (1ll << 31) && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    CUtensorMap tm;
    rc = make_map_f16(&tm, qkv, (unsigned long long)B * (Q + 1 + L), (unsigned long long)3 * heads * 64,
                      (unsigned long long)3 * heads * 64, 128);
    if (rc) return rc;
    static bool attr_tc[64] = {false};
    if (!attr_tc[dev]) {
      cudaError_t e = cudaFuncSetAttribute(san_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM);
      if (e != cudaSuccess) return fail(OVIS_ERR_CUDA, "%s: cannot raise the shared-memory limit", "san_attn_tc_kernel");
      attr_tc[dev] = true;
    }
    int sms = 148;
    rc = device_info(&sms);
    if (rc) return rc;
    const long long items = (long long)B * heads;
    san_attn_tc_kernel<<<(unsigned)(items < sms ? items : sms), ST_THREADS, ST_SMEM, (cudaStream_t)stream>>>(tm, a);
    return check_launch("san_attn_tc_kernel");
  }
  if (!simt) {
    const int ntiles = (1 + L + SA_KT - 1) / SA_KT;
    const size_t smem2 = (size_t)ntiles * SA_KT * SA_LD * 2 * sizeof(__half);
    static size_t attr2[64] = {0};
    if (smem2 > 48 * 1024 && attr2[dev] < smem2) {
      cudaError_t e = cudaFuncSetAttribute(san_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "san_attn: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
        return OVIS_ERR_CUDA;
      }
      attr2[dev] = smem2;
    }
    const int Lt = Q + 1 + L;
    san_attn_mma_kernel<<<dim3((Lt + 63) / 64, heads, B), 128, smem2, (cudaStream_t)stream>>>(a);
    return check_launch("san_attn_mma_kernel");
  }
  const size_t smem = (size_t)(1 + L) * 64 * 2 * sizeof(__half);
  static size_t attr_smem[64] = {0};
  if (smem > 48 * 1024 && attr_smem[dev] < smem) {
    cudaError_t e = cudaFuncSetAttribute(san_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "san_attn: cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      return OVIS_ERR_CUDA;
    }
    attr_smem[dev] = smem;
  }
  san_attn_kernel<<<dim3(heads, B), 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("san_attn_kernel");
}

int ovis_san_attn_bias(const float* bias, float* out, int BN, int Q, int h, int w, int gh, int gw, void* stream) {
  CHECK_ARG(bias && out && BN > 0 && Q > 0 && h > 0 && w > 0 && gh > 0 && gw > 0, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  san_attn_bias_kernel<<<dim3(Q + 1 + gh * gw, BN), 256, 0, (cudaStream_t)stream>>>(bias, out, Q, h, w, gh, gw);
  return check_launch("san_attn_bias_kernel");
}

}  // extern "C"

// ---- temporal association ---------------------------------------------------------
int ovis_temporal_unfold_f16(const void* in, void* out, int G, int T, int C, int taps, void* stream) {
  CHECK_ARG(in && out && G > 0 && T > 0 && C > 0 && C % 8 == 0 && taps > 0 && (taps & 1), "bad arguments");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "pointers must be 16-byte aligned");
  int rc = device_info(nullptr);
  if (rc) return rc;
  const long long total = (long long)G * T * taps * (C / 8);
  temporal_unfold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)in, (__half*)out, T, C / 8, taps, total);
  return check_launch("temporal_unfold_kernel");
}

int ovis_match_embeds(const float* en, int B, int T, int n, int C, float* cost, int* pi, void* stream) {
  CHECK_ARG(en && pi && B > 0 && T > 0 && n > 0 && C > 0, "bad arguments");
  CHECK_ARG(n <= 1024 && B <= 65535, "at most 1024 queries, 65535 clips");
  int rc = device_info(nullptr);
  if (rc) return rc;
  MatchArgs a;
  a.en = en; a.cost = cost; a.pi = pi; a.T = T; a.n = n; a.C = C;
  a.cost_in_smem = match_smem_bytes(n, 1) <= (size_t)227 * 1024;
  CHECK_ARG(a.cost_in_smem || cost, "a cost scratch buffer is required when the n x n matrix does not fit shared memory");
  const size_t smem = match_smem_bytes(n, a.cost_in_smem);
  if (smem > 48 * 1024) {
    static bool attr_done[64] = {false};          // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev]) {
      cudaError_t e = cudaFuncSetAttribute(match_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return fail(OVIS_ERR_CUDA, "%s: cannot raise the shared-memory limit", "match_assign_kernel");
      attr_done[dev] = true;
    }
  }
  match_assign_kernel<<<dim3(T, B), MATCH_THREADS, smem, (cudaStream_t)stream>>>(a);
  return check_launch("match_assign_kernel");
}

int ovis_match_compose(const int* pi, long long* indices, int B, int T, int n, void* stream) {
  CHECK_ARG(pi && indices && B > 0 && T > 0 && n > 0 && B <= 65535, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  match_compose_kernel<<<dim3((n + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(pi, indices, T, n);
  return check_launch("match_compose_kernel");
}

int ovis_reorder_queries_f32(const float* in, const long long* idx, float* out, int B, int T, int n, long long inner,
                             long long stride_b, long long stride_t, long long stride_q, void* stream) {
  CHECK_ARG(in && idx && out && in != out && B > 0 && T > 0 && n > 0 && inner > 0, "bad arguments");
  CHECK_ARG(T <= 65535 && B <= 65535, "at most 65535 frames / clips");
  int rc = device_info(nullptr);
  if (rc) return rc;
  reorder_queries_kernel<<<dim3(n, T, B), 128, 0, (cudaStream_t)stream>>>(in, idx, out, T, n, inner, stride_b, stride_t, stride_q);
  return check_launch("reorder_queries_kernel");
}

int ovis_gn_stats(const float* x, double* stats, int B, int S, void* stream) {
  CHECK_ARG(x && stats && B > 0 && S > 0 && B <= 65535, "bad arguments");
  CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  int sms = 0;
  int rc = device_info(&sms);
  if (rc) return rc;
  // about four CTAs per SM over the whole batch, at least 32 rows each
  int per = (int)(((long long)S * B + 4ll * sms - 1) / (4ll * sms));
  per = ((per < 32 ? 32 : per) + 3) / 4 * 4;
  gn_stats_kernel<<<dim3((S + per - 1) / per, B), 256, 0, (cudaStream_t)stream>>>(x, stats, S, per);
  return check_launch("gn_stats_kernel");
}

int ovis_gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, float eps, int B, int H, int W,
                  int relu, const float* add, long long add_bs, long long add_cs, long long add_ps, int hs, int ws,
                  float* out32, void* out16, long long out_bs, long long out_off, void* stream) {
  CHECK_ARG(x && stats && gamma && beta && B > 0 && H > 0 && W > 0 && B <= 65535 && (out32 || out16), "bad arguments");
  CHECK_ARG(!add || (hs > 0 && ws > 0 && add_ps > 0 && add_cs > 0), "the added map needs its size and strides");
  CHECK_ARG(!add || add_cs != 1 || (add_ps % 4 == 0 && add_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(add) & 15) == 0),
            "a token-major added map must be 16-byte aligned");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out32) | reinterpret_cast<uintptr_t>(out16) |
              reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "pointers must be 16-byte aligned");
  int sms = 0;
  int rc = device_info(&sms);
  if (rc) return rc;
  GnApplyArgs a;
  a.x = x; a.stats = stats; a.gamma = gamma; a.beta = beta; a.eps = eps; a.S = H * W; a.W = W; a.relu = relu;
  a.add = add; a.add_bs = add_bs; a.add_cs = add_cs; a.add_ps = add_ps; a.hs = hs; a.ws = ws;
  a.out32 = out32; a.out16 = (__half*)out16; a.out_bs = out_bs; a.out_off = out_off;
  const int S = H * W;
  int gx = (S + 7) / 8;
  const int cap = (8 * sms + B - 1) / B;
  if (gx > cap) gx = cap;
  gn_apply_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("gn_apply_kernel");
}

int ovis_tokens_to_nchw_f32(const float* in, float* out, int B, int C, int N, long long in_bs, long long in_off, void* stream) {
  CHECK_ARG(in && out && B > 0 && C > 0 && N > 0 && B <= 65535 && in_bs >= in_off + N && in_off >= 0, "bad arguments");
  int rc = device_info(nullptr);
  if (rc) return rc;
  tokens_to_nchw_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B), 256, 0, (cudaStream_t)stream>>>(in, out, C, N, in_bs, in_off);
  return check_launch("tokens_to_nchw_kernel");
}

int ovis_conv3x3_unfold_f16(const void* in, void* out, int B, int H, int W, int C, void* stream) {
  CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad arguments");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "pointers must be 16-byte aligned");
  int rc = device_info(nullptr);
  if (rc) return rc;
  const long long total = (long long)B * H * W * 9 * (C / 8);
  CHECK_ARG((total + 255) / 256 < (1ll << 31), "too many elements for one launch");
  conv3x3_unfold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, H, W,
                                                                                             C / 8, total);
  return check_launch("conv3x3_unfold_kernel");
}

int ovis_tokens_pool_f16(const void* ft, void* g, int B, int H, int W, int s, void* stream) {
  CHECK_ARG(ft && g && B > 0 && H > 0 && W > 0 && s >= 2 && s % 2 == 0 && H % s == 0 && W % s == 0, "bad arguments");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(ft) | reinterpret_cast<uintptr_t>(g)) & 15) == 0, "pointers must be 16-byte aligned");
  int rc = device_info(nullptr);
  if (rc) return rc;
  const long long total = (long long)B * (H / s) * (W / s) * 32;
  CHECK_ARG((total + 255) / 256 < (1ll << 31), "too many elements for one launch");
  tokens_pool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)ft, (__half*)g, H, W, s, total);
  return check_launch("tokens_pool_kernel");
}

int ovis_tokens_add_pos_f16(const void* xt, const float* pos, const float* pos_t, void* xp, int B, int N, void* stream) {
  CHECK_ARG(xt && pos && xp && B > 0 && N > 0, "bad arguments");
  CHECK_ARG(((reinterpret_cast<uintptr_t>(xt) | reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(pos) |
              reinterpret_cast<uintptr_t>(pos_t)) & 15) == 0, "pointers must be 16-byte aligned");
  int rc = device_info(nullptr);
  if (rc) return rc;
  const long long total = (long long)B * N * 32;
  CHECK_ARG((total + 255) / 256 < (1ll << 31), "too many elements for one launch");
  tokens_add_pos_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)xt, pos, pos_t, (__half*)xp, N, total);
  return check_launch("tokens_add_pos_kernel");
}
