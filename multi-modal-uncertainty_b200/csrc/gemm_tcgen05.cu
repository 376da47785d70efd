// Host side of the tcgen05 GEMM: tensor-map encoding and launch.
#include "gemm_tcgen05.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.h"

namespace mmu {

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime API so that the shared library
// still loads (and exports its symbols) on a machine without a driver.
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

}  // namespace

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld`
// elements; box = box_inner x box_outer, 128-byte swizzle, out-of-bounds reads give zeros.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, long long inner, long long outer,
                      long long ld, int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return MMU_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) return MMU_ERR_ALIGN;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MMU_ERR_TMAP;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// 3-D bf16 tensor map (inner contiguous), 128-byte swizzle, zero fill out of bounds.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, long long inner, long long mid,
                      long long outer, long long mid_stride, long long outer_stride, int box_inner,
                      int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return MMU_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (mid_stride * 2) % 16 != 0 ||
      (outer_stride * 2) % 16 != 0)
    return MMU_ERR_ALIGN;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(mid),
                        static_cast<cuuint64_t>(outer)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(mid_stride) * 2,
                           static_cast<cuuint64_t>(outer_stride) * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_inner), 1u, static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MMU_ERR_TMAP;
}

// SMs the persistent GEMM grids may occupy (0 = all).  The data-parallel wrapper lowers it while a
// gradient all-reduce runs next to the backward: a persistent grid of 148 CTAs that finds some SMs
// taken by NCCL's CTAs runs its last CTAs -- each with a full static share of the tiles -- AFTER the
// others, which can double the kernel; a grid sized to the SMs that are really free cannot.
static std::atomic<int> g_sm_limit{0};
void set_gemm_sm_limit(int n) { g_sm_limit.store(n < 0 ? 0 : n, std::memory_order_relaxed); }
int gemm_sm_limit() { return g_sm_limit.load(std::memory_order_relaxed); }
int gemm_sms() {
  const int lim = g_sm_limit.load(std::memory_order_relaxed);
  const int n = sm_count();
  return (lim > 0 && lim < n) ? (lim < 2 ? 2 : lim) : n;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// 4-D output / epilogue-operand map: (columns, rows, batch % hdiv, batch / hdiv), element strides
// (1, ld, hstride, mid_stride); box = 32 rows x 64 bytes, 64-byte swizzle.  Stores clip at the
// extents, loads zero-fill, so the kernel needs no edge predicates.
int make_tmap_out_4d(CUtensorMap* out, const void* base, int is_bf16, long long cols, long long rows,
                     long long ld, long long hdiv, long long hstride, long long nmid,
                     long long mid_stride) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return MMU_ERR_DRIVER;
  const long long es = is_bf16 ? 2 : 4;
  if (hdiv < 1) hdiv = 1;
  if (nmid < 1) nmid = 1;
  if (hdiv == 1 || hstride <= 0) hstride = ld;     // size-1 dimensions still need a legal stride
  if (nmid == 1 || mid_stride <= 0) mid_stride = ld;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * es) % 16 != 0 ||
      (hstride * es) % 16 != 0 || (mid_stride * es) % 16 != 0)
    return MMU_ERR_ALIGN;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(hdiv), static_cast<cuuint64_t>(nmid)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(ld * es), static_cast<cuuint64_t>(hstride * es),
                           static_cast<cuuint64_t>(mid_stride * es)};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(is_bf16 ? 32 : 16), 32u, 1u, 1u};  // 64-byte rows
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  4, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MMU_ERR_TMAP;
}

namespace {
// MMU_GEMM_NCTA=1 forces the single-CTA kernel (A/B measurements); default: CTA pairs where they fit
int pair_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("MMU_GEMM_NCTA");
    v = (s != nullptr && s[0] == '1') ? 0 : 1;
  }
  return v;
}

// CTA pairs (256x256 tiles) for large unbatched problems
// (N <= 128 would waste half of every 256-wide tile: those run the single-CTA kernel at N = 128)
bool use_pair(const GemmProblem& p) {
  return p.batch == 0 && p.M >= 512 && p.N > gemm::BN / 2 && pair_mode_enabled();
}

template <int MODE, bool OBF, int NCTA, bool FOLD = false>
int launch_kernel(int grid, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& c0,
                  const CUtensorMap& c1, const CUtensorMap& c2, const GemmProblem& p,
                  const GemmEpilogue& e, cudaStream_t stream) {
  using namespace gemm;
  auto kernel = gemm_bf16_tcgen05_kernel<MODE, OBF, NCTA, FOLD>;
  static cudaError_t attr_err =
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (attr_err != cudaSuccess) {
    fprintf(stderr, "mmu: cudaFuncSetAttribute(smem=%d): %s\n", SMEM_BYTES,
            cudaGetErrorString(attr_err));
    return MMU_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS + (FOLD ? 32 : 0));  // FOLD: + the statistics warp
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  // programmatic dependent launch: the kernel's prologue may overlap the previous kernel's tail
  // (it calls griddepcontrol.wait before it touches global memory)
  static const bool pdl = getenv("MMU_GEMM_NO_PDL") == nullptr;  // A/B switch
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (NCTA == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = NCTA;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, ta, tb, c0, c1, c2, p, e);
  if (err != cudaSuccess) {
    fprintf(stderr, "mmu: gemm launch failed: %s\n", cudaGetErrorString(err));
    return MMU_ERR_CUDA;
  }
  count_launch();
  return 0;
}

// One CTA per 128x256 tile, or -- for large unbatched problems -- a CTA pair per 256x256 tile.
template <int MODE, bool OBF>
int launch_mode(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& c0,
                const CUtensorMap& c1, const GemmProblem& p, const GemmEpilogue& e,
                cudaStream_t stream, const CUtensorMap* c2p = nullptr) {
  const CUtensorMap& c2 = c2p != nullptr ? *c2p : c1;
  using namespace gemm;
  const int nbatch = p.batch > 0 ? p.batch : 1;
  const long long n_tiles = (p.N + BN - 1) / BN;
  constexpr bool CAN_FOLD = (MODE == EPI_STORE || MODE == EPI_QUICKGELU) && OBF;
  const bool fold = e.ln_stats != nullptr;
  if (fold && !CAN_FOLD) return MMU_ERR_ARG;
  if (use_pair(p)) {
    const long long tiles = ((p.M + 2 * BM - 1) / (2 * BM)) * n_tiles * p.splits;
    const long long clusters = gemm_sms() / 2;
    const int grid = 2 * static_cast<int>(tiles < clusters ? tiles : clusters);
    if constexpr (CAN_FOLD) {
      if (fold) return launch_kernel<MODE, OBF, 2, true>(grid, ta, tb, c0, c1, c2, p, e, stream);
    }
    return launch_kernel<MODE, OBF, 2>(grid, ta, tb, c0, c1, c2, p, e, stream);
  }
  const long long tiles = nbatch * ((p.M + BM - 1) / BM) * n_tiles * p.splits;
  const int grid = static_cast<int>(tiles < gemm_sms() ? tiles : gemm_sms());
  if constexpr (CAN_FOLD) {
    if (fold) return launch_kernel<MODE, OBF, 1, true>(grid, ta, tb, c0, c1, c2, p, e, stream);
  }
  return launch_kernel<MODE, OBF, 1>(grid, ta, tb, c0, c1, c2, p, e, stream);
}

// Output maps + mode dispatch shared by the plain and the batched entry points.  `rows`/`cols`
// are the per-problem output extents; batch geometry as in GemmProblem.
int launch_with_epilogue(const CUtensorMap& ta, const CUtensorMap& tb, const GemmProblem& p,
                         const GemmEpilogue& e, cudaStream_t stream) {
  const int obf = e.out_bf16 ? 1 : 0;
  const long long hdiv = p.batch > 0 ? p.out_hdiv : 1;
  const long long nmid = p.batch > 0 ? (p.batch + hdiv - 1) / hdiv : 1;
  auto omap = [&](CUtensorMap* m, const void* base, long long ld) {
    return make_tmap_out_4d(m, base, obf, p.N, p.M, ld, hdiv, p.out_hstride, nmid, p.out_mid_stride);
  };
  CUtensorMap c0, c1;
  int rc;
  if (e.ln_stats != nullptr) {  // LayerNorm folded into the epilogue: see gemm_api.h
    if ((e.mode != EPI_STORE && e.mode != EPI_QUICKGELU) || !obf || p.batch > 0 || e.ln_cw == nullptr ||
        e.ln_nt < 1 || e.ln_nt > gemm::LN_MAX_NT || e.alpha != 1.0f || (reinterpret_cast<uintptr_t>(e.ln_cw) & 3) != 0 ||
        (reinterpret_cast<uintptr_t>(e.ln_stats) & 7) != 0)
      return MMU_ERR_ARG;
  }
  switch (e.mode) {
    case EPI_RESID_LN: {
      // out fp32 (c0), optional bf16 copy out2 (c1), fp32 residual aux (c2)
      if (obf || e.out == nullptr || e.aux == nullptr || p.batch > 0 || p.splits > 1) return MMU_ERR_ARG;
      if (e.stats_out != nullptr && (e.stats_nt < (p.N + gemm::BN / 2 - 1) / (gemm::BN / 2) ||
                                     (reinterpret_cast<uintptr_t>(e.stats_out) & 7) != 0))
        return MMU_ERR_ARG;
      CUtensorMap c2;
      if ((rc = omap(&c0, e.out, e.ld_out)) != 0) return rc;
      if ((rc = omap(&c2, e.aux, e.ld_aux)) != 0) return rc;
      c1 = c0;
      if (e.out2 != nullptr &&
          (rc = make_tmap_out_4d(&c1, e.out2, 1, p.N, p.M, e.ld_out2, hdiv, p.out_hstride, nmid,
                                 p.out_mid_stride)) != 0)
        return rc;
      return launch_mode<EPI_RESID_LN, false>(ta, tb, c0, c1, p, e, stream, &c2);
    }
    case EPI_STORE:
      if (e.out == nullptr) return MMU_ERR_ARG;
      if ((rc = omap(&c0, e.out, e.ld_out)) != 0) return rc;
      c1 = c0;
      return obf ? launch_mode<EPI_STORE, true>(ta, tb, c0, c1, p, e, stream)
                 : launch_mode<EPI_STORE, false>(ta, tb, c0, c1, p, e, stream);
    case EPI_QUICKGELU:
      if (!obf || e.out2 == nullptr) return MMU_ERR_ARG;  // bf16 activations only on this path
      if ((rc = omap(&c1, e.out2, e.ld_out2)) != 0) return rc;
      c0 = c1;
      if (e.out != nullptr && (rc = omap(&c0, e.out, e.ld_out)) != 0) return rc;
      return launch_mode<EPI_QUICKGELU, true>(ta, tb, c0, c1, p, e, stream);
    case EPI_DGELU:
      if (!obf || e.out == nullptr || e.aux == nullptr) return MMU_ERR_ARG;
      if ((rc = omap(&c0, e.out, e.ld_out)) != 0) return rc;
      if ((rc = omap(&c1, e.aux, e.ld_aux)) != 0) return rc;
      return launch_mode<EPI_DGELU, true>(ta, tb, c0, c1, p, e, stream);
    case EPI_SOFTMAX:
      if (!obf || e.out == nullptr || p.batch <= 0 || p.N > gemm::BN / 2 || e.n_valid < 1 ||
          e.n_valid > p.N)
        return MMU_ERR_ARG;
      if ((rc = omap(&c0, e.out, e.ld_out)) != 0) return rc;
      c1 = c0;
      return launch_mode<EPI_SOFTMAX, true>(ta, tb, c0, c1, p, e, stream);
    case EPI_ATOMIC:
      if (obf || e.out == nullptr) return MMU_ERR_ARG;
      if ((rc = omap(&c0, e.out, e.ld_out)) != 0) return rc;
      c1 = c0;
      return launch_mode<EPI_ATOMIC, false>(ta, tb, c0, c1, p, e, stream);
    default:
      // EPI_RESIDUAL exists on the fp32 path only: the bf16 engine fuses the residual add into
      // the LayerNorm kernel that consumes the sum (rowops.cu add_layernorm_fwd)
      return MMU_ERR_ARG;
  }
}
}  // namespace

int gemm_bf16_launch(const void* A, long long lda, const void* B, long long ldb,
                     const GemmProblem& p_in, const GemmEpilogue& e, cudaStream_t stream) {
  using namespace gemm;
  GemmProblem p = p_in;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return MMU_ERR_SHAPE;
  if (e.seg_len > 0) {
    // Row remap (fused torch.cat): one batched problem per segment -- batch g holds rows
    // [g*seg_len, (g+1)*seg_len) of A and writes rows g*seg_stride + seg_off + l of `out`.
    if (p.a_mn_major || p.b_mn_major || p.splits > 1 || e.mode != EPI_STORE || p.M % e.seg_len != 0)
      return MMU_ERR_SHAPE;
    const int nb = p.M / e.seg_len;
    BatchedOperand a{}, b{};
    a.base = A; a.inner = p.K; a.mid = nb; a.outer = e.seg_len;
    a.mid_stride = static_cast<long long>(e.seg_len) * lda; a.outer_stride = lda;
    a.mn_major = 0; a.hdiv = 1; a.hstride = 0; a.col0 = 0;
    b.base = B; b.inner = p.K; b.mid = 1; b.outer = p.N;
    b.mid_stride = static_cast<long long>(p.N) * ldb; b.outer_stride = ldb;
    b.mn_major = 0; b.hdiv = nb; b.hstride = 0; b.col0 = 0;  // g / nb == 0: every batch reads B
    GemmEpilogue eb = e;
    eb.seg_len = 0;
    const long long es = e.out_bf16 ? 2 : 4;
    eb.out = static_cast<char*>(e.out) + static_cast<long long>(e.seg_off) * e.ld_out * es;
    return gemm_bf16_batched_launch(a, b, nb, e.seg_len, p.N, p.K, eb, 1, 0,
                                    static_cast<long long>(e.seg_stride) * e.ld_out, stream);
  }
  const int kb_total = (p.K + BK - 1) / BK;
  if (p.splits < 1) p.splits = 1;
  if (p.splits > 1 && e.mode != EPI_ATOMIC) return MMU_ERR_SHAPE;
  if (p.splits > kb_total) p.splits = kb_total;
  {  // no split may be empty: the epilogue would add an undefined accumulator
    const int kb_per = (kb_total + p.splits - 1) / p.splits;
    p.splits = (kb_total + kb_per - 1) / kb_per;
  }
  p.batch = 0;
  p.half_n = !use_pair(p) && p.N <= BN / 2;
  CUtensorMap ta, tb;
  int rc;
  if (!p.a_mn_major) rc = make_tmap_bf16_2d(&ta, A, p.K, p.M, lda, BK, BM);
  else               rc = make_tmap_bf16_2d(&ta, A, p.M, p.K, lda, 64, BK);
  if (rc != 0) return rc;
  // a CTA of a pair stages only its half of the B tile: the K-major box shrinks to 128 rows
  const int b_rows = (use_pair(p) || p.half_n) ? BN / 2 : BN;
  if (!p.b_mn_major) rc = make_tmap_bf16_2d(&tb, B, p.K, p.N, ldb, BK, b_rows);
  else               rc = make_tmap_bf16_2d(&tb, B, p.N, p.K, ldb, 64, BK);
  if (rc != 0) return rc;

  return launch_with_epilogue(ta, tb, p, e, stream);
}

int gemm_bf16_batched_launch(const BatchedOperand& A, const BatchedOperand& B, int batch, int M,
                             int N, int K, const GemmEpilogue& e, int out_hdiv, int out_hstride,
                             long long out_mid_stride, cudaStream_t stream) {
  using namespace gemm;
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return MMU_ERR_SHAPE;
  if (e.mode != EPI_STORE && e.mode != EPI_SOFTMAX) return MMU_ERR_ARG;
  GemmProblem p{};
  p.M = M; p.N = N; p.K = K;
  p.a_mn_major = A.mn_major; p.b_mn_major = B.mn_major;
  p.splits = 1;
  p.batch = batch;
  p.a_hdiv = A.hdiv; p.a_hstride = A.hstride; p.a_col0 = A.col0;
  p.b_hdiv = B.hdiv; p.b_hstride = B.hstride; p.b_col0 = B.col0;
  p.out_hdiv = out_hdiv < 1 ? 1 : out_hdiv; p.out_hstride = out_hstride;
  p.out_mid_stride = out_mid_stride;
  p.half_n = N <= BN / 2;  // the MMA runs at N = 128 and only 128 rows of B are staged
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_3d(&ta, A.base, A.inner, A.mid, A.outer, A.mid_stride, A.outer_stride,
                             A.mn_major ? 64 : BK, A.mn_major ? BK : BM);
  if (rc != 0) return rc;
  rc = make_tmap_bf16_3d(&tb, B.base, B.inner, B.mid, B.outer, B.mid_stride, B.outer_stride,
                         B.mn_major ? 64 : BK, B.mn_major ? BK : (p.half_n ? BN / 2 : BN));
  if (rc != 0) return rc;
  return launch_with_epilogue(ta, tb, p, e, stream);
}

}  // namespace mmu
