// fp32 SIMT GEMM (FFMA) with the same operand-majorness options and fused epilogues as the
// tcgen05 kernel.  This is the fp32 PARITY path: every result that has to match the reference
// to 1e-3 relative (north_star) is produced with fp32 inputs, fp32 products and fp32
// accumulation; the bf16 tensor-core path is the throughput path.
#include <cstdio>

#include "common.h"
#include "gemm_api.h"

namespace mmu {
namespace sgemm {

constexpr int BM = 128, BN = 128, BK = 8, THREADS = 256;
constexpr int LDS = BM + 4;

__device__ __forceinline__ float quick_gelu(float z) { return z / (1.0f + expf(-1.702f * z)); }
__device__ __forceinline__ float quick_gelu_grad(float z) {
  const float s = 1.0f / (1.0f + expf(-1.702f * z));
  return s * (1.0f + 1.702f * z * (1.0f - s));
}

// Load a (rows x BK) operand tile into smem as S[k][row].  K-major: element (r,k) at
// P[r*ld + k]; MN-major: P[k*ld + r].  Out-of-range elements are zero.
__device__ __forceinline__ void load_tile(const float* __restrict__ P, long long ld, int mn_major,
                                          int r0, int k0, int R, int K, int k_end, float (&reg)[4]) {
  const int t = threadIdx.x;
  if (!mn_major) {
    const int r = r0 + (t >> 1), k = k0 + (t & 1) * 4;
    const float* src = P + (long long)r * ld + k;
    if (r < R && k + 3 < k_end && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) reg[j] = (r < R && k + j < k_end) ? src[j] : 0.f;
    }
  } else {
    const int k = k0 + (t >> 5), r = r0 + (t & 31) * 4;
    const float* src = P + (long long)k * ld + r;
    if (k < k_end && r + 3 < R && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      const float4 v = *reinterpret_cast<const float4*>(src);
      reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) reg[j] = (k < k_end && r + j < R) ? src[j] : 0.f;
    }
  }
  (void)K;
}

__device__ __forceinline__ void store_tile(float (*S)[LDS], int mn_major, const float (&reg)[4]) {
  const int t = threadIdx.x;
  if (!mn_major) {
    const int r = t >> 1, k = (t & 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) S[k + j][r] = reg[j];
  } else {
    const int k = t >> 5, r = (t & 31) * 4;
    *reinterpret_cast<float4*>(&S[k][r]) = make_float4(reg[0], reg[1], reg[2], reg[3]);
  }
}

template <int MODE>
__device__ __forceinline__ void epi4(const GemmEpilogue& e, int grow, int gcol, int N, float (&v)[4]) {
  long long orow = grow;
  if (e.seg_len > 0)
    orow = (long long)(grow / e.seg_len) * e.seg_stride + e.seg_off + grow % e.seg_len;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = gcol + j;
    if (c >= N) break;
    float x = v[j] * e.alpha + (e.bias != nullptr ? e.bias[c] : 0.f);
    if constexpr (MODE == EPI_QUICKGELU || MODE == EPI_DGELU) {  // dropout before the activation
      if (e.drop.on()) x *= e.drop.mult(static_cast<unsigned int>(grow) * static_cast<unsigned int>(N) +
                                        static_cast<unsigned int>(c));
    }
    float* out = static_cast<float*>(e.out);
    if constexpr (MODE == EPI_STORE) {
      out[orow * e.ld_out + c] = x;
    } else if constexpr (MODE == EPI_QUICKGELU) {
      if (out != nullptr) out[orow * e.ld_out + c] = x;
      static_cast<float*>(e.out2)[orow * e.ld_out2 + c] = quick_gelu(x);
    } else if constexpr (MODE == EPI_RESIDUAL) {
      out[orow * e.ld_out + c] = x + static_cast<const float*>(e.aux)[orow * e.ld_aux + c];
    } else if constexpr (MODE == EPI_DGELU) {
      out[orow * e.ld_out + c] = x * quick_gelu_grad(static_cast<const float*>(e.aux)[orow * e.ld_aux + c]);
    } else if constexpr (MODE == EPI_ATOMIC) {
      atomicAdd(out + orow * e.ld_out + c, x);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS)
sgemm_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
             const GemmProblem p, const GemmEpilogue e) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kchunk = ((p.K + p.splits - 1) / p.splits + BK - 1) / BK * BK;
  const int k_begin = blockIdx.z * kchunk;
  const int k_end = min(p.K, k_begin + kchunk);
  if (k_begin >= k_end) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  load_tile(A, lda, p.a_mn_major, m0, k_begin, p.M, p.K, k_end, ra);
  load_tile(B, ldb, p.b_mn_major, n0, k_begin, p.N, p.K, k_end, rb);
  store_tile(As[0], p.a_mn_major, ra);
  store_tile(Bs[0], p.b_mn_major, rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      load_tile(A, lda, p.a_mn_major, m0, k0 + BK, p.M, p.K, k_end, ra);
      load_tile(B, ldb, p.b_mn_major, n0, k0 + BK, p.N, p.K, k_end, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      store_tile(As[buf ^ 1], p.a_mn_major, ra);
      store_tile(Bs[buf ^ 1], p.b_mn_major, rb);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int grow = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (grow >= p.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gcol = n0 + h * 64 + tx * 4;
      if (gcol >= p.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      epi4<MODE>(e, grow, gcol, p.N, v);
    }
  }
}

template <int MODE>
int launch(const float* A, long long lda, const float* B, long long ldb, const GemmProblem& p,
           const GemmEpilogue& e, cudaStream_t s) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.splits);
  sgemm_kernel<MODE><<<grid, THREADS, 0, s>>>(A, lda, B, ldb, p, e);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    fprintf(stderr, "mmu: sgemm launch failed: %s\n", cudaGetErrorString(err));
    return MMU_ERR_CUDA;
  }
  count_launch();
  return 0;
}

}  // namespace sgemm

int gemm_f32_launch(const float* A, long long lda, const float* B, long long ldb,
                    const GemmProblem& p_in, const GemmEpilogue& e, cudaStream_t stream) {
  GemmProblem p = p_in;
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return MMU_ERR_SHAPE;
  if (e.out_bf16) return MMU_ERR_ARG;
  if (p.splits < 1) p.splits = 1;
  if (p.splits > 1 && e.mode != EPI_ATOMIC) return MMU_ERR_SHAPE;
  switch (e.mode) {
    case EPI_STORE: return sgemm::launch<EPI_STORE>(A, lda, B, ldb, p, e, stream);
    case EPI_QUICKGELU: return sgemm::launch<EPI_QUICKGELU>(A, lda, B, ldb, p, e, stream);
    case EPI_RESIDUAL: return sgemm::launch<EPI_RESIDUAL>(A, lda, B, ldb, p, e, stream);
    case EPI_DGELU: return sgemm::launch<EPI_DGELU>(A, lda, B, ldb, p, e, stream);
    case EPI_ATOMIC: return sgemm::launch<EPI_ATOMIC>(A, lda, B, ldb, p, e, stream);
    default: return MMU_ERR_ARG;
  }
}

}  // namespace mmu
