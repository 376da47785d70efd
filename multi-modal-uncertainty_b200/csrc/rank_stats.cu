// Rank statistics of the reference's post-hoc analysis, on device:
//   * AUROC per variant            notebooks/hatefulmeme_robustness.py:22-41 (sklearn roc_auc_score),
//                                  and the eval loop's own src/framework.py:195-198
//   * Kendall tau-b @ top-5 between the heads' predictions
//                                  notebooks/analysis_round_1.py:74-113 (scipy.stats.kendalltau on the
//                                  flattened, top-5-truncated, true-class-muted (S, C) arrays)
// Both are functions of the same five INTEGER pair counts over all unordered pairs {i, j} of two
// vectors x, y: concordant, discordant, tied in x, tied in y, tied in both.  With x = the 0/1
// labels, AUROC = (conc + ties_in_y_only / 2) / (conc + disc + ties_in_y_only); tau-b =
// (conc - disc) / sqrt((T - tx)(T - ty)).  Counting pairs directly (O(n^2) compares, no sort) keeps
// the result independent of any summation order: the counts are exact, so the device result is the
// oracle's bit for bit and sklearn / scipy's up to their final floating-point division.
//
// Work decomposition: a block owns PC_IT = 1024 "i" elements (4 per thread, in registers) and one
// chunk of PC_JC = 2048 "j" elements staged in shared memory (every lane reads the same j: a
// broadcast).  Blocks entirely below the diagonal exit; blocks entirely above it and inside the
// vector run an unpredicated loop; the diagonal / ragged blocks mask pairs outside i < j < n.
// Per-thread accumulators (at most 4 * 2048 pairs each, see pair_loop) are folded into 64-bit
// totals with one warp reduction and one atomic per accumulator per warp; a one-thread-per-problem
// pass turns them into the four counts.
#include <cstdio>

#include "common.h"
#include "kernels.h"

namespace mmu {
namespace {

constexpr int PC_THREADS = 256;
constexpr int PC_IPT = 4;                     // i elements per thread
constexpr int PC_IT = PC_THREADS * PC_IPT;    // i elements per block
constexpr int PC_JC = 2048;                   // j elements per block

// sign(a - b) as a float in {-1, 0, +1} without forming the difference (no underflow, NaN -> 0):
// two FSET (float-valued compares) and one FADD
__device__ __forceinline__ float sgnf(float a, float b) {
  return (a > b ? 1.f : 0.f) - (a < b ? 1.f : 0.f);
}

// Per pair: p = sx * sy; the four accumulators  D += p (concordant - discordant),  Q += p * p
// (concordant + discordant),  NX += sx * sx (pairs NOT tied in x),  NY += sy * sy  determine every
// count (11 instructions per pair instead of 29 for predicate-and-count).  They are fp32 but only
// ever hold integers below 2^13 (a thread sees at most 4 * 2048 pairs), so they are exact.
template <bool FULL>
__device__ __forceinline__ void pair_loop(const float (&xi)[PC_IPT], const float (&yi)[PC_IPT],
                                          const long long (&ii)[PC_IPT], const float* sx,
                                          const float* sy, long long j0, int jn, long long n,
                                          float (&acc)[4]) {
#pragma unroll 4
  for (int j = 0; j < jn; ++j) {
    const float xj = sx[j], yj = sy[j];
#pragma unroll
    for (int k = 0; k < PC_IPT; ++k) {
      float a = sgnf(xi[k], xj), b = sgnf(yi[k], yj);
      if (!FULL) {  // pairs outside i < j < n contribute to no accumulator
        const bool ok = (j0 + j > ii[k]) && (ii[k] < n);
        a = ok ? a : 0.f;
        b = ok ? b : 0.f;
      }
      const float p = a * b;
      acc[0] += p;
      acc[1] = fmaf(p, p, acc[1]);
      acc[2] = fmaf(a, a, acc[2]);
      acc[3] = fmaf(b, b, acc[3]);
    }
  }
}

__global__ void __launch_bounds__(PC_THREADS)
pair_concordance_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n,
                        long long x_batch_stride, long long y_batch_stride,
                        unsigned long long* __restrict__ counts /* [batch][4]: raw D, Q, NX, NY */) {
  __shared__ float sx[PC_JC], sy[PC_JC];
  const long long i0 = static_cast<long long>(blockIdx.x) * PC_IT;
  const long long j0 = static_cast<long long>(blockIdx.y) * PC_JC;
  if (j0 + PC_JC <= i0 + 1) return;  // every j of the chunk is <= every i of the tile
  x += blockIdx.z * x_batch_stride;
  y += blockIdx.z * y_batch_stride;
  const int jn = static_cast<int>(n - j0 < PC_JC ? n - j0 : PC_JC);
  for (int j = threadIdx.x; j < jn; j += PC_THREADS) {
    sx[j] = x[j0 + j];
    sy[j] = y[j0 + j];
  }
  float xi[PC_IPT], yi[PC_IPT];
  long long ii[PC_IPT];
#pragma unroll
  for (int k = 0; k < PC_IPT; ++k) {
    ii[k] = i0 + k * PC_THREADS + threadIdx.x;
    xi[k] = ii[k] < n ? x[ii[k]] : 0.f;
    yi[k] = ii[k] < n ? y[ii[k]] : 0.f;
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (j0 >= i0 + PC_IT && i0 + PC_IT <= n) {
    pair_loop<true>(xi, yi, ii, sx, sy, j0, jn, n, acc);
  } else {
    pair_loop<false>(xi, yi, ii, sx, sy, j0, jn, n, acc);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int v = static_cast<int>(acc[c]);  // exact: an integer of magnitude <= 8192
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0)  // D may be negative: two's-complement 64-bit add
      atomicAdd(&counts[blockIdx.z * 4 + c], static_cast<unsigned long long>(static_cast<long long>(v)));
  }
}

// raw (D, Q, NX, NY) -> (concordant, discordant, tied in x, tied in y), in place
__global__ void pair_finalize_kernel(unsigned long long* __restrict__ counts, long long n, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const long long T = n * (n - 1) / 2;
  const long long D = static_cast<long long>(counts[4 * b + 0]);
  const long long Q = static_cast<long long>(counts[4 * b + 1]);
  const long long NX = static_cast<long long>(counts[4 * b + 2]);
  const long long NY = static_cast<long long>(counts[4 * b + 3]);
  counts[4 * b + 0] = static_cast<unsigned long long>((Q + D) / 2);
  counts[4 * b + 1] = static_cast<unsigned long long>((Q - D) / 2);
  counts[4 * b + 2] = static_cast<unsigned long long>(T - NX);
  counts[4 * b + 3] = static_cast<unsigned long long>(T - NY);
}

// notebooks/analysis_round_1.py:74-85 `trunk_pred_top`: per row, the threshold is the top-th largest
// value of the ORIGINAL row (np.partition(row, -top)[-top], duplicates counted); the true class is
// zeroed first when mute_true; entries below the threshold become 0.  One thread per row; the
// threshold is max{ v_c : #{k : v_k >= v_c} >= top } (O(C^2) compares, C is 10..101 here).
__global__ void top_truncate_kernel(const float* __restrict__ pred, const long long* __restrict__ labels,
                                    int N, int C, int top, int mute_true, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float* row = pred + static_cast<size_t>(r) * C;
  float value = -INFINITY;
  for (int c = 0; c < C; ++c) {
    const float v = row[c];
    if (v <= value) continue;
    int ge = 0;
    for (int k = 0; k < C; ++k) ge += row[k] >= v;
    if (ge >= top) value = v;
  }
  const int y = mute_true ? static_cast<int>(labels[r]) : -1;
  for (int c = 0; c < C; ++c) {
    const float v = c == y ? 0.f : row[c];
    out[static_cast<size_t>(r) * C + c] = v >= value ? v : 0.f;
  }
}

}  // namespace

int pair_concordance(const float* x, const float* y, long long n, int batch, long long x_batch_stride,
                     long long y_batch_stride, unsigned long long* counts, cudaStream_t stream) {
  if (n < 0 || batch < 1 || batch > 65535 || n > (1ll << 31)) return MMU_ERR_SHAPE;
  if (cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 4 * batch, stream) != cudaSuccess)
    return MMU_ERR_CUDA;
  if (n < 2) return 0;
  const unsigned gi = static_cast<unsigned>((n + PC_IT - 1) / PC_IT);
  const unsigned gj = static_cast<unsigned>((n + PC_JC - 1) / PC_JC);
  if (gj > 65535u) return MMU_ERR_SHAPE;
  pair_concordance_kernel<<<dim3(gi, gj, batch), PC_THREADS, 0, stream>>>(
      x, y, n, x_batch_stride, y_batch_stride, counts);
  pair_finalize_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(counts, n, batch);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch(2);
  return 0;
}

int top_truncate(const float* pred, const long long* labels, int N, int C, int top, int mute_true,
                 float* out, cudaStream_t stream) {
  if (N < 1 || C < 1 || top < 1 || top > C) return MMU_ERR_SHAPE;
  top_truncate_kernel<<<(N + 127) / 128, 128, 0, stream>>>(pred, labels, N, C, top, mute_true, out);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}

}  // namespace mmu
