// Public (library-internal) description of one GEMM launch: problem, operand majorness and
// the fused epilogue.  Shared by the kernel, the engine and the test harness.
#pragma once
#include <cuda_runtime.h>

#include "dropout.cuh"

namespace mmu {

enum GemmEpiMode : int {
  EPI_STORE = 0,      // out = alpha*acc + bias
  EPI_QUICKGELU = 1,  // z = alpha*acc + bias ; out = z (optional) ; out2 = z*sigmoid(1.702 z)
  EPI_RESIDUAL = 2,   // out(f32) = aux(f32) + alpha*acc + bias
  EPI_DGELU = 3,      // out = alpha*acc * d/dz[z*sigmoid(1.702 z)] with z = aux (bf16)
  EPI_ATOMIC = 4,     // out(f32) += alpha*acc   (split-K partial sums)
  EPI_SOFTMAX = 5,    // out(bf16) = softmax over columns [0, n_valid) of alpha*acc, 0 beyond
                      // (batched bf16 path, N <= 128: the attention probabilities, fused)
  EPI_RESID_LN = 6,   // bf16 kernel, unbatched: out(f32) = aux(f32) + alpha*acc + bias (the residual
                      // stream, src/model.py:210-211; aux may alias out), out2(bf16, optional) =
                      // bf16(out) = the RAW operand of the next LN-folded GEMM, stats_out
                      // (optional)[row][n/128] = (sum, sum of squares) of out over a 128-column slab
};

struct GemmEpilogue {
  int mode;
  int out_bf16;  // 1: out/out2 are bf16, 0: f32
  void* out;
  void* out2;
  const float* bias;
  const void* aux;
  long long ld_out, ld_out2, ld_aux;
  // optional row remap (fuses torch.cat of the per-modality projections, src/model.py:273):
  // out_row = (r / seg_len) * seg_stride + seg_off + r % seg_len      (seg_len <= 0: identity)
  int seg_len, seg_stride, seg_off;
  float alpha;
  int n_valid;  // EPI_SOFTMAX: number of real columns (N is padded to a multiple of 8)
  // EPI_QUICKGELU / EPI_DGELU (bf16 kernel): 0 = QuickGELU z*sigmoid(1.702 z) (src/model.py:185),
  // 1 = erf-GELU z*Phi(z) (BERT's gelu of the MMBT path, src/mmbt.py:124-128)
  int act;
  // EPI_QUICKGELU / EPI_DGELU: dropout applied to z BEFORE the activation (src/model.py:195-201);
  // element counter = row * N + column.  thresh == 0: off.
  dropout::Site drop;
  // LayerNorm FOLDED into the GEMM that consumes the normalised rows (EPI_STORE / EPI_QUICKGELU of
  // the bf16 kernel, bf16 out, unbatched; src/model.py:209-212 ln_1 -> in_proj, ln_2 -> c_fc):
  // A holds the raw rows x, B holds W diag(gamma), and with the row statistics
  //   mean_r = S1_r / D, rstd_r = rsqrt(S2_r / D - mean_r^2 + eps)   (S1, S2 summed over ln_nt partials)
  //   out = rstd_r * (acc - mean_r * ln_cw[n]) + bias[n],  ln_cw[n] = sum_k B[n][k],
  //   bias[n] = b[n] + sum_k beta[k] W[n][k]      (so that out = LN(x) W^T + b exactly)
  const float* ln_stats;  // [M][ln_nt][2] partial (sum, sum of squares) per row; nullptr: off
  const float* ln_cw;     // [N]
  int ln_nt;
  float ln_inv_d, ln_eps;
  // EPI_RESID_LN: partial row statistics of `out`, [M][stats_nt][2], slab n/128 (nullptr: off)
  float* stats_out;
  int stats_nt;
};

struct GemmProblem {
  int M, N, K;
  int a_mn_major, b_mn_major;
  int splits;  // split-K factor (only with EPI_ATOMIC)
  // Batched mode (batch > 0): `batch` independent M x N x K problems addressed through 3-D tensor
  // maps (inner, mid, outer).  Batch id g selects mid coordinate g / hdiv and adds
  // (g % hdiv) * hstride + col0 to the inner coordinate; the output base moves by
  // (g / out_hdiv) * out_mid_stride + (g % out_hdiv) * out_hstride elements.  This is how the
  // per-(token position, head) attention problems of the batch-axis MHA are expressed.
  int batch;
  int a_hdiv, a_hstride, a_col0;
  int b_hdiv, b_hstride, b_col0;
  int out_hdiv, out_hstride;
  long long out_mid_stride;
  // single-CTA kernel, N <= 128: the MMA runs at N = 128 and only 128 rows of B are staged
  int half_n;
};

// Operand of a batched GEMM: a 3-D view (inner contiguous; strides in elements).
struct BatchedOperand {
  const void* base;
  long long inner, mid, outer;          // extents
  long long mid_stride, outer_stride;   // element strides (inner stride is 1)
  int mn_major;                         // 0: rows on `outer`, K on `inner`; 1: K on `outer`
  int hdiv, hstride, col0;
};
int gemm_bf16_batched_launch(const BatchedOperand& A, const BatchedOperand& B, int batch, int M,
                             int N, int K, const GemmEpilogue& e, int out_hdiv, int out_hstride,
                             long long out_mid_stride, cudaStream_t stream);

// ---------------------------------------------------------------- host side
// Returns 0 on success, negative error code otherwise (never throws).
int gemm_bf16_launch(const void* A, long long lda, const void* B, long long ldb,
                     const GemmProblem& p, const GemmEpilogue& e, cudaStream_t stream);

// fp32 SIMT twin (parity path): fp32 operands, fp32 out/aux (out_bf16 must be 0).
int gemm_f32_launch(const float* A, long long lda, const float* B, long long ldb,
                    const GemmProblem& p, const GemmEpilogue& e, cudaStream_t stream);

}  // namespace mmu
