// Host-callable launchers of the row-wise / elementwise kernels (library-internal C++ API;
// the exported C ABI lives in capi.cu and include/mmu_b200.h).  All functions enqueue on
// `stream`, never allocate and never synchronise; they return 0 or a negative MMU_ERR_* code.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "dropout.cuh"

namespace mmu {

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };

// ---- input staging: token-subset gather + per-sample modality zero-fill + cast
// src fp32 [B, l_src, d] -> dst (dtype) [B, n_sel, d].  idx (device int32[n_sel]) may be null
// (identity).  keep (device int32[B, 2]) may be null; modality selects its column.
// pos_major: destination rows ordered (token position, sample) instead of (sample, position).
// src_dtype: DT_F32 (the reference's embeddings) or DT_BF16 (bf16 host staging: half the H2D bytes).
int cast_gather(const void* src, void* dst, int dst_dtype, int B, int l_src, int d, const int* idx,
                int n_sel, const int* keep, int modality, cudaStream_t stream, int pos_major = 0,
                int src_dtype = DT_F32);
int cast_f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t stream);
// guided / random modality-dropout keep mask int32[B,2] from host-drawn uniforms u, r (fp32[B]) and,
// for mode 1 (guided), device-resident per-sample scores (element b at score_*[b*score_stride])
int modality_keep_mask(const float* u, const float* r, const float* score_img, const float* score_txt,
                       int score_stride, int B, float p_drop, int mode, int* keep, cudaStream_t stream);
// packed ragged rows + int32 offsets[B+1] -> zero-padded (B, max_l, d) (src/dataset.py:216-226)
int ragged_pad(const float* packed, const int* offsets, float* out, int B, int max_l, int d,
               cudaStream_t stream);

// ---- LayerNorm (eps 1e-5, biased variance, src/model.py:174-180, :252-253)
int layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                  float* mean, float* rstd, int M, int D, cudaStream_t stream);
// y1 (fp32) = LN(x; g1, b1); y2 (y2_dtype) = LN(y1; g2, b2): ln_pre chained with the first ln_1.
int layernorm2_fwd(const float* x, const float* g1, const float* b1, float* y1, float* mean1,
                   float* rstd1, const float* g2, const float* b2, void* y2, int y2_dtype,
                   float* mean2, float* rstd2, int M, int D, cudaStream_t stream);
// Eval path with LayerNorm folded into the consumer GEMM (gemm_api.h: GemmEpilogue::ln_stats):
// y (fp32) = LN(x; gamma, beta), yraw = bf16(y) (RAW rows, the consumer's A operand),
// stats[row][nt][2]: (sum, sum of squares) of y in partial 0, zeros in partials 1..nt-1.
int layernorm_raw_stats_fwd(const float* x, const float* gamma, const float* beta, float* y,
                            void* yraw_bf16, float* stats, int nt, int M, int D, cudaStream_t stream);
// Folded weights of up to two Linears that consume LayerNorm output (N1 == 0: one):
// Wf = bf16(W diag(gamma)) [N][K], cw[n] = sum_k Wf[n][k], bf[n] = bias[n] + sum_k beta[k] W[n][k].
// n_layers > 1: the same pair for every layer in ONE launch, input pointers advancing by pstride
// floats and output pointers by wstride bytes per layer.
int ln_fold_weights(const float* W0, const float* gamma0, const float* beta0, const float* bias0,
                    void* Wf0, float* cw0, float* bf0, int N0, const float* W1, const float* gamma1,
                    const float* beta1, const float* bias1, void* Wf1, float* cw1, float* bf1, int N1,
                    int K, cudaStream_t stream, int n_layers = 1, long long pstride = 0,
                    long long wstride = 0);
// x_out = x_in + y (y and h in `dtype`); h = LN(x_out) unless gamma == nullptr (sum only).
int add_layernorm_fwd(const float* x_in, const void* y, float* x_out, const float* gamma,
                      const float* beta, void* h, int dtype, float* mean, float* rstd, int M, int D,
                      cudaStream_t stream);
// dx (fp32) = [accumulate ? dx : 0] + LN'(dy); optional low-precision copy of the final dx,
// optional column sum of the final dx (bias gradient of the layer that produced x).
int layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* mean,
                  const float* rstd, const float* gamma, float* dx, int accumulate, void* dx_lp,
                  int lp_dtype, float* dgamma, float* dbeta, float* dcolsum, int M, int D,
                  cudaStream_t stream);
int colsum_accumulate(const void* x, int dtype, float* out, int M, int N, cudaStream_t stream);

// ---- post-LN residual blocks (BERT encoder of the MMBT path, reference src/mmbt.py:90-128;
// arithmetic of pytorch_pretrained_bert's BertSelfOutput / BertOutput / BertLayerNorm, eps 1e-12)
// s_out (fp32, may be null) = x_in + y (both `dtype`; y may be null); h = LN(s; gamma, beta, eps).
// ydrop (training, hidden dropout): y is multiplied by the counter-based mask (element counter
// row * D + column, csrc/dropout.cuh) before the residual add.
int postln_fwd(const void* x_in, const void* y, float* s_out, const float* gamma, const float* beta,
               void* h, int dtype, float* mean, float* rstd, int M, int D, float eps,
               cudaStream_t stream, dropout::Site ydrop = dropout::Site{0u, 0u, 0u, 1.0f});
// dx (fp32) and dx_lp (`dtype`, may be null; with dtype fp32 pass null) = LN'(dy_branch + dy_res);
// dy_branch (`dtype`) or dy_res (fp32) may be null.  dgamma / dbeta / dcolsum(dx) accumulate.
// Dropout in the backward of a post-LN sub-layer s = x + dropout(y), h = [dropout](LN(s)):
//   in_a / in_b : mask the forward applied to the LayerNorm OUTPUT (embedding dropout); rows with
//                 row_side[row] == 0 use in_b, all others (or row_side == null) in_a;
//   out         : mask the forward applied to the branch y: dx_lp and dcolsum receive dx * mask
//                 (what flows into the branch and its bias), dx itself (the residual path) does not.
struct PostLnDropout {
  dropout::Site in_a{0u, 0u, 0u, 1.0f}, in_b{0u, 0u, 0u, 1.0f}, out{0u, 0u, 0u, 1.0f};
  const int* row_side = nullptr;
};
int postln_bwd(const void* dy_branch, const float* dy_res, int dtype, const float* x, const float* mean,
               const float* rstd, const float* gamma, float* dx, void* dx_lp, float* dgamma,
               float* dbeta, float* dcolsum, int M, int D, cudaStream_t stream,
               PostLnDropout dr = PostLnDropout{});

// ---- heads: LayerNorm(ln_post) + row gather / segment mean pooling + E small Linears.  x rows are
// position-major: row (l, b) = l*B + b (the engine's layout).
struct HeadSegments {
  int E;
  int seg_begin[16];
  int seg_end[16];  // rows [begin, end) of each sample's L rows feed head e
};
// Packed-variant pooling: segs = device int32 [V][E][2]; vec out is [V][B][E][D].
int pool_ln_fwd_variants(const float* x, const float* gamma, const float* beta, const int* segs,
                         int V, int E, float* vec, int B, int L, int D, cudaStream_t stream);
struct HeadParams {
  const float* w[16];  // (C, D) each
  const float* b[16];
  float* dw[16];
  float* db[16];
};
int pool_ln_fwd(const float* x, const float* gamma, const float* beta, const HeadSegments& seg,
                float* vec, float* mean, float* rstd, int B, int L, int D, cudaStream_t stream);
int pool_ln_bwd(const float* dvec, const float* x, const float* mean, const float* rstd,
                const float* gamma, const HeadSegments& seg, float* dx, float* dgamma, float* dbeta,
                int B, int L, int D, cudaStream_t stream);
int heads_fwd(const float* vec, const HeadParams& hp, float* logits, int B, int E, int C, int D,
              cudaStream_t stream);
int heads_bwd(const float* dlogits, const float* vec, const HeadParams& hp, float* dvec, int B,
              int E, int C, int D, cudaStream_t stream);

// ---- CLS rows (FlavaFusionTransfomerwithCLSToken, src/model.py:327-347)
int cls_fill(const float* class_emb /*(D,E)*/, float* mm_x, int B, int L, int D, int E,
             cudaStream_t stream);
int cls_bwd(const float* dmm, float* dclass_emb, int B, int L, int D, int E, cudaStream_t stream);
// split the gradient of the concatenated sequence (position-major rows (l, b)) into per-modality
// compact buffers (rows (l, b) as well)
int split_rows(const float* dmm, void* dimg, void* dtxt, int dtype, int B, int L, int off_img,
               int l_img, int l_txt, int D, cudaStream_t stream);

// ---- batch-axis attention (src/model.py:193,205-207: MHA with batch_first=False on (B,L,D))
// qkv (dtype) [B*L, 3D] packed q|k|v; out (dtype) [B*L, D]; lse fp32 [L*H*B].
// bf16 with head_dim % 64 == 0 and non-null probs/scores buffers runs on the tensor cores
// (batched tcgen05 GEMMs): probs bf16 [L*H, B, Bp] (kept for the backward), scores fp32 and dprobs
// bf16 scratch of the same shape, Bp = B rounded up to 8.  Otherwise the fp32 SIMT kernels run
// (B <= 256) and need lse / delta_ws.
// pos_major = 0: rows of qkv / out are (sample b, position l) -> b*L + l; 1: (position, sample) ->
// l*B + b, the engine's layout: the 128 rows of one (position, head) problem are then ADJACENT in
// memory (a contiguous B x 3D block per position) instead of L*3D elements apart.
// bf16, head_dim 256, B <= 128 runs ONE fused kernel (battn_fused.cu); keep_probs = 0 (eval: no
// backward will read the probabilities) leaves `probs` untouched.
int attention_fwd(const void* qkv, void* out, float* lse, void* probs, float* scores, int dtype,
                  int B, int L, int D, int H, cudaStream_t stream, int pos_major = 0,
                  int keep_probs = 1);
// the fused kernel itself: 0 launched, > 0 not applicable, < 0 error; probs (bf16 [L*H][B][Bp]) may
// be null (eval)
int fused_batch_attention_fwd(const void* qkv, void* out, void* probs, int B, int L, int D, int H,
                              int pos_major, cudaStream_t stream);
int attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                  float* delta_ws, const void* probs, float* scores, void* dprobs, void* dqkv,
                  int dtype, int B, int L, int D, int H, cudaStream_t stream, int pos_major = 0);

// ---- sequence-axis attention (BERT encoder of the MMBT path; reference call site
// src/mmbt.py:124-128 with the additive mask of :103-107).  qkv (dtype) [B*S, 3D] packed q|k|v,
// addmask fp32 [B, S] = (1 - mask) * -10000, out (dtype) [B*S, D].
// bf16 (head_dim % 64 == 0): batched tcgen05 GEMMs; probs bf16 [B*H, S, Sp] kept for the backward,
// scores fp32 / dprobs bf16 scratch of the same shape (Sp = S rounded up to 8).
// fp32: probs fp32 [B*H, S, S] kept for the backward, scores fp32 scratch (backward only).
// keep_probs == 0 (inference): the fused bf16 kernel does not write the probabilities at all.
// drop (attention-probability dropout, training): element counter ((b*H + h)*S + query)*S + key;
// the fused kernel is bypassed, `pdrop` (same shape / dtype as probs) receives the dropped copy
// that feeds P V while `probs` keeps the undropped probabilities for the backward.
int seq_attention_fwd(const void* qkv, const float* addmask, void* out, void* probs, float* scores,
                      int dtype, int B, int S, int D, int H, cudaStream_t stream, int keep_probs = 1,
                      int allow_fused = 1, dropout::Site drop = dropout::Site{0u, 0u, 0u, 1.0f},
                      void* pdrop = nullptr);
// Fused forward (fused_attention.cu): head_dim 64, S <= 512; probs may be null (inference: nothing
// but O is written).  Returns 1 when it does not apply, 0 on success, < 0 on error.
// drop (training): the undropped probabilities are stored, the dropped ones feed P V / dP is
// multiplied by the regenerated mask.
int fused_seq_attention_fwd(const void* qkv, const float* addmask, void* out, void* probs, int B, int S,
                            int D, int H, cudaStream_t stream,
                            dropout::Site drop = dropout::Site{0u, 0u, 0u, 1.0f}, void* pdrop = nullptr);
int fused_seq_attention_bwd_ds(const void* qkv, const void* dout, const void* probs, void* dprobs, void* dqkv,
                               int B, int S, int D, int H, cudaStream_t stream,
                               dropout::Site drop = dropout::Site{0u, 0u, 0u, 1.0f});
// pdrop_saved (bf16 path with dropout): the dropped probabilities the forward wrote to its `pdrop`
// argument, if the caller kept them; else they are regenerated from `probs`.
int seq_attention_bwd(const void* qkv, const void* dout, const void* probs, float* scores, void* dprobs,
                      void* dqkv, int dtype, int B, int S, int D, int H, cudaStream_t stream,
                      dropout::Site drop = dropout::Site{0u, 0u, 0u, 1.0f}, const void* pdrop_saved = nullptr);

// ---- fused softmax-CE / accuracy / uncertainty / calibration-histogram epilogue
struct MetricAccum {  // lives in device memory; all-reduced (sum) across ranks
  unsigned long long conf_count[15];
  unsigned long long conf_correct[15];
  unsigned long long hpred_count[32];
  unsigned long long mi_count[32];
  unsigned long long n_samples;
  unsigned long long n_rows;          // rows that contributed to loss_sum
  unsigned long long n_correct_rows;  // `acc` numerator (train: per head row; eval: mean logits)
  unsigned long long n_correct_prob;  // argmax of the mean probability == label
  double conf_sum[15];
  double loss_sum;
  double sum_h_pred, sum_h_exp, sum_mi;
};
// mode 0 (train): CE per (sample, head) row vs labels[n*label_stride + e*label_estride];
// mode 1 (eval): CE on the head-mean logits vs labels[n*label_stride].
int ce_uncertainty(const float* logits, const long long* labels, int label_stride,
                   int label_estride, int N, int E, int C, int mode, float grad_scale,
                   float* dlogits, int* pred_out, float* scores_out /*[N,4]*/, MetricAccum* acc,
                   cudaStream_t stream);

// ---- post-hoc robustness scoring (notebooks/utils.py:22-34, food101_robustness.py:24-77)
struct PosthocAccum {  // device memory; sums are all-reducible
  double sx[2], sy[2], sxx[2], syy[2], sxy[2];  // [image, text] Pearson sufficient statistics
  unsigned long long n_samples;
  unsigned long long correct[128];               // head-mean-logit argmax == label, per variant
};
int posthoc_scoring(const float* logits /*(V,B,E,C)*/, const long long* labels /*(B)*/, int V, int B,
                    int E, int C, int n_repeats, float* p_true_out /*(B,V) or null*/,
                    PosthocAccum* acc, cudaStream_t stream);

// ---- rank statistics (AUROC: notebooks/hatefulmeme_robustness.py:22-41, src/framework.py:195-198;
// Kendall tau-b @ top-k: notebooks/analysis_round_1.py:74-113).  counts[b] = {concordant,
// discordant, tied in x (joint ties included), tied in y (joint ties included)} over all unordered
// pairs of (x + b*x_batch_stride, y + b*y_batch_stride)[0..n); overwritten, exact integers.
int pair_concordance(const float* x, const float* y, long long n, int batch, long long x_batch_stride,
                     long long y_batch_stride, unsigned long long* counts /*[batch][4]*/,
                     cudaStream_t stream);
int top_truncate(const float* pred /*(N,C)*/, const long long* labels /*(N) or null*/, int N, int C,
                 int top, int mute_true, float* out /*(N,C)*/, cudaStream_t stream);

// ---- fused AdamW over the flat parameter buffer (train.py:196-202 hyper-parameters)
int adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr,
               float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
               cudaStream_t stream);

}  // namespace mmu
