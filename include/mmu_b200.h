/* mmu_b200 -- C ABI of the B200-native hot path of wooginawunan/multi-modal-uncertainty.
 *
 * The reference has no FFI / plugin registry: its boundary is a Python protocol
 * (SURVEY.md section 8b).  This header is therefore the interface a maintainer would bind from
 * Python (ctypes) to replace, call site by call site, the ATen operators
 * the reference path dispatches to.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless stated otherwise;
 *   - the caller owns every buffer; the library never allocates, frees or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); calls are re-entrant per
 *     stream;
 *   - return value: 0 on success, a negative MMU_ERR_* code otherwise; nothing throws;
 *   - dtype arguments: MMU_F32 or MMU_BF16 (storage of GEMM operands / activations); all
 *     accumulation, statistics, losses and optimiser state are fp32;
 *   - there is NO CPU fallback: without a B200 and libcuda every compute entry point returns
 *     MMU_ERR_DRIVER / MMU_ERR_CUDA.
 */
#ifndef MMU_B200_H_
#define MMU_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MMU_API __attribute__((visibility("default")))
#else
#define MMU_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MMU_OK 0
#define MMU_ERR_SHAPE (-1)
#define MMU_ERR_ALIGN (-2)
#define MMU_ERR_DRIVER (-3)
#define MMU_ERR_TMAP (-4)
#define MMU_ERR_CUDA (-5)
#define MMU_ERR_ARG (-6)
#define MMU_ERR_WORKSPACE (-7)

#define MMU_F32 0
#define MMU_BF16 1

/* GEMM epilogue modes */
#define MMU_EPI_STORE 0
#define MMU_EPI_QUICKGELU 1
#define MMU_EPI_RESIDUAL 2
#define MMU_EPI_DGELU 3
#define MMU_EPI_ATOMIC 4
#define MMU_EPI_RESID_LN 6

MMU_API const char* mmu_version(void);
MMU_API const char* mmu_error_string(int code);
/* Number of CUDA kernels this library has launched so far in this process (host counter). */
MMU_API long long mmu_launch_count(void);
/* SM budget of the persistent tensor-core GEMM grids: n > 0 limits them to n SMs, 0 restores all.
 * Returns the budget in effect.  Host-side, process-wide; used by the data-parallel wrapper while an
 * NCCL gradient all-reduce shares the GPU with the backward pass (no reference counterpart: the
 * reference has no distributed code). */
MMU_API int mmu_set_gemm_sm_limit(int n);
/* sizeof() of the ABI structs as compiled into the library, so a binding can verify its mirror:
 * 0 mmu_flava_config, 1 mmu_flava_inputs, 2 mmu_gemm_epilogue, 3 mmu_metric_accum,
 * 4 mmu_param_entry, 5 mmu_posthoc_accum; -1 for anything else. */
MMU_API int mmu_struct_size(int which);

/* ------------------------------------------------------------------------------------------
 * Dense contraction  C[M,N] = epilogue( sum_k A(m,k) B(n,k) )
 * Replaces nn.Linear / F.linear (src/model.py:262,264 projections; :193 MHA in_proj/out_proj;
 * :195-201 MLP c_fc/c_proj) and their autograd backward (dgrad / wgrad).
 *   a_mn_major = 0: A stored [M][K] (lda = row pitch);  1: A stored [K][M].  Same for B/[N].
 *   dtype MMU_BF16: tcgen05/TMEM/TMA tensor-core kernel, fp32 accumulate;
 *   dtype MMU_F32 : fp32 FFMA kernel (the 1e-3 parity path).
 * Epilogues (alpha scales the accumulator, bias is fp32[N] or NULL):
 *   STORE      out = alpha*acc + bias                      (out: dtype if out_lp else fp32)
 *   QUICKGELU  z = alpha*acc + bias; out = z (may be NULL); out2 = z*sigmoid(1.702 z)
 *              (src/model.py:183-185 QuickGELU fused behind c_fc)
 *   RESIDUAL   out(fp32) = aux(fp32) + alpha*acc + bias    (src/model.py:210-211); dtype MMU_F32
 *              only -- the bf16 path returns MMU_ERR_ARG: its residual add is fused into the
 *              LayerNorm kernel that consumes the sum
 *   DGELU      out = alpha*acc * d/dz QuickGELU(z), z = aux (dtype)
 *   ATOMIC     out(fp32) += alpha*acc, split-K `splits` ways (weight gradients)
 *   RESID_LN   (MMU_BF16 only; the eval path of src/model.py:209-212) out(fp32) = aux(fp32) + alpha*acc
 *              + bias -- the residual stream; aux may alias out --, out2 (bf16, may be NULL) = the
 *              same values rounded, i.e. the RAW rows the next GEMM consumes, stats_out (may be NULL)
 *              [M][stats_nt][2] = (sum, sum of squares) of out over each 128-column slab
 * LayerNorm FOLDED into the consumer (STORE / QUICKGELU, MMU_BF16, out_lp = 1): with ln_stats !=
 * NULL, A holds raw rows x, B holds W diag(gamma) (mmu_ln_fold_weights), and
 *   out = rstd_r (acc - mean_r ln_cw[n]) + bias[n],   mean_r, rstd_r from the ln_nt partial
 *   (sum, sum of squares) pairs of row r (eps ln_eps, 1/D = ln_inv_d), ln_cw[n] = sum_k B[n][k],
 *   bias = the folded bias of mmu_ln_fold_weights  =>  out = LayerNorm(x) W^T + b.
 * seg_len > 0 remaps output rows r -> (r/seg_len)*seg_stride + seg_off + r%seg_len, which writes
 * a per-modality projection straight into the concatenated sequence (fuses torch.cat :273). */
typedef struct {
  int mode;
  int out_lp;      /* 1: out/out2 stored in `dtype`; 0: fp32 */
  void* out;
  void* out2;
  const float* bias;
  const void* aux;
  long long ld_out, ld_out2, ld_aux;
  int seg_len, seg_stride, seg_off;
  float alpha;
  /* QUICKGELU / DGELU only: dropout on z BEFORE the activation (src/model.py:195-201); element
   * counter = row * N + column, mask function of csrc/dropout.cuh.  drop_p == 0: off. */
  float drop_p;
  int drop_site;
  unsigned long long drop_seed;
  /* LayerNorm folded into the epilogue (see above); ln_stats == NULL: off */
  const float* ln_stats;
  const float* ln_cw;
  int ln_nt;
  float ln_inv_d, ln_eps;
  /* RESID_LN: per-slab row sums of out; NULL: not written.  stats_nt >= ceil(N / 128) */
  int stats_nt;
  float* stats_out;
} mmu_gemm_epilogue;

MMU_API int mmu_gemm(int dtype, const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
             int b_mn_major, int M, int N, int K, int splits, const mmu_gemm_epilogue* epi,
             void* stream);

/* Companions of the folded-LayerNorm eval path.
 * mmu_ln_fold_weights: Wf[n][k] = bf16(W[n][k] gamma[k]), cw[n] = sum_k Wf[n][k],
 *   bf[n] = bias[n] + sum_k beta[k] W[n][k]   (W fp32 [N][K]; bias may be NULL).
 * mmu_layernorm_raw_stats: y (fp32) = LayerNorm(x; gamma, beta) (eps 1e-5), yraw = bf16(y),
 *   stats[row][nt][2] = (sum, sum of squares) of y in partial 0, zeros in partials 1..nt-1 -- the
 *   ln_stats layout a following folded GEMM expects (ln_pre feeding the first block's ln_1). */
MMU_API int mmu_ln_fold_weights(const float* W, const float* gamma, const float* beta, const float* bias,
                                void* Wf_bf16, float* cw, float* bf, int N, int K, void* stream);
MMU_API int mmu_layernorm_raw_stats(const float* x, const float* gamma, const float* beta, float* y,
                                    void* yraw_bf16, float* stats, int nt, int M, int D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input staging: token-subset gather + per-sample modality zero-fill + cast.
 * Replaces the fancy indexing `img[:, indices_img, :]` of eval_transformer_robustness.py:118-119,
 * the zero-fill masking of eval_robustness.py:92-97 and the H2D-side dtype handling.
 * src fp32 (B, l_src, d) -> dst (B, n_sel, d), or (n_sel, B, d) when pos_major != 0 (the engine's
 * position-major row order).  idx: int32[n_sel] or NULL (identity); keep: int32[B,2] or NULL;
 * `modality` (0 image, 1 text) selects the keep column. */
MMU_API int mmu_mask_gather_tokens(const float* src, void* dst, int dst_dtype, int B, int l_src, int d,
                           const int* idx, int n_sel, const int* keep, int modality, int pos_major,
                           void* stream);

/* Guided / random modality dropout (BASELINE.json configs[1]; the reference only NAMES it --
 * configs/training_guided.gin:10-18 -- so the definition is this repo's, oracle/shaping.py
 * modality_dropout_mask): keep int32[B,2] over (image, text) for mmu_mask_gather_tokens /
 * mmu_flava_inputs.keep.  u, r: fp32[B] uniform draws of the HOST generator (device copies).
 * Sample b keeps both modalities unless u[b] < p_drop; then mode 0 (random) drops image when
 * r[b] < 0.5 else text; mode 1 (guided) drops the modality with the HIGHER score
 * (score_img[b*score_stride] vs score_txt[b*score_stride], device fp32, ties -> image) -- the
 * scores stay on the device, e.g. the confidence column of mmu_heads_uncertainty_epilogue's
 * scores_out for the image-only / text-only variants of the batch. */
MMU_API int mmu_modality_keep_mask(const float* u, const float* r, const float* score_img,
                                   const float* score_txt, int score_stride, int B, float p_drop,
                                   int mode, int* keep, void* stream);

/* Ragged batch assembly on device: replaces torch pad_sequence(batch_first=True, padding_value=0)
 * in collate_fn_flava (src/dataset.py:216-226).  packed: fp32 [offsets[B], d] rows of the batch's
 * samples back to back; offsets: int32[B+1] (device); out: fp32 (B, max_l, d), tails zero-filled. */
MMU_API int mmu_ragged_pad(const float* packed, const int* offsets, float* out, int B, int max_l, int d,
                   void* stream);

/* bf16 copy of a flat fp32 buffer (the GEMM-operand shadow of the parameters); n elements. */
MMU_API int mmu_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream);

/* LayerNorm, eps 1e-5, biased variance, fp32 statistics (src/model.py:174-180 LayerNorm
 * subclass; :252-253 ln_pre / ln_post).  mean/rstd: fp32[M], saved for the backward. */
MMU_API int mmu_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                      float* mean, float* rstd, int M, int D, void* stream);
/* dx(fp32) = (accumulate ? dx : 0) + dLN(dy); dgamma/dbeta += ; dcolsum (may be NULL) += column
 * sums of the final dx; dx_lp (may be NULL): copy of the final dx in lp_dtype. */
MMU_API int mmu_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* mean,
                      const float* rstd, const float* gamma, float* dx, int accumulate, void* dx_lp,
                      int lp_dtype, float* dgamma, float* dbeta, float* dcolsum, int M, int D,
                      void* stream);

/* Batch-axis multi-head attention (src/model.py:193,205-207: nn.MultiheadAttention with
 * batch_first=False fed (B, L, D), i.e. attention across the mini-batch, per token position).
 * qkv: [B*L, 3D] packed q|k|v; out: [B*L, D].  pos_major = 0: row of (sample b, position l) is
 * b*L + l (the reference's (B, L, .) tensors); pos_major = 1: l*B + b -- the engine's layout: the B
 * rows of one (position, head) problem are adjacent instead of L*3D elements apart, which is what
 * lets these HBM-bound kernels stream (profiles/r02_attention_layout.md).
 * Tensor-core path (dtype MMU_BF16, head_dim % 64 == 0, probs/scores non-NULL): batched tcgen05
 * GEMMs over the L*H (position, head) problems; probs: bf16 [L*H, B, Bp] written by the forward and
 * read by the backward, scores: fp32 scratch, dprobs: bf16 scratch of the same shape (Bp = B
 * rounded up to 8).  Otherwise (fp32, or NULL buffers): fp32 SIMT kernels, B <= 256, which need
 * lse fp32[L*H*B] (forward output) and delta_ws fp32[L*H*B].
 * The forward's `pos_major` argument is a flag word: bit 0 = position-major rows, bit 1 = the
 * probabilities are not needed (eval).  bf16 with head_dim 256 and B <= 128 runs as ONE fused kernel
 * (QK^T -> softmax -> PV on chip, csrc/battn_fused.cu); with bit 1 set `probs` is then not written. */
MMU_API int mmu_batchaxis_attention_fwd(const void* qkv, void* out, float* lse, void* probs,
                                        float* scores, int dtype, int B, int L, int D, int H,
                                        int pos_major, void* stream);
MMU_API int mmu_batchaxis_attention_bwd(const void* qkv, const void* out, const void* dout,
                                        const float* lse, float* delta_ws, const void* probs,
                                        float* scores, void* dprobs, void* dqkv, int dtype, int B,
                                        int L, int D, int H, int pos_major, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused softmax / cross-entropy (+gradient) / accuracy / uncertainty / calibration epilogue.
 * Replaces CrossEntropyLoss in compute_loss (src/model.py:293-304), `acc` (train.py:119-130)
 * and the .cpu().numpy() logits dump + notebook scoring of the robustness scripts
 * (eval_transformer_robustness.py:123-137) with on-device accumulators.
 *   mode 0 (train): one CE row per (sample, head) against labels[n*label_stride + e*label_estride]
 *   mode 1 (eval) : CE on the head-mean logits against labels[n*label_stride]
 *   dlogits (mode 0, may be NULL): (softmax - onehot) * grad_scale, same layout as logits.
 *   pred_out: int32[N,2] = (prediction behind `acc`, argmax of the mean probability) or NULL.
 *   scores_out: fp32[N,4] = (confidence, H_pred, H_exp, MI) or NULL.
 *   accum: device mmu_metric_accum, accumulated with atomics (zero it first); may be NULL. */
typedef struct {
  unsigned long long conf_count[15];   /* confidence histogram, bin = min(floor(conf*15), 14) */
  unsigned long long conf_correct[15];
  unsigned long long hpred_count[32];  /* H_pred / log C    in 32 equal bins */
  unsigned long long mi_count[32];     /* MI / log max(E,2) in 32 equal bins */
  unsigned long long n_samples;
  unsigned long long n_rows;
  unsigned long long n_correct_rows;
  unsigned long long n_correct_prob;
  double conf_sum[15];
  double loss_sum;
  double sum_h_pred, sum_h_exp, sum_mi;
} mmu_metric_accum;

MMU_API int mmu_heads_uncertainty_epilogue(const float* logits, const long long* labels, int label_stride,
                                   int label_estride, int N, int E, int C, int mode,
                                   float grad_scale, float* dlogits, int* pred_out,
                                   float* scores_out, mmu_metric_accum* accum, void* stream);

/* ------------------------------------------------------------------------------------------
 * Post-hoc robustness scoring on device: p(true label) from head-averaged probabilities, the
 * Pearson statistics of experimental vs mean-control delta-p per modality, and per-variant
 * accuracy of the head-mean logits -- what notebooks/utils.py:22-34 and
 * notebooks/food101_robustness.py:24-77 compute on the CPU from the dumped (S, 43, K, C) array.
 * logits: fp32 (V, B, E, C) with V = 3 + 2*n_repeats in the order of
 * eval_transformer_robustness.py:103-121; labels int64 (B); p_true_out: fp32 (B, V) or NULL.
 * The accumulator is += (zero it first; sums all-reduce across ranks). */
typedef struct {
  double sx[2], sy[2], sxx[2], syy[2], sxy[2]; /* [image, text]: x = dp(experiment), y = mean dp(controls) */
  unsigned long long n_samples;
  unsigned long long correct[128];            /* per variant */
} mmu_posthoc_accum;
MMU_API int mmu_posthoc_scoring(const float* logits, const long long* labels, int V, int B, int E, int C,
                        int n_repeats, float* p_true_out, mmu_posthoc_accum* acc, void* stream);

/* ------------------------------------------------------------------------------------------
 * Rank statistics of the post-hoc analysis, as exact integer pair counts.
 * mmu_pair_concordance: for each of `batch` problems b, over all unordered pairs {i, j} of
 *   xb = x + b*x_batch_stride, yb = y + b*y_batch_stride (fp32, n elements each; a stride of 0
 *   shares one vector between the problems), counts[4*b + 0..3] = { concordant, discordant,
 *   tied in x, tied in y } (both tie counts include the jointly tied pairs; joint =
 *   conc + disc + tx + ty - n(n-1)/2).  counts is overwritten.  NaNs compare as ties.
 *   - AUROC (sklearn.metrics.roc_auc_score as called at src/framework.py:195-198 and
 *     notebooks/hatefulmeme_robustness.py:22-41): x = labels as 0/1 floats, y = scores;
 *     AUROC = (conc + (ty - joint)/2) / (conc + disc + ty - joint).
 *   - Kendall tau-b (scipy.stats.kendalltau as called at notebooks/analysis_round_1.py:87-90):
 *     (conc - disc) / sqrt((T - tx)(T - ty)), T = n(n-1)/2.
 * mmu_top_truncate: notebooks/analysis_round_1.py:74-85 `trunk_pred_top` on device -- per row of
 *   pred (N, C): threshold = the top-th largest value of the original row, the true class zeroed
 *   first when mute_true (labels int64 (N), may be NULL otherwise), entries below the threshold
 *   zeroed; out (N, C). */
MMU_API int mmu_pair_concordance(const float* x, const float* y, long long n, int batch,
                         long long x_batch_stride, long long y_batch_stride,
                         unsigned long long* counts, void* stream);
MMU_API int mmu_top_truncate(const float* pred, const long long* labels, int N, int C, int top,
                     int mute_true, float* out, void* stream);

/* Fused AdamW over a flat fp32 buffer: torch.optim.AdamW as configured in train.py:196-202.
 * `step` is the 1-based step count; grad_scale multiplies g first (1/world_size for DDP);
 * p_bf16 (may be NULL) receives the bf16 shadow of the updated parameters.  n % 4 == 0. */
MMU_API int mmu_adamw_flat_step(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                        float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * FLAVA-fusion engine: FlavaFusionTransfomer / FlavaFusionTransfomerwithCLSToken forward and
 * backward (src/model.py:225-374) over caller-owned flat buffers. */
typedef struct {
  int B, l_img, l_txt, d_img, d_txt, D, n_head, n_layers, E, C;
  int avg_pool;   /* kwargs["avg_pool"], src/model.py:256,281-284 */
  int cls_token;  /* 1: the CLS-token variant, src/model.py:306-361 */
  int precision;  /* MMU_F32 or MMU_BF16 */
  int max_variants; /* capacity for packed-variant evaluation (0 or 1: single variant) */
  int group_pool;   /* > 0: MIMOTransfomer head wiring (src/model.py:148-153): head e = mean of
                       token positions [e*group_pool, (e+1)*group_pool); with d_txt == 0 the text
                       projection does not exist (single-modality model) */
} mmu_flava_config;

typedef struct {
  char name[96];      /* reference state_dict key, e.g. "mm_encoder.resblocks.0.attn.in_proj_weight" */
  long long offset;   /* element offset into the flat parameter / gradient buffers */
  long long numel;
  int rows, cols;     /* cols == 0: vector */
  int stage;          /* backward stage that completes this gradient */
} mmu_param_entry;

MMU_API long long mmu_flava_param_count(const mmu_flava_config* cfg);
MMU_API int mmu_flava_param_table(const mmu_flava_config* cfg, mmu_param_entry* out /* host */, int max);
MMU_API long long mmu_flava_workspace_bytes(const mmu_flava_config* cfg, int training);
MMU_API int mmu_flava_num_stages(const mmu_flava_config* cfg);

typedef struct {
  const void* img;     /* (B, l_img, d_img) fp32 (bf16 with src_bf16) or NULL (modality absent) */
  const void* txt;     /* (B, l_txt, d_txt) fp32 (bf16 with src_bf16) or NULL */
  const int* idx_img;  /* int32[n_img] token subset or NULL (first n_img tokens) */
  const int* idx_txt;
  int n_img, n_txt;    /* tokens fed to the model (<= l_img / l_txt) */
  const int* keep;     /* int32[B,2] modality keep mask (0: zero-fill) or NULL */
  const void* params_bf16; /* optional bf16 copy of `params` kept fresh by the caller (the fused
                              AdamW writes it); NULL: cast from the fp32 master every forward */
  /* Packed-variant evaluation (replaces the 43 forwards per batch of
   * eval_transformer_robustness.py:99-125 by one): attention runs over the batch axis and every
   * other op is row-wise, so token positions never interact and V token-subset variants can be
   * concatenated along L.  idx_img / idx_txt then hold the concatenated per-variant index lists
   * (n_img / n_txt are the totals; cfg->l_img / l_txt act as capacities), src_l_* give the token
   * counts of the source tensors, var_segments the rows [begin, end) feeding head e of variant v. */
  int src_l_img, src_l_txt;  /* 0: cfg->l_img / cfg->l_txt */
  int n_variants;            /* 0 or 1: ordinary forward */
  const int* var_segments;   /* device int32[n_variants][E][2] */
  /* nn.Dropout(drop) between c_fc and QuickGELU (src/model.py:195-201; training forward and its
   * backward only): element (row r, column c) of layer i's [B*L, 4D] pre-activation is kept iff
   * hash(seed, site = i, r * 4D + c) >= floor(p * 2^32) and scaled by 1 / (1 - p) -- the mask
   * function of csrc/dropout.cuh, restated in oracle/dropout.py.  drop_p == 0: no dropout. */
  float drop_p;
  int src_bf16;  /* 1: img / txt point to bf16 tensors of the same shapes (bf16 host staging: half the
                    host->device bytes; the bf16 engine rounds its inputs to bf16 in the stem anyway, so
                    the results are bit-identical to feeding the fp32 values that round to them) */
  unsigned long long drop_seed;
} mmu_flava_inputs;

/* logits: fp32 (B, E, C), or (n_variants, B, E, C) for a packed-variant forward (eval only).
 * training != 0 keeps the activations the backward needs. */
MMU_API int mmu_flava_forward(const mmu_flava_config* cfg, const float* params, const mmu_flava_inputs* in,
                      void* workspace, long long workspace_bytes, int training, float* logits,
                      void* stream);
/* grads (flat fp32, same layout as params) are ACCUMULATED (+=), like autograd's .grad.
 * Stages [stage_begin, stage_end): 0 heads+ln_post, 1..n_layers the blocks in reverse order,
 * n_layers+1 the stem; running them in separate calls lets the caller overlap a bucketed
 * gradient all-reduce with the rest of the backward. */
MMU_API int mmu_flava_backward(const mmu_flava_config* cfg, const float* params, const mmu_flava_inputs* in,
                       void* workspace, long long workspace_bytes, const float* dlogits,
                       float* grads, int stage_begin, int stage_end, void* stream);

/* ------------------------------------------------------------------------------------------
 * MIMOResNet engine: the four-view FashionMNIST ResNet of train_fashionmnist.py (src/model.py:17-100
 * ResNet / MultiHeadFC / MIMOResNet, src/layers.py:7-38 BasicBlock), fp32.  Same conventions as
 * the FLAVA engine: flat fp32 parameter / gradient buffers described by a table of reference
 * state_dict names, a flat buffer for the BatchNorm running statistics (second table), a
 * caller-owned workspace.  x: fp32 (B, cin, 14, 14); logits: fp32 (B, E, C).  training != 0 uses
 * batch statistics and moves the running statistics (momentum 0.1, unbiased variance). */
typedef struct {
  int B, cin, H, W, E, C;
} mmu_resnet_config;
MMU_API long long mmu_resnet_param_count(const mmu_resnet_config* cfg);
MMU_API long long mmu_resnet_stat_count(const mmu_resnet_config* cfg);
MMU_API int mmu_resnet_param_table(const mmu_resnet_config* cfg, mmu_param_entry* out /* host */, int max);
MMU_API int mmu_resnet_stat_table(const mmu_resnet_config* cfg, mmu_param_entry* out /* host */, int max);
MMU_API long long mmu_resnet_workspace_bytes(const mmu_resnet_config* cfg, int training);
/* params_bf16: optional bf16 copy of params (mmu_cast_f32_to_bf16): the convolutions then run on
 * the tcgen05 tensor-core GEMM (bf16 operands, fp32 accumulation); NULL: fp32 parity path. */
MMU_API int mmu_resnet_forward(const mmu_resnet_config* cfg, const float* params, const void* params_bf16,
                       float* stats, const float* x, void* workspace, long long workspace_bytes,
                       int training, float* logits, void* stream);
MMU_API int mmu_resnet_backward(const mmu_resnet_config* cfg, const float* params, const void* params_bf16,
                        float* stats, const float* x, void* workspace, long long workspace_bytes,
                        const float* dlogits, float* grads, void* stream);

/* ------------------------------------------------------------------------------------------
 * MMBT engine: reference src/mmbt.py:47-262 (ImageBertEmbeddings, MultimodalBertEncoder.forward /
 * forward_img_only / forward_txt_only / forward_control, MultimodalBertClf) from the pooled image
 * tokens onward.  The BERT arithmetic is that of the reference's un-vendored, unpinned dependency
 * pytorch_pretrained_bert (call sites src/mmbt.py:13,90-96,124-128): post-LN blocks, LayerNorm
 * eps 1e-12, erf-GELU, additive mask (1 - m) * -10000 (src/mmbt.py:103-107).  Flat fp32
 * parameter / gradient buffers described by a table of reference state_dict names; the query /
 * key / value tensors of a layer are adjacent (one [3D, D] GEMM).  logits: fp32 (B, C). */
typedef struct {
  int B, S_txt, n_img, d_img, D, n_head, n_layers, d_ff, vocab, max_pos, n_types, C;
  int cls_id, sep_id; /* args.vocab.stoi["[CLS]"], ["[SEP]"] (src/mmbt.py:62-66) */
  int precision;      /* 0 fp32 (parity path), 1 bf16 operands on tcgen05 */
  int max_seq;        /* workspace capacity in sequence positions; 0 = n_img + 2 + S_txt */
  /* Dropout (training forward + its backward only; masks = csrc/dropout.cuh, oracle/dropout.py):
   * drop_hidden = BertConfig.hidden_dropout_prob (text embeddings; attention-output and FFN-output
   * dense layers before the residual add), drop_attn = attention_probs_dropout_prob (softmax(QK^T)
   * before P V; applied inside the fused attention kernels, mask regenerated in the backward),
   * drop_img = args.dropout of ImageBertEmbeddings (src/mmbt.py:56,82).  Sites: 0 embeddings,
   * 4l+1 attention probabilities, 4l+2 attention output, 4l+3 FFN output of layer l; element
   * counters: row * D + column, resp. ((b * H + h) * S + query) * S + key. */
  float drop_hidden, drop_attn, drop_img;
  int drop_reserved;
} mmu_mmbt_config;
typedef struct {
  const long long* txt;     /* (B, S_txt) token ids */
  const long long* mask;    /* (B, S_txt) 1 = attend */
  const long long* segment; /* (B, S_txt) token types */
  const float* img;         /* (B, n_img, d_img): ImageEncoder output, src/mmbt.py:40-45 */
  const int* indices;       /* device int32[n_sel] positions of [CLS img.. SEP | text..] that enter the
                               encoder; NULL = all (forward).  img_only = first n_img + 2; txt_only =
                               {0} + text; forward_control = {0} + sorted sample (src/mmbt.py:198-201) */
  int n_sel;
  int indices_per_sample;   /* 1: indices is int32[B][n_sel], one list per sample (equally long robustness
                               variants of a batch packed along the batch axis into one forward) */
  const void* params_bf16;  /* optional bf16 shadow of params */
  float* dimg;              /* backward: d loss / d img, or NULL */
  unsigned long long drop_seed; /* seed of this forward's dropout masks; pass the same to the backward */
} mmu_mmbt_inputs;
MMU_API long long mmu_mmbt_param_count(const mmu_mmbt_config* cfg);
MMU_API int mmu_mmbt_param_table(const mmu_mmbt_config* cfg, mmu_param_entry* out /* host */, int max);
MMU_API long long mmu_mmbt_workspace_bytes(const mmu_mmbt_config* cfg, int training);
MMU_API int mmu_mmbt_forward(const mmu_mmbt_config* cfg, const float* params, const mmu_mmbt_inputs* in,
                     void* workspace, long long workspace_bytes, int training, float* logits,
                     void* stream);
MMU_API int mmu_mmbt_backward(const mmu_mmbt_config* cfg, const float* params, const mmu_mmbt_inputs* in,
                      void* workspace, long long workspace_bytes, const float* dlogits, float* grads,
                      void* stream);
/* Sequence-axis self-attention of the MMBT path's BERT encoder (pytorch_pretrained_bert
 * BertSelfAttention as called from src/mmbt.py:124-128, additive mask of :103-107).
 * qkv (dtype) [B*S, 3D] packed q|k|v; addmask fp32 [B, S]; out (dtype) [B*S, D].
 * dtype 1 (bf16, head_dim % 64 == 0): tcgen05; probs bf16 [B*H, S, Sp] (Sp = S rounded up to 8),
 * scores fp32 / dprobs bf16 scratch of that shape.  dtype 0: fp32 SIMT parity path, probs fp32
 * [B*H, S, S].  flags bit 0: keep the probabilities for the backward; bit 1: do NOT use the fused
 * one-kernel forward (head_dim 64, S <= 512) -- A/B and test switch. */
MMU_API int mmu_seq_attention_fwd(const void* qkv, const float* addmask, void* out, void* probs,
                          float* scores, int dtype, int B, int S, int D, int H, int flags, void* stream);
MMU_API int mmu_seq_attention_bwd(const void* qkv, const void* dout, const void* probs, float* scores,
                          void* dprobs, void* dqkv, int dtype, int B, int S, int D, int H, void* stream);

/* MMBT image encoder (reference src/mmbt.py:15-45 ImageEncoder): torchvision Bottleneck ResNet trunk
 * (resnet152: layers {3, 8, 36, 3}, children()[:-2]) + AdaptiveAvg/MaxPool2d to num_image_embeds
 * cells, flattened to (B, N, 2048) tokens.  Parameter names are the nn.Sequential keys
 * ("model.0.weight", "model.1.weight", "model.4.0.conv1.weight", ...); BatchNorm running statistics
 * live in a second flat buffer (second table), updated in training mode as torch does. */
typedef struct {
  int B, H;             /* images, square input size */
  int layers[4];        /* {3, 8, 36, 3} */
  int width_per_group;  /* 64 */
  int pool_h, pool_w;   /* src/mmbt.py:28-37 */
  int pool_max;         /* img_embed_pool_type != "avg" */
} mmu_imgenc_config;
MMU_API long long mmu_imgenc_param_count(const mmu_imgenc_config* cfg);
MMU_API long long mmu_imgenc_stat_count(const mmu_imgenc_config* cfg);
MMU_API int mmu_imgenc_param_table(const mmu_imgenc_config* cfg, mmu_param_entry* out /* host */, int max);
MMU_API int mmu_imgenc_stat_table(const mmu_imgenc_config* cfg, mmu_param_entry* out /* host */, int max);
MMU_API long long mmu_imgenc_workspace_bytes(const mmu_imgenc_config* cfg, int training);
MMU_API int mmu_imgenc_forward(const mmu_imgenc_config* cfg, const float* params, const void* params_bf16,
                       float* stats, const float* x /* (B,3,H,H) */, void* workspace,
                       long long workspace_bytes, int training, float* tokens, void* stream);
MMU_API int mmu_imgenc_backward(const mmu_imgenc_config* cfg, const float* params, const void* params_bf16,
                        float* stats, const float* x, void* workspace, long long workspace_bytes,
                        const float* dtokens, float* grads, void* stream);
/* BertAdam step over a flat buffer (pytorch_pretrained_bert.optimization.BertAdam as configured by
 * train.py:136-147): per-tensor clip_grad_norm_(max_grad_norm) applied to g in place, m/v without
 * bias correction, update = m / (sqrt(v) + eps) + decay * p, p -= lr * update.  segs: device
 * int64[n_seg][2] (offset, numel); seg_hyper: device fp32[n_seg][2] = (weight decay, scheduled lr --
 * the reference keeps state['step'] per tensor); norms: device fp32[n_seg] scratch. */
MMU_API int mmu_bertadam_flat_step(float* p, float* g, float* m, float* v, void* p_bf16,
                           const long long* segs, const float* seg_hyper, float* norms, int n_seg,
                           long long max_seg_numel, float b1, float b2, float eps,
                           float max_grad_norm, float grad_scale /* applied to g first: 1/world */,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMU_B200_H_ */
