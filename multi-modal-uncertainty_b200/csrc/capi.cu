// extern "C" surface of libmmu_b200.so: thin, exception-free adapters over the C++ launchers.
#include <cstring>

#include "../../include/mmu_b200.h"
#include "common.h"
#include "engine.h"
#include "gemm_api.h"
#include "kernels.h"
#include "mmbt.h"
#include "resnet.h"

using namespace mmu;

static_assert(sizeof(mmu_metric_accum) == sizeof(MetricAccum), "metric accumulator layout");
static_assert(sizeof(mmu_flava_config) == sizeof(FlavaConfig), "config layout");
static_assert(sizeof(mmu_resnet_config) == sizeof(ResNetConfig), "resnet config layout");
static_assert(sizeof(mmu_posthoc_accum) == sizeof(PosthocAccum), "post-hoc accumulator layout");
static_assert(sizeof(mmu_param_entry) == sizeof(ParamEntry), "param entry layout");
static_assert(sizeof(mmu_mmbt_config) == sizeof(MmbtConfig), "mmbt config layout");
static_assert(sizeof(mmu_imgenc_config) == sizeof(ImgEncConfig), "image encoder config layout");
static_assert(sizeof(mmu_mmbt_inputs) == sizeof(MmbtInputs), "mmbt inputs layout");
static_assert(sizeof(mmu_flava_inputs) == sizeof(FlavaInputs), "inputs layout");

namespace {
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
inline FlavaConfig cfg_of(const mmu_flava_config* c) {
  FlavaConfig r;
  std::memcpy(&r, c, sizeof(r));
  return r;
}
inline FlavaInputs in_of(const mmu_flava_inputs* i) {
  FlavaInputs r;
  std::memcpy(&r, i, sizeof(r));
  return r;
}
}  // namespace

extern "C" {

const char* mmu_version(void) { return "mmu_b200 0.1 (sm_100a)"; }

long long mmu_launch_count(void) { return launch_count(); }

int mmu_set_gemm_sm_limit(int n) {
  set_gemm_sm_limit(n);
  return gemm_sms();
}

int mmu_struct_size(int which) {
  switch (which) {
    case 0: return static_cast<int>(sizeof(mmu_flava_config));
    case 1: return static_cast<int>(sizeof(mmu_flava_inputs));
    case 2: return static_cast<int>(sizeof(mmu_gemm_epilogue));
    case 3: return static_cast<int>(sizeof(mmu_metric_accum));
    case 4: return static_cast<int>(sizeof(mmu_param_entry));
    case 5: return static_cast<int>(sizeof(mmu_posthoc_accum));
    case 6: return static_cast<int>(sizeof(mmu_mmbt_config));
    case 7: return static_cast<int>(sizeof(mmu_mmbt_inputs));
    case 8: return static_cast<int>(sizeof(mmu_imgenc_config));
    default: return -1;
  }
}

const char* mmu_error_string(int code) {
  switch (code) {
    case MMU_OK: return "ok";
    case MMU_ERR_SHAPE: return "unsupported or inconsistent shape";
    case MMU_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case MMU_ERR_DRIVER: return "CUDA driver entry point unavailable (no GPU?)";
    case MMU_ERR_TMAP: return "tensor-map encoding failed";
    case MMU_ERR_CUDA: return "CUDA launch failed";
    case MMU_ERR_ARG: return "bad argument";
    case MMU_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

int mmu_gemm(int dtype, const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
             int b_mn_major, int M, int N, int K, int splits, const mmu_gemm_epilogue* epi,
             void* stream) {
  if (A == nullptr || B == nullptr || epi == nullptr || epi->out == nullptr && epi->out2 == nullptr)
    return MMU_ERR_ARG;
  GemmProblem p{M, N, K, a_mn_major, b_mn_major, splits};
  GemmEpilogue e{};
  e.mode = epi->mode;
  e.out_bf16 = (dtype == MMU_BF16 && epi->out_lp) ? 1 : 0;
  e.out = epi->out;
  e.out2 = epi->out2;
  e.bias = epi->bias;
  e.aux = epi->aux;
  e.ld_out = epi->ld_out;
  e.ld_out2 = epi->ld_out2;
  e.ld_aux = epi->ld_aux;
  e.seg_len = epi->seg_len;
  e.seg_stride = epi->seg_stride;
  e.seg_off = epi->seg_off;
  e.alpha = epi->alpha;
  e.drop = dropout::make_site(epi->drop_p, epi->drop_seed, static_cast<unsigned int>(epi->drop_site));
  e.ln_stats = epi->ln_stats;
  e.ln_cw = epi->ln_cw;
  e.ln_nt = epi->ln_nt;
  e.ln_inv_d = epi->ln_inv_d;
  e.ln_eps = epi->ln_eps;
  e.stats_out = epi->stats_out;
  e.stats_nt = epi->stats_nt;
  if (dtype != MMU_BF16 && (e.ln_stats != nullptr || e.mode == EPI_RESID_LN)) return MMU_ERR_ARG;
  if (epi->drop_p < 0.f || epi->drop_p >= 1.f) return MMU_ERR_ARG;
  if (dtype == MMU_BF16) return gemm_bf16_launch(A, lda, B, ldb, p, e, S(stream));
  if (dtype == MMU_F32)
    return gemm_f32_launch(static_cast<const float*>(A), lda, static_cast<const float*>(B), ldb, p,
                           e, S(stream));
  return MMU_ERR_ARG;
}

int mmu_ln_fold_weights(const float* W, const float* gamma, const float* beta, const float* bias,
                        void* Wf_bf16, float* cw, float* bf, int N, int K, void* stream) {
  if (W == nullptr || gamma == nullptr || beta == nullptr || Wf_bf16 == nullptr || cw == nullptr ||
      bf == nullptr)
    return MMU_ERR_ARG;
  return ln_fold_weights(W, gamma, beta, bias, Wf_bf16, cw, bf, N, nullptr, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr, 0, K, S(stream));
}

int mmu_layernorm_raw_stats(const float* x, const float* gamma, const float* beta, float* y,
                            void* yraw_bf16, float* stats, int nt, int M, int D, void* stream) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr || yraw_bf16 == nullptr ||
      stats == nullptr)
    return MMU_ERR_ARG;
  return layernorm_raw_stats_fwd(x, gamma, beta, y, yraw_bf16, stats, nt, M, D, S(stream));
}

int mmu_mask_gather_tokens(const float* src, void* dst, int dst_dtype, int B, int l_src, int d,
                           const int* idx, int n_sel, const int* keep, int modality, int pos_major,
                           void* stream) {
  if (src == nullptr || dst == nullptr) return MMU_ERR_ARG;
  return cast_gather(src, dst, dst_dtype, B, l_src, d, idx, n_sel, keep, modality, S(stream), pos_major != 0);
}

int mmu_modality_keep_mask(const float* u, const float* r, const float* score_img,
                           const float* score_txt, int score_stride, int B, float p_drop, int mode,
                           int* keep, void* stream) {
  return modality_keep_mask(u, r, score_img, score_txt, score_stride, B, p_drop, mode, keep, S(stream));
}

int mmu_ragged_pad(const float* packed, const int* offsets, float* out, int B, int max_l, int d,
                   void* stream) {
  if (packed == nullptr || offsets == nullptr || out == nullptr) return MMU_ERR_ARG;
  return ragged_pad(packed, offsets, out, B, max_l, d, S(stream));
}

int mmu_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream) {
  if (src == nullptr || dst == nullptr) return MMU_ERR_ARG;
  return cast_f32_to_bf16(src, dst, n, S(stream));
}

int mmu_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                      float* mean, float* rstd, int M, int D, void* stream) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || y == nullptr) return MMU_ERR_ARG;
  return layernorm_fwd(x, gamma, beta, y, y_dtype, mean, rstd, M, D, S(stream));
}

int mmu_layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* mean,
                      const float* rstd, const float* gamma, float* dx, int accumulate, void* dx_lp,
                      int lp_dtype, float* dgamma, float* dbeta, float* dcolsum, int M, int D,
                      void* stream) {
  if (dy == nullptr || x == nullptr || mean == nullptr || rstd == nullptr || gamma == nullptr ||
      dx == nullptr || dgamma == nullptr || dbeta == nullptr)
    return MMU_ERR_ARG;
  return layernorm_bwd(dy, dy_dtype, x, mean, rstd, gamma, dx, accumulate, dx_lp, lp_dtype, dgamma,
                       dbeta, dcolsum, M, D, S(stream));
}

int mmu_batchaxis_attention_fwd(const void* qkv, void* out, float* lse, void* probs, float* scores,
                                int dtype, int B, int L, int D, int H, int pos_major, void* stream) {
  if (qkv == nullptr || out == nullptr) return MMU_ERR_ARG;
  return attention_fwd(qkv, out, lse, probs, scores, dtype, B, L, D, H, S(stream), (pos_major & 1) != 0,
                       (pos_major & 2) == 0);
}

int mmu_batchaxis_attention_bwd(const void* qkv, const void* out, const void* dout,
                                const float* lse, float* delta_ws, const void* probs, float* scores,
                                void* dprobs, void* dqkv, int dtype, int B, int L, int D, int H,
                                int pos_major, void* stream) {
  if (qkv == nullptr || out == nullptr || dout == nullptr || dqkv == nullptr) return MMU_ERR_ARG;
  return attention_bwd(qkv, out, dout, lse, delta_ws, probs, scores, dprobs, dqkv, dtype, B, L, D,
                       H, S(stream), pos_major != 0);
}

int mmu_heads_uncertainty_epilogue(const float* logits, const long long* labels, int label_stride,
                                   int label_estride, int N, int E, int C, int mode,
                                   float grad_scale, float* dlogits, int* pred_out,
                                   float* scores_out, mmu_metric_accum* accum, void* stream) {
  if (logits == nullptr || labels == nullptr) return MMU_ERR_ARG;
  return ce_uncertainty(logits, labels, label_stride, label_estride, N, E, C, mode, grad_scale,
                        dlogits, pred_out, scores_out, reinterpret_cast<MetricAccum*>(accum),
                        S(stream));
}

int mmu_adamw_flat_step(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                        float grad_scale, void* stream) {
  if (p == nullptr || g == nullptr || m == nullptr || v == nullptr) return MMU_ERR_ARG;
  return adamw_flat(p, g, m, v, p_bf16, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                    S(stream));
}

long long mmu_flava_param_count(const mmu_flava_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : flava_param_count(cfg_of(cfg));
}
int mmu_flava_param_table(const mmu_flava_config* cfg, mmu_param_entry* out, int max) {
  return cfg == nullptr ? MMU_ERR_ARG
                        : flava_param_table(cfg_of(cfg), reinterpret_cast<ParamEntry*>(out), max);
}
long long mmu_flava_workspace_bytes(const mmu_flava_config* cfg, int training) {
  return cfg == nullptr ? MMU_ERR_ARG : flava_workspace_bytes(cfg_of(cfg), training);
}
int mmu_flava_num_stages(const mmu_flava_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : flava_num_stages(cfg_of(cfg));
}

int mmu_posthoc_scoring(const float* logits, const long long* labels, int V, int B, int E, int C,
                        int n_repeats, float* p_true_out, mmu_posthoc_accum* acc, void* stream) {
  if (logits == nullptr || labels == nullptr || acc == nullptr) return MMU_ERR_ARG;
  return posthoc_scoring(logits, labels, V, B, E, C, n_repeats, p_true_out,
                         reinterpret_cast<PosthocAccum*>(acc), S(stream));
}

int mmu_pair_concordance(const float* x, const float* y, long long n, int batch,
                         long long x_batch_stride, long long y_batch_stride,
                         unsigned long long* counts, void* stream) {
  if (x == nullptr || y == nullptr || counts == nullptr) return MMU_ERR_ARG;
  return pair_concordance(x, y, n, batch, x_batch_stride, y_batch_stride, counts, S(stream));
}

int mmu_top_truncate(const float* pred, const long long* labels, int N, int C, int top, int mute_true,
                     float* out, void* stream) {
  if (pred == nullptr || out == nullptr || (mute_true && labels == nullptr)) return MMU_ERR_ARG;
  return top_truncate(pred, labels, N, C, top, mute_true, out, S(stream));
}

int mmu_flava_forward(const mmu_flava_config* cfg, const float* params, const mmu_flava_inputs* in,
                      void* workspace, long long workspace_bytes, int training, float* logits,
                      void* stream) {
  if (cfg == nullptr || in == nullptr) return MMU_ERR_ARG;
  return flava_forward(cfg_of(cfg), params, in_of(in), workspace, workspace_bytes, training, logits,
                       S(stream));
}

int mmu_flava_backward(const mmu_flava_config* cfg, const float* params, const mmu_flava_inputs* in,
                       void* workspace, long long workspace_bytes, const float* dlogits,
                       float* grads, int stage_begin, int stage_end, void* stream) {
  if (cfg == nullptr || in == nullptr) return MMU_ERR_ARG;
  return flava_backward(cfg_of(cfg), params, in_of(in), workspace, workspace_bytes, dlogits, grads,
                        stage_begin, stage_end, S(stream));
}

namespace {
inline ResNetConfig rcfg_of(const mmu_resnet_config* c) {
  ResNetConfig r;
  std::memcpy(&r, c, sizeof(r));
  return r;
}
}  // namespace

long long mmu_resnet_param_count(const mmu_resnet_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : resnet_param_count(rcfg_of(cfg));
}
long long mmu_resnet_stat_count(const mmu_resnet_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : resnet_stat_count(rcfg_of(cfg));
}
int mmu_resnet_param_table(const mmu_resnet_config* cfg, mmu_param_entry* out, int max) {
  return cfg == nullptr ? MMU_ERR_ARG
                        : resnet_param_table(rcfg_of(cfg), reinterpret_cast<ParamEntry*>(out), max);
}
int mmu_resnet_stat_table(const mmu_resnet_config* cfg, mmu_param_entry* out, int max) {
  return cfg == nullptr ? MMU_ERR_ARG
                        : resnet_stat_table(rcfg_of(cfg), reinterpret_cast<ParamEntry*>(out), max);
}
long long mmu_resnet_workspace_bytes(const mmu_resnet_config* cfg, int training) {
  return cfg == nullptr ? MMU_ERR_ARG : resnet_workspace_bytes(rcfg_of(cfg), training);
}
int mmu_resnet_forward(const mmu_resnet_config* cfg, const float* params, const void* params_bf16,
                       float* stats, const float* x, void* workspace, long long workspace_bytes,
                       int training, float* logits, void* stream) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return resnet_forward(rcfg_of(cfg), params, params_bf16, stats, x, workspace, workspace_bytes,
                        training, logits, S(stream));
}
int mmu_resnet_backward(const mmu_resnet_config* cfg, const float* params, const void* params_bf16,
                        float* stats, const float* x, void* workspace, long long workspace_bytes,
                        const float* dlogits, float* grads, void* stream) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return resnet_backward(rcfg_of(cfg), params, params_bf16, stats, x, workspace, workspace_bytes,
                         dlogits, grads, S(stream));
}

namespace {
inline MmbtConfig mcfg_of(const mmu_mmbt_config* c) {
  MmbtConfig r;
  std::memcpy(&r, c, sizeof(r));
  return r;
}
inline MmbtInputs min_of(const mmu_mmbt_inputs* i) {
  MmbtInputs r;
  std::memcpy(&r, i, sizeof(r));
  return r;
}
}  // namespace

long long mmu_mmbt_param_count(const mmu_mmbt_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : mmbt_param_count(mcfg_of(cfg));
}
int mmu_mmbt_param_table(const mmu_mmbt_config* cfg, mmu_param_entry* out, int max) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return mmbt_param_table(mcfg_of(cfg), reinterpret_cast<ParamEntry*>(out), max);
}
long long mmu_mmbt_workspace_bytes(const mmu_mmbt_config* cfg, int training) {
  return cfg == nullptr ? MMU_ERR_ARG : mmbt_workspace_bytes(mcfg_of(cfg), training);
}
int mmu_mmbt_forward(const mmu_mmbt_config* cfg, const float* params, const mmu_mmbt_inputs* in,
                     void* workspace, long long workspace_bytes, int training, float* logits,
                     void* stream) {
  if (cfg == nullptr || in == nullptr) return MMU_ERR_ARG;
  return mmbt_forward(mcfg_of(cfg), params, min_of(in), workspace, workspace_bytes, training, logits,
                      S(stream));
}
int mmu_mmbt_backward(const mmu_mmbt_config* cfg, const float* params, const mmu_mmbt_inputs* in,
                      void* workspace, long long workspace_bytes, const float* dlogits, float* grads,
                      void* stream) {
  if (cfg == nullptr || in == nullptr) return MMU_ERR_ARG;
  return mmbt_backward(mcfg_of(cfg), params, min_of(in), workspace, workspace_bytes, dlogits, grads,
                       S(stream));
}
int mmu_bertadam_flat_step(float* p, float* g, float* m, float* v, void* p_bf16, const long long* segs,
                           const float* seg_hyper, float* norms, int n_seg, long long max_seg_numel,
                           float b1, float b2, float eps, float max_grad_norm, float grad_scale,
                           void* stream) {
  return bertadam_flat(p, g, m, v, p_bf16, segs, seg_hyper, norms, n_seg, max_seg_numel, b1, b2, eps,
                       max_grad_norm, grad_scale, S(stream));
}

namespace {
inline ImgEncConfig icfg_of(const mmu_imgenc_config* c) {
  ImgEncConfig r;
  std::memcpy(&r, c, sizeof(r));
  return r;
}
}  // namespace
long long mmu_imgenc_param_count(const mmu_imgenc_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : imgenc_param_count(icfg_of(cfg));
}
long long mmu_imgenc_stat_count(const mmu_imgenc_config* cfg) {
  return cfg == nullptr ? MMU_ERR_ARG : imgenc_stat_count(icfg_of(cfg));
}
int mmu_imgenc_param_table(const mmu_imgenc_config* cfg, mmu_param_entry* out, int max) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return imgenc_param_table(icfg_of(cfg), reinterpret_cast<ParamEntry*>(out), max);
}
int mmu_imgenc_stat_table(const mmu_imgenc_config* cfg, mmu_param_entry* out, int max) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return imgenc_stat_table(icfg_of(cfg), reinterpret_cast<ParamEntry*>(out), max);
}
long long mmu_imgenc_workspace_bytes(const mmu_imgenc_config* cfg, int training) {
  return cfg == nullptr ? MMU_ERR_ARG : imgenc_workspace_bytes(icfg_of(cfg), training);
}
int mmu_imgenc_forward(const mmu_imgenc_config* cfg, const float* params, const void* params_bf16,
                       float* stats, const float* x, void* workspace, long long workspace_bytes,
                       int training, float* tokens, void* stream) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return imgenc_forward(icfg_of(cfg), params, params_bf16, stats, x, workspace, workspace_bytes, training,
                        tokens, S(stream));
}
int mmu_imgenc_backward(const mmu_imgenc_config* cfg, const float* params, const void* params_bf16,
                        float* stats, const float* x, void* workspace, long long workspace_bytes,
                        const float* dtokens, float* grads, void* stream) {
  if (cfg == nullptr) return MMU_ERR_ARG;
  return imgenc_backward(icfg_of(cfg), params, params_bf16, stats, x, workspace, workspace_bytes, dtokens,
                         grads, S(stream));
}

int mmu_seq_attention_fwd(const void* qkv, const float* addmask, void* out, void* probs, float* scores,
                          int dtype, int B, int seq, int D, int H, int flags, void* stream) {
  return seq_attention_fwd(qkv, addmask, out, probs, scores, dtype, B, seq, D, H, S(stream), flags & 1,
                           (flags & 2) ? 0 : 1);
}
int mmu_seq_attention_bwd(const void* qkv, const void* dout, const void* probs, float* scores, void* dprobs,
                          void* dqkv, int dtype, int B, int seq, int D, int H, void* stream) {
  return seq_attention_bwd(qkv, dout, probs, scores, dprobs, dqkv, dtype, B, seq, D, H, S(stream));
}

}  // extern "C"
