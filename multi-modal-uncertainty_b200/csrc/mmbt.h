// MMBT engine: the multimodal bitransformer of reference src/mmbt.py (MultimodalBertEncoder /
// MultimodalBertClf, :86-262) from the pooled image tokens onward -- ImageBertEmbeddings, BERT text
// embeddings, the BERT encoder (post-LN blocks, sequence-axis attention with the additive mask of
// :103-107), pooler and classifier -- forward and backward over caller-owned flat buffers.
// The BERT arithmetic itself lives in the reference's un-vendored dependency
// `pytorch_pretrained_bert` (version unpinned; call sites src/mmbt.py:13,90-96,124-128): this
// engine follows that package's published BertEmbeddings / BertLayer / BertPooler definitions
// (LayerNorm eps 1e-12 inside the square root, erf-GELU, 1/sqrt(head_dim) scaling).
// No allocation, no synchronisation, no exceptions; everything is enqueued on the caller's stream.
#pragma once
#include <cuda_runtime.h>

#include "engine.h"  // ParamEntry, Precision

namespace mmu {

struct MmbtConfig {
  int B;         // samples
  int S_txt;     // text tokens per sample (padded length of this batch)
  int n_img;     // args.num_image_embeds
  int d_img;     // args.img_hidden_sz (2048)
  int D;         // args.hidden_sz (768); must be a multiple of 64
  int n_head;    // 12
  int n_layers;  // 12
  int d_ff;      // 3072
  int vocab;     // 30522
  int max_pos;   // 512
  int n_types;   // 2
  int C;         // args.n_classes
  int cls_id;    // args.vocab.stoi["[CLS]"]
  int sep_id;    // args.vocab.stoi["[SEP]"]
  int precision; // Precision
  int max_seq;   // workspace capacity in sequence positions; 0 = n_img + 2 + S_txt.  A caller that
                 // only ever runs short index lists (packed robustness variants) sizes it to n_sel.
  // Dropout probabilities, applied by a TRAINING forward and its backward only (csrc/dropout.cuh):
  float drop_hidden;  // BertConfig.hidden_dropout_prob: text embeddings, attention-output and
                      // FFN-output dense layers (before the residual add + LayerNorm)
  float drop_attn;    // BertConfig.attention_probs_dropout_prob: softmax(QK^T) before P V
  float drop_img;     // args.dropout: ImageBertEmbeddings.dropout (src/mmbt.py:56,82)
  int drop_reserved;
};

struct MmbtInputs {
  const long long* txt;      // (B, S_txt) token ids
  const long long* mask;     // (B, S_txt) attention mask, 1 = attend
  const long long* segment;  // (B, S_txt) token type ids
  const float* img;          // (B, n_img, d_img) fp32: ImageEncoder output (src/mmbt.py:40-45)
  // Positions of the full sequence [CLS img.. SEP | text..] that enter the encoder, ascending or
  // not, device int32[n_sel]; null = all n_img + 2 + S_txt positions (forward, :98-129).
  // forward_img_only (:131-153) = the first n_img + 2; forward_txt_only (:155-184) = {0} + text;
  // forward_control (:186-234) = {0} + the sampled subset.
  const int* indices;
  int n_sel;
  // 0: one index list shared by the batch; 1: `indices` is int32[B][n_sel], one list per sample
  // (packs several equally long robustness variants of a batch into ONE forward along the batch axis)
  int indices_per_sample;
  const void* params_bf16;   // optional caller-maintained bf16 shadow of params
  float* dimg;               // backward only: d loss / d img (B, n_img, d_img) fp32, or null
  unsigned long long drop_seed;  // seed of this forward's dropout masks (the backward needs the same)
};

int mmbt_param_table(const MmbtConfig& c, ParamEntry* out, int max_entries);  // returns count
long long mmbt_param_count(const MmbtConfig& c);
long long mmbt_workspace_bytes(const MmbtConfig& c, int training);
// logits: fp32 (B, C).  training != 0 keeps the activations for mmbt_backward.
int mmbt_forward(const MmbtConfig& c, const float* params, const MmbtInputs& in, void* ws,
                 long long ws_bytes, int training, float* logits, cudaStream_t stream);
// grads (same layout as params) are ACCUMULATED.
int mmbt_backward(const MmbtConfig& c, const float* params, const MmbtInputs& in, void* ws,
                  long long ws_bytes, const float* dlogits, float* grads, cudaStream_t stream);

// BertAdam (pytorch_pretrained_bert.optimization.BertAdam.step as called from train.py:142-147):
// per-tensor gradient-norm clipping (max_grad_norm), Adam moments WITHOUT bias correction,
// weight decay added to the update.  Per-tensor state['step'] means a per-tensor scheduled lr, so
// the caller supplies it per tensor.
// segs: device int64 [n_seg][2] = (offset, numel) of every tensor in the flat buffer;
// seg_hyper: device float [n_seg][2] = (weight_decay, scheduled lr) per tensor; norms: device float
// [n_seg] scratch; max_seg_numel: the largest tensor's element count (sizes the grid); grad_scale:
// factor applied to g before everything else (1 / world size after a sum-all-reduce).
int bertadam_flat(float* p, float* g, float* m, float* v, void* p_bf16, const long long* segs,
                  const float* seg_hyper, float* norms, int n_seg, long long max_seg_numel, float b1,
                  float b2, float eps, float max_grad_norm, float grad_scale, cudaStream_t stream);

}  // namespace mmu
