// FLAVA-fusion engine: host-side orchestration of the forward / backward pass of
// FlavaFusionTransfomer[withCLSToken] (reference src/model.py:225-374) over caller-owned
// flat buffers.  No allocation, no synchronisation, no exceptions: everything is enqueued on
// the caller's stream and errors come back as negative codes.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace mmu {

enum Precision : int { PREC_FP32 = 0, PREC_BF16 = 1 };

struct FlavaConfig {
  int B;          // mini-batch (the attended axis!)
  int l_img;      // max image tokens per sample
  int l_txt;      // max text tokens per sample
  int d_img;      // image_hidden_size
  int d_txt;      // text_hidden_size
  int D;          // multimodal_hidden_size
  int n_head;     // multimodal_num_attention_heads
  int n_layers;   // multimodal_num_hidden_layers
  int E;          // out_dim (number of heads / ensemble members)
  int C;          // num_classes
  int avg_pool;   // kwargs["avg_pool"]
  int cls_token;  // FlavaFusionTransfomerwithCLSToken
  int precision;  // Precision
  int max_variants;  // capacity for packed-variant evaluation (0 or 1: single variant)
  // MIMOTransfomer (reference src/model.py:114-159): d_txt == 0 means "no text modality" (its
  // projection does not exist), and head e pools the mean of token positions
  // [e*group_pool, (e+1)*group_pool) -- the `x.view(b, e, c, -1).mean(2)` of :148-149.
  int group_pool;    // 0: FLAVA head wiring (token e / avg_pool)
};

struct ParamEntry {
  char name[96];  // reference state_dict key
  long long offset;  // in elements, into the flat fp32 parameter / gradient buffers
  long long numel;
  int rows, cols;    // cols == 0 for vectors
  int stage;         // backward stage that finishes this gradient (for bucketed all-reduce)
};

int flava_param_table(const FlavaConfig& c, ParamEntry* out, int max_entries);  // returns count
long long flava_param_count(const FlavaConfig& c);  // padded flat length (elements)
long long flava_workspace_bytes(const FlavaConfig& c, int training);
int flava_num_stages(const FlavaConfig& c);  // backward stages: heads, layers (reversed), stem

struct FlavaInputs {
  const void* img;      // (B, l_img, d_img) fp32 (bf16 with src_bf16) or null
  const void* txt;      // (B, l_txt, d_txt) fp32 (bf16 with src_bf16) or null
  const int* idx_img;   // device int32[n_img] token subset, or null for identity
  const int* idx_txt;
  int n_img;            // tokens used (0: modality absent); <= l_img
  int n_txt;
  const int* keep;      // device int32[B,2] modality keep mask (zero-fill), or null
  // optional caller-maintained bf16 shadow of `params` (same element offsets).  Null: the engine
  // casts the fp32 master into the workspace at every forward (22.8 M params = 137 MB of traffic).
  const void* params_bf16;
  // Packed-variant evaluation (robustness sweeps): token positions never interact (attention
  // runs over the BATCH axis, every other op is row-wise), so V token-subset variants of one
  // batch are evaluated in ONE pass by concatenating their tokens along L: idx_img / idx_txt
  // hold the concatenated per-variant index lists (n_img / n_txt = totals, <= config l_img /
  // l_txt, which act as capacities), and var_segments says which positions feed each head.
  int src_l_img;            // token count of the `img` source tensor (0: config l_img)
  int src_l_txt;
  int n_variants;           // 0 or 1: ordinary forward
  const int* var_segments;  // device int32 [n_variants][E][2]: rows [begin, end) of head e
  // nn.Dropout(drop) between c_fc and QuickGELU (src/model.py:195-201), training only: counter-
  // based masks (csrc/dropout.cuh), site = layer index; the backward must see the same values.
  float drop_p;
  int src_bf16;  // 1: img / txt point to bf16 tensors (bf16 host staging), 0: fp32
  unsigned long long drop_seed;
};

int flava_forward(const FlavaConfig& c, const float* params, const FlavaInputs& in, void* ws,
                  long long ws_bytes, int training, float* logits, cudaStream_t stream);
int flava_backward(const FlavaConfig& c, const float* params, const FlavaInputs& in, void* ws,
                   long long ws_bytes, const float* dlogits, float* grads, int stage_begin,
                   int stage_end, cudaStream_t stream);

}  // namespace mmu
