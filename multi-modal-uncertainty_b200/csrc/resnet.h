// Four-view FashionMNIST ResNet engine (reference MIMOResNet, src/model.py:17-100 +
// src/layers.py:7-38).  See resnet.cu.
#pragma once
#include <cuda_runtime.h>

#include "engine.h"  // ParamEntry

namespace mmu {

struct ResNetConfig {
  int B;     // mini-batch
  int cin;   // input channels = num_channels * emb_dim (the views become channels, src/model.py:85)
  int H, W;  // 14 x 14 (quarter crops of FashionMNIST)
  int E;     // out_dim (heads of MultiHeadFC)
  int C;     // num_classes
};

int resnet_param_table(const ResNetConfig& c, ParamEntry* out, int max_entries);  // returns count
int resnet_stat_table(const ResNetConfig& c, ParamEntry* out, int max_entries);   // BN running stats
long long resnet_param_count(const ResNetConfig& c);   // padded flat length of params / grads
long long resnet_stat_count(const ResNetConfig& c);    // padded flat length of the statistics buffer
long long resnet_workspace_bytes(const ResNetConfig& c, int training);
// x: fp32 (B, cin, 14, 14) NCHW; logits: fp32 (B, E, C).  training != 0: batch statistics, running
// statistics updated in `stats`, activations kept for the backward.
// params_bf16: optional bf16 copy of `params` (same offsets): every convolution whose GEMM K is a
// multiple of 8 (all but the stem) then runs on the tcgen05 tensor-core kernel; null = fp32 path.
int resnet_forward(const ResNetConfig& c, const float* params, const void* params_bf16, float* stats,
                   const float* x_nchw, void* ws, long long ws_bytes, int training, float* logits,
                   cudaStream_t stream);
// grads (same layout as params) are ACCUMULATED.
int resnet_backward(const ResNetConfig& c, const float* params, const void* params_bf16, float* stats,
                    const float* x_nchw, void* ws, long long ws_bytes, const float* dlogits,
                    float* grads, cudaStream_t stream);

// ---- MMBT image encoder (reference src/mmbt.py:15-45): Bottleneck ResNet trunk + adaptive pool.
struct ImgEncConfig {
  int B;                // images
  int H;                // square input size (224)
  int layers[4];        // resnet152: {3, 8, 36, 3}
  int width_per_group;  // 64 (torchvision's base width; thinner nets for tests)
  int pool_h, pool_w;   // adaptive pool grid: num_image_embeds cells (src/mmbt.py:28-37)
  int pool_max;         // args.img_embed_pool_type != "avg"
};
int imgenc_param_table(const ImgEncConfig& c, ParamEntry* out, int max_entries);
int imgenc_stat_table(const ImgEncConfig& c, ParamEntry* out, int max_entries);
long long imgenc_param_count(const ImgEncConfig& c);
long long imgenc_stat_count(const ImgEncConfig& c);
long long imgenc_workspace_bytes(const ImgEncConfig& c, int training);
// x: fp32 (B, 3, H, H) NCHW; tokens: fp32 (B, pool_h * pool_w, 2048).
int imgenc_forward(const ImgEncConfig& c, const float* params, const void* params_bf16, float* stats,
                   const float* x_nchw, void* ws, long long ws_bytes, int training, float* tokens,
                   cudaStream_t stream);
int imgenc_backward(const ImgEncConfig& c, const float* params, const void* params_bf16, float* stats,
                    const float* x_nchw, void* ws, long long ws_bytes, const float* dtokens, float* grads,
                    cudaStream_t stream);

}  // namespace mmu
