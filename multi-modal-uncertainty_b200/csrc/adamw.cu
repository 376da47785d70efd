// Fused AdamW over the flat fp32 parameter buffer (one launch per optimiser step).
//
// Reference: torch.optim.AdamW(lr, betas=(0.9, 0.98), eps=1e-9, weight_decay=wd) built in
// train.py:196-202, PyTorch's update rule:
//     p <- p (1 - lr wd);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2
//     p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: reads p, g, m, v and writes p, m, v = 28 B per parameter (+2 B for the optional
// bf16 shadow copy consumed by the tensor-core GEMMs).  128-bit loads/stores, L1 bypassed
// for the streaming operands, grid = a multiple of the SM count.
// `grad_scale` folds the 1/world_size of data-parallel gradient averaging into the same pass.
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>

#include "common.h"
#include "kernels.h"

namespace mmu {
namespace {

__device__ __forceinline__ float4 ldcs4(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void stcs4(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float lr_wd_keep,
                                          float b1, float b2, float step_size, float inv_bc2_sqrt,
                                          float eps) {
  p *= lr_wd_keep;
  m = b1 * m + (1.0f - b1) * g;
  v = b2 * v + (1.0f - b2) * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p -= step_size * (m / denom);
  return p;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, __nv_bfloat16* __restrict__ p_lp, size_t n4, float lr_wd_keep,
             float b1, float b2, float step_size, float inv_bc2_sqrt, float eps, float grad_scale) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    float4 pp = *reinterpret_cast<const float4*>(p + 4 * i);
    float4 gg = ldcs4(g + 4 * i);
    float4 mm = ldcs4(m + 4 * i);
    float4 vv = ldcs4(v + 4 * i);
    gg.x *= grad_scale; gg.y *= grad_scale; gg.z *= grad_scale; gg.w *= grad_scale;
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_wd_keep, b1, b2, step_size, inv_bc2_sqrt, eps);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_wd_keep, b1, b2, step_size, inv_bc2_sqrt, eps);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_wd_keep, b1, b2, step_size, inv_bc2_sqrt, eps);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_wd_keep, b1, b2, step_size, inv_bc2_sqrt, eps);
    *reinterpret_cast<float4*>(p + 4 * i) = pp;
    stcs4(m + 4 * i, mm);
    stcs4(v + 4 * i, vv);
    if (p_lp != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p_lp + 4 * i) = pk;
    }
  }
}

}  // namespace

int adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, size_t n, float lr,
               float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
               cudaStream_t stream) {
  if (n == 0) return 0;
  if (n % 4 != 0 || step < 1) return MMU_ERR_SHAPE;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
        reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) != 0)
    return MMU_ERR_ALIGN;
  // bias corrections in double on the host, exactly as torch computes them in Python floats
  const double bc1 = 1.0 - std::pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - std::pow(static_cast<double>(beta2), step);
  const float step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / std::sqrt(bc2));
  const float keep = static_cast<float>(1.0 - static_cast<double>(lr) * weight_decay);
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  const size_t cap = static_cast<size_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), n4, keep, beta1, beta2, step_size,
      inv_bc2_sqrt, eps, grad_scale);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    fprintf(stderr, "mmu: adamw launch failed: %s\n", cudaGetErrorString(err));
    return MMU_ERR_CUDA;
  }
  count_launch();
  return 0;
}

}  // namespace mmu
