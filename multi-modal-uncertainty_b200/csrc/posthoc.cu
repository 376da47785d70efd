// Post-hoc robustness scoring on device: what the reference's notebooks
// compute on the CPU from the dumped (S, 43, K, C) logits --
//   p(true label) after averaging the head PROBABILITIES   (notebooks/food101_robustness.py:24-46,
//                                                            notebooks/utils.py:22-23 softmax)
//   Pearson r between the experimental delta-p and the mean control delta-p, per modality
//                                                           (notebooks/utils.py:26-34)
//   accuracy of the head-mean LOGITS per variant            (food101_robustness.py:48-77)
// -- accumulated per batch straight from the packed-variant logits (V, B, E, C), so the
// (S, 43, K, C) array never has to leave the GPU.  Variant order as in
// eval_transformer_robustness.py:103-121: 0 full, 1 image only, 2 text only, then n_repeats
// image-controlled draws, then n_repeats text-controlled draws.
//
// One warp per sample; per variant the warp reduces the E*C logits with shuffles (max / sum-exp
// per head for the label's probability, sequential head sum + correctly rounded division for the
// head-mean logits, first-index argmax).  Sufficient statistics for the two Pearson coefficients
// are accumulated in fp64.
#include <cstdio>

#include "common.h"
#include "kernels.h"

namespace mmu {
namespace {

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int PH_WARPS = 4;
constexpr int PH_MAX_V = 128;

__global__ void __launch_bounds__(PH_WARPS * 32)
posthoc_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int V, int B,
               int E, int C, int n_rep, float* __restrict__ p_true_out, PosthocAccum* __restrict__ acc) {
  __shared__ float pt[PH_WARPS][PH_MAX_V];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * PH_WARPS + warp;
  if (b >= B) return;
  const int y = static_cast<int>(labels[b]);
  const float fE = static_cast<float>(E);
  for (int v = 0; v < V; ++v) {
    const float* z = logits + (static_cast<size_t>(v) * B + b) * E * C;
    float p_lab = 0.f;
    // head-mean logits (for the accuracy): sequential sum over heads, then a true division
    float best = -INFINITY;
    int best_c = 0x7fffffff;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      if (c < C) {
        float s = 0.f;
        for (int k = 0; k < E; ++k) s += z[k * C + c];
        s = s / fE;
        if (s > best) { best = s; best_c = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
      if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
    }
    for (int k = 0; k < E; ++k) {
      const float* zk = z + k * C;
      float m = -INFINITY;
      for (int c = lane; c < C; c += 32) m = fmaxf(m, zk[c]);
      m = wmax(m);
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s += expf(zk[c] - m);
      s = wsum(s);
      p_lab += expf(zk[y] - m) / s;
    }
    p_lab /= fE;
    if (lane == 0) {
      pt[warp][v] = p_lab;
      if (p_true_out != nullptr) p_true_out[static_cast<size_t>(b) * V + v] = p_lab;
      if (best_c == y) atomicAdd(&acc->correct[v], 1ull);
    }
  }
  __syncwarp();
  if (lane < 2) {  // lane 0: image, lane 1: text
    const float ori = pt[warp][0];
    const float x = pt[warp][1 + lane] - ori;
    float ctl = 0.f;
    const int base = 3 + lane * n_rep;
    for (int r = 0; r < n_rep; ++r) ctl += pt[warp][base + r] - ori;
    const double xd = x, yd = n_rep > 0 ? static_cast<double>(ctl) / n_rep : 0.0;
    atomicAdd(&acc->sx[lane], xd);
    atomicAdd(&acc->sy[lane], yd);
    atomicAdd(&acc->sxx[lane], xd * xd);
    atomicAdd(&acc->syy[lane], yd * yd);
    atomicAdd(&acc->sxy[lane], xd * yd);
  }
  if (lane == 0) atomicAdd(&acc->n_samples, 1ull);
}

}  // namespace

int posthoc_scoring(const float* logits, const long long* labels, int V, int B, int E, int C,
                    int n_repeats, float* p_true_out, PosthocAccum* acc, cudaStream_t stream) {
  if (V < 3 || V > PH_MAX_V || V != 3 + 2 * n_repeats || E < 1 || C < 1 || B < 1) return MMU_ERR_SHAPE;
  posthoc_kernel<<<(B + PH_WARPS - 1) / PH_WARPS, PH_WARPS * 32, 0, stream>>>(
      logits, labels, V, B, E, C, n_repeats, p_true_out, acc);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}

}  // namespace mmu
