// Row-wise HBM-bound kernels: input staging (gather/mask/cast), LayerNorm forward/backward,
// column sums (bias gradients), ln_post + pooling + classifier heads, CLS rows.
// One warp owns one row; every global access is a 128-bit (or 64-bit for bf16) vector,
// reductions are warp shuffles, per-column partial sums are reduced through shared memory and
// finished with one atomicAdd per column per block.
#include <cuda_bf16.h>

#include <cstdio>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace mmu {

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

#define MMU_CHECK_LAUNCH()                                                      \
  do {                                                                          \
    const cudaError_t err__ = cudaGetLastError();                               \
    if (err__ != cudaSuccess) {                                                 \
      fprintf(stderr, "mmu: launch failed at %s:%d: %s\n", __FILE__, __LINE__, \
              cudaGetErrorString(err__));                                       \
      return MMU_ERR_CUDA;                                                      \
    }                                                                           \
    count_launch();                                                             \
  } while (0)

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) {
    return *reinterpret_cast<const float4*>(p);
  }
  static __device__ __forceinline__ void st(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 pk = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = pk;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

int grid_for(size_t work_items, int per_block) {
  size_t blocks = (work_items + per_block - 1) / per_block;
  const size_t cap = static_cast<size_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------ cast / gather / mask
template <typename T, typename TS = float>
__global__ void cast_gather_kernel(const TS* __restrict__ src, T* __restrict__ dst, int B,
                                   int l_src, int d, const int* __restrict__ idx, int n_sel,
                                   const int* __restrict__ keep, int modality, int pos_major) {
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue early
  const int dv = d >> 2;
  const size_t total = static_cast<size_t>(B) * n_sel * dv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % dv);
    const size_t row = i / dv;  // destination row: (b, j) batch-major, (j, b) position-major
    const int j = static_cast<int>(pos_major ? row / B : row % n_sel);
    const int b = static_cast<int>(pos_major ? row % B : row / n_sel);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (keep == nullptr || keep[b * 2 + modality] != 0) {
      const int l = idx != nullptr ? idx[j] : j;
      v = Vec4<TS>::ld(src + (static_cast<size_t>(b) * l_src + l) * d + 4 * c);
    }
    Vec4<T>::st(dst + row * d + 4 * c, v);
  }
}

// ------------------------------------------------------------------ modality-dropout keep mask
// Guided / random modality dropout (north-star extension; definition: oracle/shaping.py
// modality_dropout_mask).  u, r: fp32[B] uniform draws made by the HOST generator; sample b keeps
// both modalities unless u[b] < p_drop; then it drops image / text by r[b] < 0.5 (mode 0, random)
// or the modality whose device-resident score is higher, ties -> image (mode 1, guided).  The
// guided scores never leave the device, so the training step does not wait on a host read.
__global__ void modality_keep_mask_kernel(const float* __restrict__ u, const float* __restrict__ r,
                                          const float* __restrict__ score_img,
                                          const float* __restrict__ score_txt, int score_stride, int B,
                                          float p_drop, int mode, int* __restrict__ keep) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int ki = 1, kt = 1;
  if (u[b] < p_drop) {
    const bool drop_text = mode == 0 ? !(r[b] < 0.5f)
                                     : score_txt[static_cast<size_t>(b) * score_stride] >
                                           score_img[static_cast<size_t>(b) * score_stride];
    if (drop_text) kt = 0; else ki = 0;
  }
  keep[2 * b] = ki;
  keep[2 * b + 1] = kt;
}

// ----------------------------------------------------------------------- LayerNorm
// Row held in registers: NV float4 per lane (D <= 128*NV).
template <int NV>
struct RowRegs {
  float4 v[NV];
  template <typename T>
  __device__ __forceinline__ void load(const T* row, int nvec, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < nvec ? Vec4<T>::ld(row + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ __forceinline__ float sum() const {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    return s;
  }
};

template <int NV>
__device__ __forceinline__ void row_stats(const RowRegs<NV>& r, int nvec, int lane, int D,
                                          float& mean, float& rstd, float eps = 1e-5f) {
  mean = warp_sum(r.sum()) / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (lane + 32 * i < nvec) {
      const float a = r.v[i].x - mean, b = r.v[i].y - mean, c = r.v[i].z - mean,
                  d = r.v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
  }
  rstd = rsqrtf(warp_sum(ss) / D + eps);
}

template <int NV, typename TO>
__global__ void __launch_bounds__(THREADS)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, TO* __restrict__ y, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int D) {
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue early
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int row = blockIdx.x * WARPS + warp; row < M; row += gridDim.x * WARPS) {
    RowRegs<NV> r;
    r.load(x + static_cast<size_t>(row) * D, nvec, lane);
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
    if (lane == 0 && mean_out != nullptr) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 b = *reinterpret_cast<const float4*>(beta + 4 * c);
        float4 o;
        o.x = (r.v[i].x - mean) * rstd * g.x + b.x;
        o.y = (r.v[i].y - mean) * rstd * g.y + b.y;
        o.z = (r.v[i].z - mean) * rstd * g.z + b.z;
        o.w = (r.v[i].w - mean) * rstd * g.w + b.w;
        Vec4<TO>::st(y + static_cast<size_t>(row) * D + 4 * c, o);
      }
    }
  }
}


// Two chained LayerNorms in one pass over the row: y1 = LN1(x) (fp32: the residual stream after
// ln_pre), y2 = LN2(y1) (the first block's ln_1 output, activation dtype).  Saves re-reading y1.
template <int NV, typename TO>
__global__ void __launch_bounds__(THREADS)
layernorm2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g1,
                      const float* __restrict__ b1, float* __restrict__ y1,
                      float* __restrict__ mean1, float* __restrict__ rstd1,
                      const float* __restrict__ g2, const float* __restrict__ b2,
                      TO* __restrict__ y2, float* __restrict__ mean2, float* __restrict__ rstd2, int M,
                      int D) {
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue early
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int row = blockIdx.x * WARPS + warp; row < M; row += gridDim.x * WARPS) {
    RowRegs<NV> r;
    r.load(x + static_cast<size_t>(row) * D, nvec, lane);
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
    if (lane == 0) {
      mean1[row] = mean;
      rstd1[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(g1 + 4 * c);
        const float4 b = *reinterpret_cast<const float4*>(b1 + 4 * c);
        float4& v = r.v[i];
        v.x = (v.x - mean) * rstd * g.x + b.x;
        v.y = (v.y - mean) * rstd * g.y + b.y;
        v.z = (v.z - mean) * rstd * g.z + b.z;
        v.w = (v.w - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(y1 + static_cast<size_t>(row) * D + 4 * c) = v;
      }
    }
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
    if (lane == 0) {
      mean2[row] = mean;
      rstd2[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(g2 + 4 * c);
        const float4 b = *reinterpret_cast<const float4*>(b2 + 4 * c);
        float4 o;
        o.x = (r.v[i].x - mean) * rstd * g.x + b.x;
        o.y = (r.v[i].y - mean) * rstd * g.y + b.y;
        o.z = (r.v[i].z - mean) * rstd * g.z + b.z;
        o.w = (r.v[i].w - mean) * rstd * g.w + b.w;
        Vec4<TO>::st(y2 + static_cast<size_t>(row) * D + 4 * c, o);
      }
    }
  }
}

// Eval path with LayerNorm folded into the consumer GEMM (gemm_api.h, ln_stats): y = LN(x; g, b)
// in fp32 (the residual stream after ln_pre), a RAW bf16 copy of y (the next GEMM's A operand) and
// the row's (sum, sum of squares) in partial slot 0 of stats[row][nt][2] (slots 1.. zeroed).
template <int NV>
__global__ void __launch_bounds__(THREADS)
layernorm_raw_stats_kernel(const float* __restrict__ x, const float* __restrict__ g,
                           const float* __restrict__ b, float* __restrict__ y,
                           __nv_bfloat16* __restrict__ yraw, float* __restrict__ stats, int nt, int M,
                           int D) {
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue early
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int row = blockIdx.x * WARPS + warp; row < M; row += gridDim.x * WARPS) {
    RowRegs<NV> r;
    r.load(x + static_cast<size_t>(row) * D, nvec, lane);
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 gg = *reinterpret_cast<const float4*>(g + 4 * c);
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * c);
        float4& v = r.v[i];
        v.x = (v.x - mean) * rstd * gg.x + bb.x;
        v.y = (v.y - mean) * rstd * gg.y + bb.y;
        v.z = (v.z - mean) * rstd * gg.z + bb.z;
        v.w = (v.w - mean) * rstd * gg.w + bb.w;
        *reinterpret_cast<float4*>(y + static_cast<size_t>(row) * D + 4 * c) = v;
        Vec4<__nv_bfloat16>::st(yraw + static_cast<size_t>(row) * D + 4 * c, v);
        s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
    }
    const float s1 = warp_sum(r.sum());
    s2 = warp_sum(s2);
    if (lane < nt) {
      float2 o = make_float2(0.f, 0.f);
      if (lane == 0) o = make_float2(s1, s2);
      reinterpret_cast<float2*>(stats)[static_cast<size_t>(row) * nt + lane] = o;
    }
  }
}

// Weights of a Linear that consumes LayerNorm output, folded for the LN-in-the-epilogue GEMM:
//   Wf[n][k] = bf16(W[n][k] * gamma[k]);  cw[n] = sum_k float(Wf[n][k]);
//   bf[n] = bias[n] + sum_k beta[k] * W[n][k]
// so that  LN(x) W^T + bias = rstd * (x Wf^T - mean * cw) + bf.   One warp per output row n;
// blockIdx.y selects one of two (W, gamma, beta, bias) sets (in_proj with ln_1, c_fc with ln_2).
struct FoldSet {
  const float* W; const float* gamma; const float* beta; const float* bias;
  __nv_bfloat16* Wf; float* cw; float* bf;
  int N;
};
// blockIdx.z = layer: every pointer of a set moves by a fixed stride per layer (pstride floats for
// the parameters, wstride BYTES for the folded outputs)
__global__ void __launch_bounds__(256)
ln_fold_weights_kernel(FoldSet s0, FoldSet s1, int K, long long pstride, long long wstride) {
  FoldSet s = blockIdx.y == 0 ? s0 : s1;
  {
    const long long po = pstride * blockIdx.z, wo = wstride * blockIdx.z;
    s.W += po; s.gamma += po; s.beta += po;
    if (s.bias != nullptr) s.bias += po;
    s.Wf = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(s.Wf) + wo);
    s.cw = reinterpret_cast<float*>(reinterpret_cast<char*>(s.cw) + wo);
    s.bf = reinterpret_cast<float*>(reinterpret_cast<char*>(s.bf) + wo);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kv = K >> 2;
  for (int n = blockIdx.x * 8 + warp; n < s.N; n += gridDim.x * 8) {
    const float* w = s.W + static_cast<size_t>(n) * K;
    float acc_w = 0.f, acc_b = 0.f;
    for (int c = lane; c < kv; c += 32) {
      const float4 wv = *reinterpret_cast<const float4*>(w + 4 * c);
      const float4 g = *reinterpret_cast<const float4*>(s.gamma + 4 * c);
      const float4 b = *reinterpret_cast<const float4*>(s.beta + 4 * c);
      const __nv_bfloat16 f0 = __float2bfloat16_rn(wv.x * g.x), f1 = __float2bfloat16_rn(wv.y * g.y),
                          f2 = __float2bfloat16_rn(wv.z * g.z), f3 = __float2bfloat16_rn(wv.w * g.w);
      __nv_bfloat162 lo, hi;
      lo.x = f0; lo.y = f1; hi.x = f2; hi.y = f3;
      uint2 packed;
      packed.x = *reinterpret_cast<const uint32_t*>(&lo);
      packed.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(s.Wf + static_cast<size_t>(n) * K + 4 * c) = packed;
      acc_w += (__bfloat162float(f0) + __bfloat162float(f1)) + (__bfloat162float(f2) + __bfloat162float(f3));
      acc_b += (wv.x * b.x + wv.y * b.y) + (wv.z * b.z + wv.w * b.w);
    }
    acc_w = warp_sum(acc_w);
    acc_b = warp_sum(acc_b);
    if (lane == 0) {
      s.cw[n] = acc_w;
      s.bf[n] = (s.bias != nullptr ? s.bias[n] : 0.f) + acc_b;
    }
  }
}

// x_out = x_in + y (branch output of the preceding projection GEMM, stored in the activation
// dtype); h = LayerNorm(x_out).  Fusing the residual add here keeps the GEMM epilogues free of the
// fp32 residual-stream traffic (a streaming row kernel moves those bytes at ~HBM peak; a GEMM
// epilogue does not).  gamma == nullptr: only the sum is written (last block -> ln_post/heads).
template <int NV, typename TY, typename TO>
__global__ void __launch_bounds__(THREADS)
add_layernorm_fwd_kernel(const float* __restrict__ x_in, const TY* __restrict__ y,
                         float* __restrict__ x_out, const float* __restrict__ gamma,
                         const float* __restrict__ beta, TO* __restrict__ h,
                         float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D) {
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue early
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int row = blockIdx.x * WARPS + warp; row < M; row += gridDim.x * WARPS) {
    RowRegs<NV> r, ry;
    r.load(x_in + static_cast<size_t>(row) * D, nvec, lane);
    ry.load(y + static_cast<size_t>(row) * D, nvec, lane);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      r.v[i].x += ry.v[i].x; r.v[i].y += ry.v[i].y; r.v[i].z += ry.v[i].z; r.v[i].w += ry.v[i].w;
      const int c = lane + 32 * i;
      if (c < nvec) *reinterpret_cast<float4*>(x_out + static_cast<size_t>(row) * D + 4 * c) = r.v[i];
    }
    if (gamma == nullptr) continue;
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 b = *reinterpret_cast<const float4*>(beta + 4 * c);
        float4 o;
        o.x = (r.v[i].x - mean) * rstd * g.x + b.x;
        o.y = (r.v[i].y - mean) * rstd * g.y + b.y;
        o.z = (r.v[i].z - mean) * rstd * g.z + b.z;
        o.w = (r.v[i].w - mean) * rstd * g.w + b.w;
        Vec4<TO>::st(h + static_cast<size_t>(row) * D + 4 * c, o);
      }
    }
  }
}

// Post-LN residual (BERT, pytorch_pretrained_bert BertSelfOutput / BertOutput; reference call
// site src/mmbt.py:124-128): s = x_in + y with x_in the previous LayerNorm output in the ACTIVATION
// dtype, s kept in fp32 for the backward, h = LayerNorm(s; eps) in the activation dtype.
// y == nullptr: plain LayerNorm of x_in.
template <int NV, typename T>
__global__ void __launch_bounds__(THREADS)
postln_fwd_kernel(const T* __restrict__ x_in, const T* __restrict__ y, float* __restrict__ s_out,
                  const float* __restrict__ gamma, const float* __restrict__ beta,
                  T* __restrict__ h, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                  int M, int D, float eps, const dropout::Site ydrop) {
  ptx::pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int row = blockIdx.x * WARPS + warp; row < M; row += gridDim.x * WARPS) {
    RowRegs<NV> r, ry;
    r.load(x_in + static_cast<size_t>(row) * D, nvec, lane);
    if (y != nullptr) {
      ry.load(y + static_cast<size_t>(row) * D, nvec, lane);
      if (ydrop.on()) {  // hidden dropout on the branch output, before the residual add
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const unsigned int e0 = static_cast<unsigned int>(row) * D + 4u * (lane + 32 * i);
          ry.v[i].x *= ydrop.mult(e0); ry.v[i].y *= ydrop.mult(e0 + 1);
          ry.v[i].z *= ydrop.mult(e0 + 2); ry.v[i].w *= ydrop.mult(e0 + 3);
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        r.v[i].x += ry.v[i].x; r.v[i].y += ry.v[i].y; r.v[i].z += ry.v[i].z; r.v[i].w += ry.v[i].w;
      }
    }
    if (s_out != nullptr) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) *reinterpret_cast<float4*>(s_out + static_cast<size_t>(row) * D + 4 * c) = r.v[i];
      }
    }
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd, eps);
    if (lane == 0 && mean_out != nullptr) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 b = *reinterpret_cast<const float4*>(beta + 4 * c);
        float4 o;
        o.x = (r.v[i].x - mean) * rstd * g.x + b.x;
        o.y = (r.v[i].y - mean) * rstd * g.y + b.y;
        o.z = (r.v[i].z - mean) * rstd * g.z + b.z;
        o.w = (r.v[i].w - mean) * rstd * g.w + b.w;
        Vec4<T>::st(h + static_cast<size_t>(row) * D + 4 * c, o);
      }
    }
  }
}

// Reduce per-warp column partials (acc[NV] float4 per lane) across the block's warps and
// atomically add into out[D].  `red` is WARPS*D floats of shared memory.
template <int NV>
__device__ __forceinline__ void block_colreduce(const float4 (&acc)[NV], float* red, float* out,
                                                int nvec, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * c) = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[w * D + c];
    atomicAdd(out + c, s);
  }
}

template <int NV>
__device__ __forceinline__ void zero_acc(float4 (&a)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// LayerNorm backward.  HBM traffic per row (bf16 path): dy 2D + x 4D + dx 4D read, dx 4D + dx_lp 2D
// written = 16 B per element.  The three per-column accumulators (dgamma, dbeta, colsum of the
// final dx) cost 72 registers per thread, so the kernel runs 4-warp blocks (3 per SM) and issues
// every load of a row -- including the previous dx it accumulates into -- before the first use:
// one memory round trip per row and ~90 KB in flight per SM.
constexpr int LNB_WARPS = 4;
constexpr int LNB_THREADS = LNB_WARPS * 32;

template <int NV, int NW>
__device__ __forceinline__ void block_colreduce_n(const float4 (&acc)[NV], float* red, float* out,
                                                  int nvec, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * c) = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w * D + c];
    atomicAdd(out + c, s);
  }
}

template <int NV, typename TDY, typename TLP>
__global__ void __launch_bounds__(LNB_THREADS, 3)
layernorm_bwd_kernel(const TDY* __restrict__ dy, const float* __restrict__ x,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, float* __restrict__ dx, int accumulate,
                     TLP* __restrict__ dx_lp, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dcolsum, int M, int D) {
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue early
  extern __shared__ float red[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  float4 acc_g[NV], acc_b[NV], acc_c[NV];
  zero_acc<NV>(acc_g);
  zero_acc<NV>(acc_b);
  zero_acc<NV>(acc_c);
  for (int row = blockIdx.x * LNB_WARPS + warp; row < M; row += gridDim.x * LNB_WARPS) {
    RowRegs<NV> rx, rdy, rp;
    rx.load(x + static_cast<size_t>(row) * D, nvec, lane);
    rdy.load(dy + static_cast<size_t>(row) * D, nvec, lane);
    if (accumulate) rp.load(dx + static_cast<size_t>(row) * D, nvec, lane);
    else zero_acc<NV>(rp.v);
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);  // L1-resident
        float4& xv = rx.v[i];
        xv.x = (xv.x - mu) * rs; xv.y = (xv.y - mu) * rs;
        xv.z = (xv.z - mu) * rs; xv.w = (xv.w - mu) * rs;  // xhat
        float4& d = rdy.v[i];
        acc_g[i].x += d.x * xv.x; acc_g[i].y += d.y * xv.y;
        acc_g[i].z += d.z * xv.z; acc_g[i].w += d.w * xv.w;
        acc_b[i].x += d.x; acc_b[i].y += d.y; acc_b[i].z += d.z; acc_b[i].w += d.w;
        d.x *= g.x; d.y *= g.y; d.z *= g.z; d.w *= g.w;  // dy * gamma from here on
        s1 += (d.x + d.y) + (d.z + d.w);
        s2 += (d.x * xv.x + d.y * xv.y) + (d.z * xv.z + d.w * xv.w);
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 d = rdy.v[i], xv = rx.v[i], prev = rp.v[i];
        float4 o;
        o.x = prev.x + rs * (d.x - s1 - xv.x * s2);
        o.y = prev.y + rs * (d.y - s1 - xv.y * s2);
        o.z = prev.z + rs * (d.z - s1 - xv.z * s2);
        o.w = prev.w + rs * (d.w - s1 - xv.w * s2);
        *reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * D + 4 * c) = o;
        if (dx_lp != nullptr) Vec4<TLP>::st(dx_lp + static_cast<size_t>(row) * D + 4 * c, o);
        acc_c[i].x += o.x; acc_c[i].y += o.y; acc_c[i].z += o.z; acc_c[i].w += o.w;
      }
    }
  }
  block_colreduce_n<NV, LNB_WARPS>(acc_g, red, dgamma, nvec, D);
  block_colreduce_n<NV, LNB_WARPS>(acc_b, red, dbeta, nvec, D);
  if (dcolsum != nullptr) block_colreduce_n<NV, LNB_WARPS>(acc_c, red, dcolsum, nvec, D);
}

// Post-LN backward: dx = LN'(dy_branch + dy_res) where dy_branch (activation dtype, may be null)
// is the gradient arriving through the sub-layer's first GEMM and dy_res (fp32, may be null) the
// gradient arriving through the residual connection -- their sum is never materialised.
// Writes dx (fp32) and its activation-dtype copy, accumulates dgamma / dbeta and the column sum of
// dx (= bias gradient of the projection that closed the sub-layer).
template <int NV, typename T>
__global__ void __launch_bounds__(LNB_THREADS, 3)
postln_bwd_kernel(const T* __restrict__ dy_branch, const float* __restrict__ dy_res,
                  const float* __restrict__ x, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                  float* __restrict__ dx, T* __restrict__ dx_lp, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, float* __restrict__ dcolsum, int M, int D,
                  const PostLnDropout dr) {
  ptx::pdl_trigger();
  extern __shared__ float red[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  float4 acc_g[NV], acc_b[NV], acc_c[NV];
  zero_acc<NV>(acc_g);
  zero_acc<NV>(acc_b);
  zero_acc<NV>(acc_c);
  for (int row = blockIdx.x * LNB_WARPS + warp; row < M; row += gridDim.x * LNB_WARPS) {
    RowRegs<NV> rx, rdy, rr;
    rx.load(x + static_cast<size_t>(row) * D, nvec, lane);
    if (dy_branch != nullptr) rdy.load(dy_branch + static_cast<size_t>(row) * D, nvec, lane);
    else zero_acc<NV>(rdy.v);
    if (dy_res != nullptr) {
      rr.load(dy_res + static_cast<size_t>(row) * D, nvec, lane);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        rdy.v[i].x += rr.v[i].x; rdy.v[i].y += rr.v[i].y; rdy.v[i].z += rr.v[i].z; rdy.v[i].w += rr.v[i].w;
      }
    }
    if (dr.in_a.on() || dr.in_b.on()) {  // dropout applied to the LayerNorm OUTPUT in the forward
      const dropout::Site site = (dr.row_side != nullptr && dr.row_side[row] == 0) ? dr.in_b : dr.in_a;
      if (site.on()) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const unsigned int e0 = static_cast<unsigned int>(row) * D + 4u * (lane + 32 * i);
          rdy.v[i].x *= site.mult(e0); rdy.v[i].y *= site.mult(e0 + 1);
          rdy.v[i].z *= site.mult(e0 + 2); rdy.v[i].w *= site.mult(e0 + 3);
        }
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        float4& xv = rx.v[i];
        xv.x = (xv.x - mu) * rs; xv.y = (xv.y - mu) * rs;
        xv.z = (xv.z - mu) * rs; xv.w = (xv.w - mu) * rs;
        float4& d = rdy.v[i];
        acc_g[i].x += d.x * xv.x; acc_g[i].y += d.y * xv.y;
        acc_g[i].z += d.z * xv.z; acc_g[i].w += d.w * xv.w;
        acc_b[i].x += d.x; acc_b[i].y += d.y; acc_b[i].z += d.z; acc_b[i].w += d.w;
        d.x *= g.x; d.y *= g.y; d.z *= g.z; d.w *= g.w;
        s1 += (d.x + d.y) + (d.z + d.w);
        s2 += (d.x * xv.x + d.y * xv.y) + (d.z * xv.z + d.w * xv.w);
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 d = rdy.v[i], xv = rx.v[i];
        float4 o;
        o.x = rs * (d.x - s1 - xv.x * s2);
        o.y = rs * (d.y - s1 - xv.y * s2);
        o.z = rs * (d.z - s1 - xv.z * s2);
        o.w = rs * (d.w - s1 - xv.w * s2);
        *reinterpret_cast<float4*>(dx + static_cast<size_t>(row) * D + 4 * c) = o;
        if (dr.out.on()) {  // the branch (and its bias) see the gradient through their dropout mask
          const unsigned int e0 = static_cast<unsigned int>(row) * D + 4u * c;
          o.x *= dr.out.mult(e0); o.y *= dr.out.mult(e0 + 1);
          o.z *= dr.out.mult(e0 + 2); o.w *= dr.out.mult(e0 + 3);
        }
        if (dx_lp != nullptr) Vec4<T>::st(dx_lp + static_cast<size_t>(row) * D + 4 * c, o);
        acc_c[i].x += o.x; acc_c[i].y += o.y; acc_c[i].z += o.z; acc_c[i].w += o.w;
      }
    }
  }
  block_colreduce_n<NV, LNB_WARPS>(acc_g, red, dgamma, nvec, D);
  block_colreduce_n<NV, LNB_WARPS>(acc_b, red, dbeta, nvec, D);
  if (dcolsum != nullptr) block_colreduce_n<NV, LNB_WARPS>(acc_c, red, dcolsum, nvec, D);
}

// Column sums (bias gradients).  Each thread owns 4 columns and walks a slab of rows with 8
// independent 8/16-byte loads in flight (the op is a pure streaming read: 2 B/elem in bf16).
template <typename T>
__global__ void __launch_bounds__(128)
colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int N, int rows_per_block) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = Vec4<T>::ld(x + static_cast<size_t>(r + u) * N + c);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
  }
  for (; r < r1; ++r) {
    const float4 v = Vec4<T>::ld(x + static_cast<size_t>(r) * N + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  atomicAdd(out + c, s.x);
  atomicAdd(out + c + 1, s.y);
  atomicAdd(out + c + 2, s.z);
  atomicAdd(out + c + 3, s.w);
}

// ------------------------------------------------- ln_post + pooling + heads (small)
template <int NV>
__global__ void __launch_bounds__(THREADS)
pool_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                   const float* __restrict__ beta, HeadSegments seg, float* __restrict__ vec,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int L, int D) {
  extern __shared__ float red[];
  const int b = blockIdx.x, e = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const int s0 = seg.seg_begin[e], s1 = seg.seg_end[e];
  float4 acc[NV];
  zero_acc<NV>(acc);
  for (int l = s0 + warp; l < s1; l += WARPS) {
    const size_t row = static_cast<size_t>(l) * gridDim.x + b;   // position-major rows (l, b)
    RowRegs<NV> r;
    r.load(x + row * D, nvec, lane);
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 bb = *reinterpret_cast<const float4*>(beta + 4 * c);
        acc[i].x += (r.v[i].x - mean) * rstd * g.x + bb.x;
        acc[i].y += (r.v[i].y - mean) * rstd * g.y + bb.y;
        acc[i].z += (r.v[i].z - mean) * rstd * g.z + bb.z;
        acc[i].w += (r.v[i].w - mean) * rstd * g.w + bb.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * c) = acc[i];
  }
  __syncthreads();
  const float inv = 1.0f / static_cast<float>(max(1, s1 - s0));
  float* out = vec + (static_cast<size_t>(b) * seg.E + e) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[w * D + c];
    out[c] = s * inv;
  }
}

// Packed-variant twin of pool_ln_fwd_kernel (eval only: no statistics are kept).  Block
// (b, e, v) pools the ln_post rows [segs[v][e][0], segs[v][e][1]) of sample b.
template <int NV>
__global__ void __launch_bounds__(THREADS)
pool_ln_fwd_variants_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const int* __restrict__ segs,
                            float* __restrict__ vec, int B, int E, int L, int D) {
  extern __shared__ float red[];
  const int b = blockIdx.x, e = blockIdx.y, v = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const int s0 = segs[(v * E + e) * 2], s1 = segs[(v * E + e) * 2 + 1];
  float4 acc[NV];
  zero_acc<NV>(acc);
  for (int l = s0 + warp; l < s1; l += WARPS) {
    const size_t row = static_cast<size_t>(l) * gridDim.x + b;   // position-major rows (l, b)
    RowRegs<NV> r;
    r.load(x + row * D, nvec, lane);
    float mean, rstd;
    row_stats<NV>(r, nvec, lane, D, mean, rstd);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c);
        const float4 bb = *reinterpret_cast<const float4*>(beta + 4 * c);
        acc[i].x += (r.v[i].x - mean) * rstd * g.x + bb.x;
        acc[i].y += (r.v[i].y - mean) * rstd * g.y + bb.y;
        acc[i].z += (r.v[i].z - mean) * rstd * g.z + bb.z;
        acc[i].w += (r.v[i].w - mean) * rstd * g.w + bb.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) *reinterpret_cast<float4*>(red + warp * D + 4 * c) = acc[i];
  }
  __syncthreads();
  const float inv = 1.0f / static_cast<float>(max(1, s1 - s0));
  float* out = vec + ((static_cast<size_t>(v) * B + b) * E + e) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[w * D + c];
    out[c] = s * inv;
  }
}

template <int NV>
__global__ void __launch_bounds__(THREADS)
pool_ln_bwd_kernel(const float* __restrict__ dvec, const float* __restrict__ x,
                   const float* __restrict__ mean, const float* __restrict__ rstd,
                   const float* __restrict__ gamma, HeadSegments seg, float* __restrict__ dx,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, int L, int D) {
  extern __shared__ float red[];
  const int b = blockIdx.x, e = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const int s0 = seg.seg_begin[e], s1 = seg.seg_end[e];
  const float inv = 1.0f / static_cast<float>(max(1, s1 - s0));
  float4 acc_g[NV], acc_b[NV], g[NV], dy[NV];
  zero_acc<NV>(acc_g);
  zero_acc<NV>(acc_b);
  const float* dv = dvec + (static_cast<size_t>(b) * seg.E + e) * D;
  float s1sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      g[i] = *reinterpret_cast<const float4*>(gamma + 4 * c);
      float4 d = *reinterpret_cast<const float4*>(dv + 4 * c);
      d.x *= inv; d.y *= inv; d.z *= inv; d.w *= inv;
      dy[i] = d;
      s1sum += (d.x * g[i].x + d.y * g[i].y) + (d.z * g[i].z + d.w * g[i].w);
    } else {
      g[i] = make_float4(0, 0, 0, 0);
      dy[i] = make_float4(0, 0, 0, 0);
    }
  }
  s1sum = warp_sum(s1sum) / D;
  for (int l = s0 + warp; l < s1; l += WARPS) {
    const size_t row = static_cast<size_t>(l) * gridDim.x + b;   // position-major rows (l, b)
    RowRegs<NV> rx;
    rx.load(x + row * D, nvec, lane);
    const float mu = mean[row], rs = rstd[row];
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        float4& xv = rx.v[i];
        xv.x = (xv.x - mu) * rs; xv.y = (xv.y - mu) * rs;
        xv.z = (xv.z - mu) * rs; xv.w = (xv.w - mu) * rs;
        const float4 d = dy[i];
        acc_g[i].x += d.x * xv.x; acc_g[i].y += d.y * xv.y;
        acc_g[i].z += d.z * xv.z; acc_g[i].w += d.w * xv.w;
        acc_b[i].x += d.x; acc_b[i].y += d.y; acc_b[i].z += d.z; acc_b[i].w += d.w;
        s2 += (d.x * g[i].x * xv.x + d.y * g[i].y * xv.y) +
              (d.z * g[i].z * xv.z + d.w * g[i].w * xv.w);
      }
    }
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 d = dy[i], xv = rx.v[i];
        float4 o;
        o.x = rs * (d.x * g[i].x - s1sum - xv.x * s2);
        o.y = rs * (d.y * g[i].y - s1sum - xv.y * s2);
        o.z = rs * (d.z * g[i].z - s1sum - xv.z * s2);
        o.w = rs * (d.w * g[i].w - s1sum - xv.w * s2);
        *reinterpret_cast<float4*>(dx + row * D + 4 * c) = o;
      }
    }
  }
  block_colreduce<NV>(acc_g, red, dgamma, nvec, D);
  block_colreduce<NV>(acc_b, red, dbeta, nvec, D);
}

// One block = one head x HEADS_TB samples x a slice of the classes: every weight row is read once
// per block and reused against all HEADS_TB pooled vectors held in shared memory.  The op is tiny
// (B*E*C*D MACs) but latency bound: the class range is split over blockIdx.z and each warp keeps
// the weight rows of TWO classes in flight (12 independent 16-byte loads per lane) so the few
// blocks that exist expose enough memory-level parallelism to pull the weights at speed.
constexpr int HEADS_TB = 8;
constexpr int HEADS_ZS = 4;
__global__ void __launch_bounds__(256)
heads_fwd_kernel(const float* __restrict__ vec, HeadParams hp, float* __restrict__ logits, int B,
                 int E, int C, int D) {
  extern __shared__ __align__(16) float sv[];  // [HEADS_TB][D]
  const int b0 = blockIdx.x * HEADS_TB, e = blockIdx.y;
  const int nb = min(HEADS_TB, B - b0);
  for (int i = threadIdx.x * 4; i < HEADS_TB * D; i += blockDim.x * 4) {
    const int t = i / D, d = i % D;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < nb) v = *reinterpret_cast<const float4*>(vec + (static_cast<size_t>(b0 + t) * E + e) * D + d);
    *reinterpret_cast<float4*>(sv + i) = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* W = hp.w[e];
  const int cper = (C + gridDim.z - 1) / gridDim.z;
  const int c_begin = blockIdx.z * cper, c_end = min(C, c_begin + cper);
  for (int c = c_begin + 2 * warp; c < c_end; c += 2 * nw) {
    const bool two = c + 1 < c_end;
    const float* w0 = W + static_cast<size_t>(c) * D;
    const float* w1 = W + static_cast<size_t>(two ? c + 1 : c) * D;
    float acc0[HEADS_TB], acc1[HEADS_TB];
#pragma unroll
    for (int t = 0; t < HEADS_TB; ++t) { acc0[t] = 0.f; acc1[t] = 0.f; }
    for (int i0 = lane * 4; i0 < D; i0 += 128 * 3) {
      float4 a[3], b[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = i0 + 128 * u;
        a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < D) {
          a[u] = *reinterpret_cast<const float4*>(w0 + i);
          b[u] = *reinterpret_cast<const float4*>(w1 + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int i = i0 + 128 * u;
        if (i < D) {
#pragma unroll
          for (int t = 0; t < HEADS_TB; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(sv + t * D + i);
            acc0[t] += (a[u].x * v.x + a[u].y * v.y) + (a[u].z * v.z + a[u].w * v.w);
            acc1[t] += (b[u].x * v.x + b[u].y * v.y) + (b[u].z * v.z + b[u].w * v.w);
          }
        }
      }
    }
    const float bias0 = hp.b[e][c], bias1 = hp.b[e][two ? c + 1 : c];
#pragma unroll
    for (int t = 0; t < HEADS_TB; ++t) {
      const float s0 = warp_sum(acc0[t]), s1 = warp_sum(acc1[t]);
      if (lane == 0 && t < nb) {
        float* o = logits + (static_cast<size_t>(b0 + t) * E + e) * C + c;
        o[0] = s0 + bias0;
        if (two) o[1] = s1 + bias1;
      }
    }
  }
}

__global__ void heads_bwd_dvec_kernel(const float* __restrict__ dlogits, HeadParams hp,
                                      float* __restrict__ dvec, int E, int C, int D) {
  extern __shared__ float sdl[];
  const int b = blockIdx.x, e = blockIdx.y;
  const float* dl = dlogits + (static_cast<size_t>(b) * E + e) * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sdl[i] = dl[i];
  __syncthreads();
  const float* W = hp.w[e];
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < C; ++c) {
      const float4 w = *reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * D + d);
      const float g = sdl[c];
      s.x += g * w.x; s.y += g * w.y; s.z += g * w.z; s.w += g * w.w;
    }
    *reinterpret_cast<float4*>(dvec + (static_cast<size_t>(b) * E + e) * D + d) = s;
  }
}

__global__ void heads_bwd_dw_kernel(const float* __restrict__ dlogits, const float* __restrict__ vec,
                                    HeadParams hp, int B, int E, int C, int D) {
  const int c = blockIdx.x, e = blockIdx.y;
  float bsum = 0.f;
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float g = dlogits[(static_cast<size_t>(b) * E + e) * C + c];
      const float4 v = *reinterpret_cast<const float4*>(vec + (static_cast<size_t>(b) * E + e) * D + d);
      s.x += g * v.x; s.y += g * v.y; s.z += g * v.z; s.w += g * v.w;
    }
    float* dw = hp.dw[e] + static_cast<size_t>(c) * D + d;
    float4 prev = *reinterpret_cast<float4*>(dw);
    prev.x += s.x; prev.y += s.y; prev.z += s.z; prev.w += s.w;
    *reinterpret_cast<float4*>(dw) = prev;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < B; ++b) bsum += dlogits[(static_cast<size_t>(b) * E + e) * C + c];
    hp.db[e][c] += bsum;
  }
}

__global__ void cls_fill_kernel(const float* __restrict__ emb, float* __restrict__ mm_x, int B, int L,
                                int D, int E) {
  const size_t total = static_cast<size_t>(B) * E * D;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const int e = static_cast<int>((i / D) % E);
    const int b = static_cast<int>(i / (static_cast<size_t>(D) * E));
    mm_x[(static_cast<size_t>(e) * B + b) * D + d] = emb[static_cast<size_t>(d) * E + e];
  }
}

__global__ void cls_bwd_kernel(const float* __restrict__ dmm, float* __restrict__ demb, int B, int L,
                               int D, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * E) return;
  const int d = i % D, e = i / D;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dmm[(static_cast<size_t>(e) * B + b) * D + d];
  demb[static_cast<size_t>(d) * E + e] += s;
}

template <typename T>
__global__ void split_rows_kernel(const float* __restrict__ dmm, T* __restrict__ dimg,
                                  T* __restrict__ dtxt, int B, int L, int off_img, int l_img,
                                  int l_txt, int D) {
  const int dv = D >> 2;
  const int ltot = l_img + l_txt;
  const size_t total = static_cast<size_t>(B) * ltot * dv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // position-major everywhere: rows (off_img + l, b) of dmm -> rows (l, b) of the modality buffers
    const int c = static_cast<int>(i % dv);
    const size_t r = i / dv;
    const int b = static_cast<int>(r % B);
    const int l = static_cast<int>(r / B);
    const float4 v =
        *reinterpret_cast<const float4*>(dmm + (static_cast<size_t>(off_img + l) * B + b) * D + 4 * c);
    if (l < l_img) Vec4<T>::st(dimg + (static_cast<size_t>(l) * B + b) * D + 4 * c, v);
    else Vec4<T>::st(dtxt + (static_cast<size_t>(l - l_img) * B + b) * D + 4 * c, v);
  }
}

int nv_for(int D) {
  if (D % 4 != 0 || D <= 0 || D > 1024) return -1;
  const int need = (D / 4 + 31) / 32;
  if (need <= 1) return 1;
  if (need <= 2) return 2;
  if (need <= 4) return 4;
  if (need <= 6) return 6;
  return 8;
}

}  // namespace

// ======================================================================= launchers
int cast_gather(const void* src, void* dst, int dst_dtype, int B, int l_src, int d, const int* idx,
                int n_sel, const int* keep, int modality, cudaStream_t stream, int pos_major, int src_dtype) {
  if (d % 4 != 0) return MMU_ERR_SHAPE;
  if (B <= 0 || n_sel <= 0) return 0;
  const size_t total = static_cast<size_t>(B) * n_sel * (d / 4);
  const int grid = grid_for(total, 256);
  using bf = __nv_bfloat16;
  if (src_dtype == DT_BF16) {  // host staging in bf16 (half the H2D bytes; the bf16 stem rounds anyway)
    if (dst_dtype == DT_BF16)
      cast_gather_kernel<bf, bf><<<grid, 256, 0, stream>>>(static_cast<const bf*>(src), static_cast<bf*>(dst), B,
                                                           l_src, d, idx, n_sel, keep, modality, pos_major);
    else
      cast_gather_kernel<float, bf><<<grid, 256, 0, stream>>>(static_cast<const bf*>(src),
                                                              static_cast<float*>(dst), B, l_src, d, idx,
                                                              n_sel, keep, modality, pos_major);
  } else if (dst_dtype == DT_BF16) {
    cast_gather_kernel<bf><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), static_cast<bf*>(dst), B,
                                                     l_src, d, idx, n_sel, keep, modality, pos_major);
  } else {
    cast_gather_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src),
                                                        static_cast<float*>(dst), B, l_src, d, idx, n_sel,
                                                        keep, modality, pos_major);
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int modality_keep_mask(const float* u, const float* r, const float* score_img, const float* score_txt,
                       int score_stride, int B, float p_drop, int mode, int* keep, cudaStream_t stream) {
  if (u == nullptr || keep == nullptr || (mode != 0 && mode != 1)) return MMU_ERR_ARG;
  if (mode == 0 && r == nullptr) return MMU_ERR_ARG;
  if (mode == 1 && (score_img == nullptr || score_txt == nullptr)) return MMU_ERR_ARG;
  if (B <= 0) return 0;
  modality_keep_mask_kernel<<<(B + 127) / 128, 128, 0, stream>>>(u, r, score_img, score_txt, score_stride,
                                                                 B, p_drop, mode, keep);
  MMU_CHECK_LAUNCH();
  return 0;
}

// Ragged -> padded: packed rows [offsets[b], offsets[b+1]) of sample b land at out[b, 0..len),
// the tail of every sample is zero-filled (torch pad_sequence(batch_first=True, padding_value=0)).
__global__ void ragged_pad_kernel(const float* __restrict__ packed, const int* __restrict__ offsets,
                                  float* __restrict__ out, int B, int max_l, int d) {
  const int dv = d >> 2;
  const size_t total = static_cast<size_t>(B) * max_l * dv;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % dv);
    const size_t row = i / dv;
    const int l = static_cast<int>(row % max_l);
    const int b = static_cast<int>(row / max_l);
    const int o0 = offsets[b], len = offsets[b + 1] - o0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < len) v = *reinterpret_cast<const float4*>(packed + (static_cast<size_t>(o0) + l) * d + 4 * c);
    *reinterpret_cast<float4*>(out + row * d + 4 * c) = v;
  }
}

int ragged_pad(const float* packed, const int* offsets, float* out, int B, int max_l, int d,
               cudaStream_t stream) {
  if (d % 4 != 0) return MMU_ERR_SHAPE;
  if (B <= 0 || max_l <= 0) return 0;
  const size_t total = static_cast<size_t>(B) * max_l * (d / 4);
  ragged_pad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(packed, offsets, out, B, max_l, d);
  MMU_CHECK_LAUNCH();
  return 0;
}

int cast_f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t stream) {
  if (n % 4 != 0) return MMU_ERR_SHAPE;
  return cast_gather(src, dst, DT_BF16, 1, static_cast<int>(n / 4), 4, nullptr,
                     static_cast<int>(n / 4), nullptr, 0, stream);
}

#define MMU_NV_DISPATCH(nv, CALL) \
  switch (nv) {                   \
    case 1: { constexpr int NV = 1; CALL; } break; \
    case 2: { constexpr int NV = 2; CALL; } break; \
    case 4: { constexpr int NV = 4; CALL; } break; \
    case 6: { constexpr int NV = 6; CALL; } break; \
    case 8: { constexpr int NV = 8; CALL; } break; \
    default: return MMU_ERR_SHAPE;                   \
  }

int layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, int y_dtype,
                  float* mean, float* rstd, int M, int D, cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  const int grid = grid_for(M, WARPS);
  if (y_dtype == DT_BF16) {
    MMU_NV_DISPATCH(nv, (layernorm_fwd_kernel<NV, __nv_bfloat16><<<grid, THREADS, 0, stream>>>(
                            x, gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd, M, D)));
  } else {
    MMU_NV_DISPATCH(nv, (layernorm_fwd_kernel<NV, float><<<grid, THREADS, 0, stream>>>(
                            x, gamma, beta, static_cast<float*>(y), mean, rstd, M, D)));
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int layernorm2_fwd(const float* x, const float* g1, const float* b1, float* y1, float* mean1,
                   float* rstd1, const float* g2, const float* b2, void* y2, int y2_dtype,
                   float* mean2, float* rstd2, int M, int D, cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  const int grid = grid_for(M, WARPS);
  if (y2_dtype == DT_BF16) {
    MMU_NV_DISPATCH(nv, (layernorm2_fwd_kernel<NV, __nv_bfloat16><<<grid, THREADS, 0, stream>>>(
                            x, g1, b1, y1, mean1, rstd1, g2, b2, static_cast<__nv_bfloat16*>(y2),
                            mean2, rstd2, M, D)));
  } else {
    MMU_NV_DISPATCH(nv, (layernorm2_fwd_kernel<NV, float><<<grid, THREADS, 0, stream>>>(
                            x, g1, b1, y1, mean1, rstd1, g2, b2, static_cast<float*>(y2), mean2,
                            rstd2, M, D)));
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int layernorm_raw_stats_fwd(const float* x, const float* gamma, const float* beta, float* y,
                            void* yraw_bf16, float* stats, int nt, int M, int D, cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0 || nt < 1 || nt > 32) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  const int grid = grid_for(M, WARPS);
  MMU_NV_DISPATCH(nv, (layernorm_raw_stats_kernel<NV><<<grid, THREADS, 0, stream>>>(
                          x, gamma, beta, y, static_cast<__nv_bfloat16*>(yraw_bf16), stats, nt, M, D)));
  MMU_CHECK_LAUNCH();
  return 0;
}

int ln_fold_weights(const float* W0, const float* gamma0, const float* beta0, const float* bias0,
                    void* Wf0, float* cw0, float* bf0, int N0, const float* W1, const float* gamma1,
                    const float* beta1, const float* bias1, void* Wf1, float* cw1, float* bf1, int N1,
                    int K, cudaStream_t stream, int n_layers, long long pstride, long long wstride) {
  if (K % 4 != 0 || N0 < 1 || N1 < 0 || n_layers < 1) return MMU_ERR_SHAPE;
  FoldSet s0{W0, gamma0, beta0, bias0, static_cast<__nv_bfloat16*>(Wf0), cw0, bf0, N0};
  FoldSet s1{W1, gamma1, beta1, bias1, static_cast<__nv_bfloat16*>(Wf1), cw1, bf1, N1};
  const int nmax = N0 > N1 ? N0 : N1;
  dim3 grid((nmax + 7) / 8, N1 > 0 ? 2 : 1, n_layers);
  ln_fold_weights_kernel<<<grid, 256, 0, stream>>>(s0, s1, K, pstride, wstride);
  MMU_CHECK_LAUNCH();
  return 0;
}

int add_layernorm_fwd(const float* x_in, const void* y, float* x_out, const float* gamma,
                      const float* beta, void* h, int dtype, float* mean, float* rstd, int M, int D,
                      cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  const int grid = grid_for(M, WARPS);
  if (dtype == DT_BF16) {
    using T = __nv_bfloat16;
    MMU_NV_DISPATCH(nv, (add_layernorm_fwd_kernel<NV, T, T><<<grid, THREADS, 0, stream>>>(
                            x_in, static_cast<const T*>(y), x_out, gamma, beta, static_cast<T*>(h),
                            mean, rstd, M, D)));
  } else {
    MMU_NV_DISPATCH(nv, (add_layernorm_fwd_kernel<NV, float, float><<<grid, THREADS, 0, stream>>>(
                            x_in, static_cast<const float*>(y), x_out, gamma, beta,
                            static_cast<float*>(h), mean, rstd, M, D)));
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int layernorm_bwd(const void* dy, int dy_dtype, const float* x, const float* mean,
                  const float* rstd, const float* gamma, float* dx, int accumulate, void* dx_lp,
                  int lp_dtype, float* dgamma, float* dbeta, float* dcolsum, int M, int D,
                  cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  int grid = (M + LNB_WARPS - 1) / LNB_WARPS;
  if (grid > sm_count() * 3) grid = sm_count() * 3;
  const size_t smem = static_cast<size_t>(LNB_WARPS) * D * sizeof(float);
  using bf = __nv_bfloat16;
  if (dy_dtype == DT_BF16 && (dx_lp == nullptr || lp_dtype == DT_BF16)) {
    MMU_NV_DISPATCH(nv, (layernorm_bwd_kernel<NV, bf, bf><<<grid, LNB_THREADS, smem, stream>>>(
                            static_cast<const bf*>(dy), x, mean, rstd, gamma, dx, accumulate,
                            static_cast<bf*>(dx_lp), dgamma, dbeta, dcolsum, M, D)));
  } else if (dy_dtype == DT_F32 && (dx_lp == nullptr || lp_dtype == DT_F32)) {
    // fp32 path: the "low precision" copy would be identical to dx itself
    MMU_NV_DISPATCH(nv, (layernorm_bwd_kernel<NV, float, float><<<grid, LNB_THREADS, smem, stream>>>(
                            static_cast<const float*>(dy), x, mean, rstd, gamma, dx, accumulate,
                            static_cast<float*>(dx_lp), dgamma, dbeta, dcolsum, M, D)));
  } else if (dy_dtype == DT_F32 && lp_dtype == DT_BF16) {
    MMU_NV_DISPATCH(nv, (layernorm_bwd_kernel<NV, float, bf><<<grid, LNB_THREADS, smem, stream>>>(
                            static_cast<const float*>(dy), x, mean, rstd, gamma, dx, accumulate,
                            static_cast<bf*>(dx_lp), dgamma, dbeta, dcolsum, M, D)));
  } else {
    return MMU_ERR_ARG;
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int postln_fwd(const void* x_in, const void* y, float* s_out, const float* gamma, const float* beta,
               void* h, int dtype, float* mean, float* rstd, int M, int D, float eps,
               cudaStream_t stream, dropout::Site ydrop) {
  const int nv = nv_for(D);
  if (nv < 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  const int grid = grid_for(M, WARPS);
  if (dtype == DT_BF16) {
    using T = __nv_bfloat16;
    MMU_NV_DISPATCH(nv, (postln_fwd_kernel<NV, T><<<grid, THREADS, 0, stream>>>(
                            static_cast<const T*>(x_in), static_cast<const T*>(y), s_out, gamma, beta,
                            static_cast<T*>(h), mean, rstd, M, D, eps, ydrop)));
  } else {
    MMU_NV_DISPATCH(nv, (postln_fwd_kernel<NV, float><<<grid, THREADS, 0, stream>>>(
                            static_cast<const float*>(x_in), static_cast<const float*>(y), s_out, gamma,
                            beta, static_cast<float*>(h), mean, rstd, M, D, eps, ydrop)));
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int postln_bwd(const void* dy_branch, const float* dy_res, int dtype, const float* x, const float* mean,
               const float* rstd, const float* gamma, float* dx, void* dx_lp, float* dgamma,
               float* dbeta, float* dcolsum, int M, int D, cudaStream_t stream, PostLnDropout dr) {
  const int nv = nv_for(D);
  if (nv < 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  int grid = (M + LNB_WARPS - 1) / LNB_WARPS;
  if (grid > sm_count() * 3) grid = sm_count() * 3;
  const size_t smem = static_cast<size_t>(LNB_WARPS) * D * sizeof(float);
  if (dtype == DT_BF16) {
    using T = __nv_bfloat16;
    MMU_NV_DISPATCH(nv, (postln_bwd_kernel<NV, T><<<grid, LNB_THREADS, smem, stream>>>(
                            static_cast<const T*>(dy_branch), dy_res, x, mean, rstd, gamma, dx,
                            static_cast<T*>(dx_lp), dgamma, dbeta, dcolsum, M, D, dr)));
  } else {
    MMU_NV_DISPATCH(nv, (postln_bwd_kernel<NV, float><<<grid, LNB_THREADS, smem, stream>>>(
                            static_cast<const float*>(dy_branch), dy_res, x, mean, rstd, gamma, dx,
                            static_cast<float*>(dx_lp), dgamma, dbeta, dcolsum, M, D, dr)));
  }
  MMU_CHECK_LAUNCH();
  return 0;
}

int colsum_accumulate(const void* x, int dtype, float* out, int M, int N, cudaStream_t stream) {
  if (N % 4 != 0) return MMU_ERR_SHAPE;
  if (M <= 0) return 0;
  const int threads = 128;
  const int gx = (N / 4 + threads - 1) / threads;
  int gy = (sm_count() * 8 + gx - 1) / gx;
  if (gy > M) gy = M;
  const int rpb = (M + gy - 1) / gy;
  gy = (M + rpb - 1) / rpb;
  dim3 grid(gx, gy);
  if (dtype == DT_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, threads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), out, M, N, rpb);
  else
    colsum_kernel<float><<<grid, threads, 0, stream>>>(static_cast<const float*>(x), out, M, N, rpb);
  MMU_CHECK_LAUNCH();
  return 0;
}

int pool_ln_fwd(const float* x, const float* gamma, const float* beta, const HeadSegments& seg,
                float* vec, float* mean, float* rstd, int B, int L, int D, cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0 || seg.E < 1 || seg.E > 16) return MMU_ERR_SHAPE;
  const size_t smem = static_cast<size_t>(WARPS) * D * sizeof(float);
  dim3 grid(B, seg.E);
  MMU_NV_DISPATCH(nv, (pool_ln_fwd_kernel<NV><<<grid, THREADS, smem, stream>>>(
                          x, gamma, beta, seg, vec, mean, rstd, L, D)));
  MMU_CHECK_LAUNCH();
  return 0;
}

int pool_ln_fwd_variants(const float* x, const float* gamma, const float* beta, const int* segs,
                         int V, int E, float* vec, int B, int L, int D, cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0 || E < 1 || E > 16 || V < 1 || V > 65535 || segs == nullptr) return MMU_ERR_SHAPE;
  const size_t smem = static_cast<size_t>(WARPS) * D * sizeof(float);
  dim3 grid(B, E, V);
  MMU_NV_DISPATCH(nv, (pool_ln_fwd_variants_kernel<NV><<<grid, THREADS, smem, stream>>>(
                          x, gamma, beta, segs, vec, B, E, L, D)));
  MMU_CHECK_LAUNCH();
  return 0;
}

int pool_ln_bwd(const float* dvec, const float* x, const float* mean, const float* rstd,
                const float* gamma, const HeadSegments& seg, float* dx, float* dgamma, float* dbeta,
                int B, int L, int D, cudaStream_t stream) {
  const int nv = nv_for(D);
  if (nv < 0 || seg.E < 1 || seg.E > 16) return MMU_ERR_SHAPE;
  const size_t smem = static_cast<size_t>(WARPS) * D * sizeof(float);
  dim3 grid(B, seg.E);
  MMU_NV_DISPATCH(nv, (pool_ln_bwd_kernel<NV><<<grid, THREADS, smem, stream>>>(
                          dvec, x, mean, rstd, gamma, seg, dx, dgamma, dbeta, L, D)));
  MMU_CHECK_LAUNCH();
  return 0;
}

int heads_fwd(const float* vec, const HeadParams& hp, float* logits, int B, int E, int C, int D,
              cudaStream_t stream) {
  if (D % 4 != 0 || E > 16) return MMU_ERR_SHAPE;
  heads_fwd_kernel<<<dim3((B + HEADS_TB - 1) / HEADS_TB, E, HEADS_ZS), 256,
                     HEADS_TB * D * sizeof(float), stream>>>(vec, hp, logits, B, E, C, D);
  MMU_CHECK_LAUNCH();
  return 0;
}

int heads_bwd(const float* dlogits, const float* vec, const HeadParams& hp, float* dvec, int B,
              int E, int C, int D, cudaStream_t stream) {
  if (D % 4 != 0 || E > 16) return MMU_ERR_SHAPE;
  heads_bwd_dvec_kernel<<<dim3(B, E), 128, C * sizeof(float), stream>>>(dlogits, hp, dvec, E, C, D);
  MMU_CHECK_LAUNCH();
  heads_bwd_dw_kernel<<<dim3(C, E), 128, 0, stream>>>(dlogits, vec, hp, B, E, C, D);
  MMU_CHECK_LAUNCH();
  return 0;
}

int cls_fill(const float* class_emb, float* mm_x, int B, int L, int D, int E, cudaStream_t stream) {
  cls_fill_kernel<<<grid_for(static_cast<size_t>(B) * E * D, 256), 256, 0, stream>>>(class_emb, mm_x,
                                                                                  B, L, D, E);
  MMU_CHECK_LAUNCH();
  return 0;
}

int cls_bwd(const float* dmm, float* dclass_emb, int B, int L, int D, int E, cudaStream_t stream) {
  cls_bwd_kernel<<<(D * E + 127) / 128, 128, 0, stream>>>(dmm, dclass_emb, B, L, D, E);
  MMU_CHECK_LAUNCH();
  return 0;
}

int split_rows(const float* dmm, void* dimg, void* dtxt, int dtype, int B, int L, int off_img,
               int l_img, int l_txt, int D, cudaStream_t stream) {
  if (D % 4 != 0) return MMU_ERR_SHAPE;
  if (l_img + l_txt <= 0) return 0;
  const size_t total = static_cast<size_t>(B) * (l_img + l_txt) * (D / 4);
  const int grid = grid_for(total, 256);
  if (dtype == DT_BF16)
    split_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        dmm, static_cast<__nv_bfloat16*>(dimg), static_cast<__nv_bfloat16*>(dtxt), B, L, off_img,
        l_img, l_txt, D);
  else
    split_rows_kernel<float><<<grid, 256, 0, stream>>>(dmm, static_cast<float*>(dimg),
                                                       static_cast<float*>(dtxt), B, L, off_img,
                                                       l_img, l_txt, D);
  MMU_CHECK_LAUNCH();
  return 0;
}

}  // namespace mmu
