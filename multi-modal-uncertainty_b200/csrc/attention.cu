// Batch-axis multi-head attention, forward and backward (SIMT fp32 math, fp32 or bf16 storage).
//
// The reference feeds (B, L, D) tensors to nn.MultiheadAttention(batch_first=False)
// (src/model.py:193, :205-207, :273-276), so the *mini-batch* is the attended axis: for every
// token position l and head h there is one independent B x B attention problem over the rows
// (b, l), b = 0..B-1.  Row (b, l) of the packed q|k|v buffer lives at
// qkv[(b*L + l) * 3D ...]; head h uses columns [h*hd, (h+1)*hd) of each third.
//
// Three kernels, all built from two shared-memory tile products:
//   nt:  R[T x B]  = X[T x hd] * Y[B x hd]^T     (reduce over the head dimension)
//   nn:  Z[T x hd] = R[T x B]  * Y[B x hd]       (reduce over the batch axis)
// forward      : S = nt(Q_tile, K) -> softmax rows (saves log-sum-exp) -> O = nn(P, V)
// backward dQ  : P = exp(S - lse); dP = nt(dO_tile, V); dS = P*(dP - delta); dQ = nn(dS, K)
// backward dKV : P^T = exp(nt(K_tile, Q) - lse); dS^T likewise; dV = nn(P^T, dO); dK = nn(dS^T, Q)
// with delta[q] = <dO[q], O[q]>.  B <= 256 (the reference trains with B <= 256).
#include <cuda_bf16.h>

#include <cstdio>

#include "common.h"
#include "gemm_api.h"
#include "kernels.h"

namespace mmu {
namespace attn {

constexpr int THREADS = 256;
constexpr int TQ = 64;      // rows of the tile owned by one CTA
constexpr int BMAX = 256;   // max attended rows (mini-batch size)
constexpr int KC = 32;      // reduction chunk of the nt product
constexpr int NC = 64;      // output-column chunk of the nn product
constexpr int XS_LD = TQ + 4;
constexpr int YS_LD = BMAX + 4;

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 pk = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

// Rows of a (B, L, *) tensor that belong to token position l: row index b -> (b*L + l).
template <typename T>
struct RowView {
  const T* base;  // already offset to the head's first column
  long long ld;   // elements between consecutive (b, l) rows of the flat [B*L, ld] buffer
  int L, l;
  __device__ __forceinline__ const T* row(int b) const {
    return base + (static_cast<long long>(b) * L + l) * ld;
  }
};

struct Smem {
  float* R0;   // [TQ][B+1]
  float* R1;   // [TQ][B+1]
  float* Xs;   // [KC][XS_LD]
  float* Ys;   // [KC][YS_LD]  (nt)  /  [KC][NC] (nn)
  float* vec;  // [TQ] + [BMAX] scratch
};

// R[t][b] = sum_d X[t0+t][d] * Y[b][d]   for t < nt_rows, b < B.   R has leading dim ldr.
template <typename T>
__device__ void tile_nt(const RowView<T>& X, int t0, int nt_rows, const RowView<T>& Y, int B, int hd,
                        float* R, int ldr, float* Xs, float* Ys) {
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  const int ncol = (B + 15) >> 4;
  for (int d0 = 0; d0 < hd; d0 += KC) {
    __syncthreads();
    {  // X chunk: TQ x KC
      const int c4 = (t & 7) * 4;
#pragma unroll
      for (int i = 0; i < TQ / 32; ++i) {
        const int r = (t >> 3) + 32 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nt_rows && d0 + c4 < hd) v = ld4<T>(X.row(t0 + r) + d0 + c4);
        Xs[(c4 + 0) * XS_LD + r] = v.x;
        Xs[(c4 + 1) * XS_LD + r] = v.y;
        Xs[(c4 + 2) * XS_LD + r] = v.z;
        Xs[(c4 + 3) * XS_LD + r] = v.w;
      }
      for (int r = (t >> 3); r < B; r += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d0 + c4 < hd) v = ld4<T>(Y.row(r) + d0 + c4);
        Ys[(c4 + 0) * YS_LD + r] = v.x;
        Ys[(c4 + 1) * YS_LD + r] = v.y;
        Ys[(c4 + 2) * YS_LD + r] = v.z;
        Ys[(c4 + 3) * YS_LD + r] = v.w;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < KC; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(Xs + k * XS_LD + ty * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < ncol) {
          const float bv = Ys[k * YS_LD + tx + 16 * j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(av[i], bv, acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = tx + 16 * j;
      if (j < ncol && c < B) R[r * ldr + c] = (r < nt_rows) ? acc[i][j] : 0.f;
    }
  }
  __syncthreads();
}

// Z[t0+t][c] = alpha * sum_b R[t][b] * Y[b][c]  -> written to out rows (t0+t) via `O` view.
template <typename T>
__device__ void tile_nn(const float* R, int ldr, int nt_rows, const RowView<T>& Y, int B, int hd,
                        T* out_base, long long out_ld, int L, int l, int t0, float alpha,
                        float* Ys) {
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  for (int c0 = 0; c0 < hd; c0 += NC) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int b0 = 0; b0 < B; b0 += KC) {
      __syncthreads();
      {  // Y chunk: KC rows x NC cols
        const int c4 = (t & 15) * 4;
#pragma unroll
        for (int i = 0; i < KC / 16; ++i) {
          const int r = (t >> 4) + 16 * i;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b0 + r < B && c0 + c4 < hd) v = ld4<T>(Y.row(b0 + r) + c0 + c4);
          *reinterpret_cast<float4*>(Ys + r * NC + c4) = v;
        }
      }
      __syncthreads();
      const int kmax = min(KC, B - b0);
      for (int k = 0; k < kmax; ++k) {
        const float4 yv = *reinterpret_cast<const float4*>(Ys + k * NC + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float rv = R[(ty * 4 + i) * ldr + b0 + k];
          acc[i][0] = fmaf(rv, yv.x, acc[i][0]);
          acc[i][1] = fmaf(rv, yv.y, acc[i][1]);
          acc[i][2] = fmaf(rv, yv.z, acc[i][2]);
          acc[i][3] = fmaf(rv, yv.w, acc[i][3]);
        }
      }
    }
    const int c = c0 + tx * 4;
    if (c < hd) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (r < nt_rows) {
          T* dst = out_base + (static_cast<long long>(t0 + r) * L + l) * out_ld + c;
          st4<T>(dst, make_float4(acc[i][0] * alpha, acc[i][1] * alpha, acc[i][2] * alpha,
                                  acc[i][3] * alpha));
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ Smem carve(float* base, int B) {
  Smem s;
  const int ldr = B + 1;
  s.R0 = base;
  s.R1 = s.R0 + TQ * ldr;
  s.Xs = s.R1 + TQ * ldr;
  s.Ys = s.Xs + KC * XS_LD;
  s.vec = s.Ys + KC * YS_LD;
  return s;
}

size_t smem_bytes(int B) {
  return sizeof(float) * (2 * TQ * (B + 1) + KC * XS_LD + KC * YS_LD + TQ + BMAX);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int B, int L_,
                int D, int H, int pos_major) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, B);
  // row of (sample b, position lp): b*L + lp (batch-major) or lp*B + b (position-major) -- both are
  // b*L + l with (L, l) = (L_, lp) resp. (1, lp*B)
  const int lh = blockIdx.x, lp = lh / H, h = lh % H;
  const int L = pos_major ? 1 : L_, l = pos_major ? lp * B : lp;
  const int hd = D / H;
  const int t0 = blockIdx.y * TQ;
  const int rows = min(TQ, B - t0);
  const int ldr = B + 1;
  const float scale = rsqrtf(static_cast<float>(hd));
  const RowView<T> Q{qkv + h * hd, 3LL * D, L, l};
  const RowView<T> K{qkv + D + h * hd, 3LL * D, L, l};
  const RowView<T> V{qkv + 2 * D + h * hd, 3LL * D, L, l};

  tile_nt<T>(Q, t0, rows, K, B, hd, s.R0, ldr, s.Xs, s.Ys);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += THREADS / 32) {
    float* row = s.R0 + r * ldr;
    float m = -INFINITY;
    for (int c = lane; c < B; c += 32) m = fmaxf(m, row[c] * scale);
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < B; c += 32) {
      const float e = expf(row[c] * scale - m);
      row[c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < B; c += 32) row[c] *= inv;
    if (lane == 0) lse[static_cast<size_t>(lh) * B + t0 + r] = m + logf(sum);
  }
  __syncthreads();
  tile_nn<T>(s.R0, ldr, rows, V, B, hd, out + h * hd, D, L, l, t0, 1.0f, s.Ys);
}

// dQ kernel; also produces delta[q] = <dO[q], O[q]> for the dKV kernel.
template <typename T>
__global__ void __launch_bounds__(THREADS)
attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                   const float* __restrict__ lse, float* __restrict__ delta, T* __restrict__ dqkv,
                   int B, int L_, int D, int H, int pos_major) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, B);
  const int lh = blockIdx.x, lp = lh / H, h = lh % H;
  const int L = pos_major ? 1 : L_, l = pos_major ? lp * B : lp;
  const int hd = D / H;
  const int t0 = blockIdx.y * TQ;
  const int rows = min(TQ, B - t0);
  const int ldr = B + 1;
  const float scale = rsqrtf(static_cast<float>(hd));
  const RowView<T> Q{qkv + h * hd, 3LL * D, L, l};
  const RowView<T> K{qkv + D + h * hd, 3LL * D, L, l};
  const RowView<T> V{qkv + 2 * D + h * hd, 3LL * D, L, l};
  const RowView<T> O{out + h * hd, static_cast<long long>(D), L, l};
  const RowView<T> dO{dout + h * hd, static_cast<long long>(D), L, l};

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += THREADS / 32) {
    const T* o = O.row(t0 + r);
    const T* g = dO.row(t0 + r);
    float acc = 0.f;
    for (int c = lane * 4; c < hd; c += 128) {
      const float4 a = ld4<T>(o + c), b = ld4<T>(g + c);
      acc += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      s.vec[r] = acc;
      delta[static_cast<size_t>(lh) * B + t0 + r] = acc;
    }
  }
  tile_nt<T>(Q, t0, rows, K, B, hd, s.R0, ldr, s.Xs, s.Ys);   // S (unscaled)
  tile_nt<T>(dO, t0, rows, V, B, hd, s.R1, ldr, s.Xs, s.Ys);  // dP
  for (int i = threadIdx.x; i < rows * B; i += THREADS) {
    const int r = i / B, c = i % B;
    const float p = expf(s.R0[r * ldr + c] * scale - lse[static_cast<size_t>(lh) * B + t0 + r]);
    s.R0[r * ldr + c] = p * (s.R1[r * ldr + c] - s.vec[r]);  // dS (without the 1/sqrt(hd) factor)
  }
  __syncthreads();
  tile_nn<T>(s.R0, ldr, rows, K, B, hd, dqkv + h * hd, 3LL * D, L, l, t0, scale, s.Ys);
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                    const float* __restrict__ lse, const float* __restrict__ delta,
                    T* __restrict__ dqkv, int B, int L_, int D, int H, int pos_major) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, B);
  const int lh = blockIdx.x, lp = lh / H, h = lh % H;
  const int L = pos_major ? 1 : L_, l = pos_major ? lp * B : lp;
  const int hd = D / H;
  const int t0 = blockIdx.y * TQ;  // key tile
  const int rows = min(TQ, B - t0);
  const int ldr = B + 1;
  const float scale = rsqrtf(static_cast<float>(hd));
  const RowView<T> Q{qkv + h * hd, 3LL * D, L, l};
  const RowView<T> K{qkv + D + h * hd, 3LL * D, L, l};
  const RowView<T> V{qkv + 2 * D + h * hd, 3LL * D, L, l};
  const RowView<T> dO{dout + h * hd, static_cast<long long>(D), L, l};
  float* lse_s = s.vec + TQ;  // [B]
  for (int i = threadIdx.x; i < B; i += THREADS) lse_s[i] = lse[static_cast<size_t>(lh) * B + i];

  tile_nt<T>(K, t0, rows, Q, B, hd, s.R0, ldr, s.Xs, s.Ys);   // S^T
  tile_nt<T>(V, t0, rows, dO, B, hd, s.R1, ldr, s.Xs, s.Ys);  // dP^T
  for (int i = threadIdx.x; i < rows * B; i += THREADS) {
    const int r = i / B, c = i % B;  // r: key, c: query
    const float p = expf(s.R0[r * ldr + c] * scale - lse_s[c]);
    s.R0[r * ldr + c] = p;
    s.R1[r * ldr + c] = p * (s.R1[r * ldr + c] - delta[static_cast<size_t>(lh) * B + c]);
  }
  __syncthreads();
  tile_nn<T>(s.R0, ldr, rows, dO, B, hd, dqkv + 2 * D + h * hd, 3LL * D, L, l, t0, 1.0f, s.Ys);
  tile_nn<T>(s.R1, ldr, rows, Q, B, hd, dqkv + D + h * hd, 3LL * D, L, l, t0, scale, s.Ys);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(bytes)) == cudaSuccess
             ? 0
             : MMU_ERR_CUDA;
}

int check(int B, int D, int H) {
  if (H <= 0 || D % H != 0) return MMU_ERR_SHAPE;
  if ((D / H) % 4 != 0) return MMU_ERR_SHAPE;
  if (B < 1 || B > BMAX) return MMU_ERR_SHAPE;
  return 0;
}


// ============================================================ tensor-core path (bf16)
// Each (token position l, head h) pair is one batch entry of the batched tcgen05 GEMM
// (3-D tensor maps over the packed qkv buffer):
//   forward : S = (Q K^T)/sqrt(hd) [GEMM, fp32 out] -> row softmax -> P (bf16, kept for the
//             backward) -> O = P V [GEMM, V as MN-major operand]
//   backward: dP = dO V^T [GEMM] -> dS = P (dP - <dP,P>) / sqrt(hd) (row kernel) ->
//             dV = P^T dO, dK = dS^T Q, dQ = dS K  [3 GEMMs, transposes via MN-major descriptors]
// S/P/dP/dS are only B x B per batch entry (23-47 MB per layer at B = 128), so keeping them in
// HBM costs far less than the qkv traffic itself.
namespace tc {

__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P, int rows, int B,
                    int Bp) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + static_cast<size_t>(row) * Bp;
  float m = -INFINITY;
  for (int c = lane; c < B; c += 32) m = fmaxf(m, s[c]);
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < B; c += 32) sum += __expf(s[c] - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* p = P + static_cast<size_t>(row) * Bp;
  for (int c = lane; c < Bp; c += 32)
    p[c] = __float2bfloat16_rn(c < B ? __expf(s[c] - m) * inv : 0.f);
}

// `drop` (attention-probability dropout, MMBT training): the incoming gradient is that of the
// DROPPED probabilities, dP = dPd * mask / (1 - p) with the mask regenerated from the counter
// row * B + column (csrc/dropout.cuh); off by default.
__global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const float* __restrict__ dP, const __nv_bfloat16* __restrict__ P,
                        __nv_bfloat16* __restrict__ dS, int rows, int B, int Bp, float scale,
                        const dropout::Site drop = dropout::Site{0u, 0u, 0u, 1.0f}) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* g = dP + static_cast<size_t>(row) * Bp;
  const __nv_bfloat16* p = P + static_cast<size_t>(row) * Bp;
  const unsigned int e0 = static_cast<unsigned int>(row) * static_cast<unsigned int>(B);
  float delta = 0.f;
  for (int c = lane; c < B; c += 32)
    delta += g[c] * (drop.on() ? drop.mult(e0 + c) : 1.0f) * __bfloat162float(p[c]);
  delta = warp_sum(delta);
  __nv_bfloat16* o = dS + static_cast<size_t>(row) * Bp;
  for (int c = lane; c < Bp; c += 32)
    o[c] = __float2bfloat16_rn(
        c < B ? __bfloat162float(p[c]) * (g[c] * (drop.on() ? drop.mult(e0 + c) : 1.0f) - delta) * scale : 0.f);
}

// Pd = P * mask / (1 - p): the dropped probabilities the P V product (and dV = Pd^T dO) consume.
__global__ void __launch_bounds__(256)
dropout_rows_kernel(const __nv_bfloat16* __restrict__ P, __nv_bfloat16* __restrict__ Pd, int rows, int n,
                    int np, const dropout::Site drop) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned int e0 = static_cast<unsigned int>(row) * static_cast<unsigned int>(n);
  const __nv_bfloat16* p = P + static_cast<size_t>(row) * np;
  __nv_bfloat16* o = Pd + static_cast<size_t>(row) * np;
  for (int c = lane; c < np; c += 32)
    o[c] = __float2bfloat16_rn(c < n ? __bfloat162float(p[c]) * drop.mult(e0 + c) : 0.f);
}

struct Views {
  BatchedOperand q_k, k_k, v_k;     // K-major views of the q / k / v thirds (rows = b, K = d)
  BatchedOperand q_mn, k_mn, v_mn;  // MN-major views (MN = d, K = b)
};

// pos: rows are (position, sample) -> l*B + b instead of (sample, position) -> b*L + l
BatchedOperand qkv_view(const void* qkv, int B, int L, int D, int H, int third, int mn, int pos) {
  BatchedOperand o{};
  o.base = qkv;
  o.inner = 3LL * D; o.mid = L; o.outer = B;
  o.mid_stride = pos ? 3LL * D * B : 3LL * D; o.outer_stride = pos ? 3LL * D : 3LL * D * L;
  o.mn_major = mn; o.hdiv = H; o.hstride = D / H; o.col0 = third * D;
  return o;
}
BatchedOperand act_view(const void* x, int B, int L, int D, int H, int mn, int pos) {  // [rows, D] tensors
  BatchedOperand o{};
  o.base = x;
  o.inner = D; o.mid = L; o.outer = B;
  o.mid_stride = pos ? static_cast<long long>(D) * B : D;
  o.outer_stride = pos ? D : static_cast<long long>(D) * L;
  o.mn_major = mn; o.hdiv = H; o.hstride = D / H; o.col0 = 0;
  return o;
}
BatchedOperand sq_view(const void* p, int G, int B, int Bp, int mn) {  // [G][B][Bp] matrices
  BatchedOperand o{};
  o.base = p;
  o.inner = Bp; o.mid = G; o.outer = B;
  o.mid_stride = static_cast<long long>(B) * Bp; o.outer_stride = Bp;
  o.mn_major = mn; o.hdiv = 1; o.hstride = 0; o.col0 = 0;
  return o;
}

GemmEpilogue store_epi(void* out, int bf16, long long ld, float alpha) {
  GemmEpilogue e{};
  e.mode = EPI_STORE; e.out_bf16 = bf16; e.out = out; e.ld_out = ld; e.alpha = alpha;
  return e;
}

bool eligible(int dtype, int B, int D, int H, const void* probs, const void* scores) {
  return dtype == DT_BF16 && probs != nullptr && scores != nullptr && (D / H) % 64 == 0 &&
         D % 8 == 0 && B >= 1;
}

int fwd(const void* qkv, void* out, void* probs, float* scores, int B, int L, int D, int H, int pos,
        cudaStream_t st) {
  const int hd = D / H, G = L * H, Bp = (B + 7) / 8 * 8;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  if (Bp <= 128) {
    // P = softmax(Q K^T / sqrt(hd)) in ONE kernel: the row softmax runs in the GEMM epilogue
    // (a lane owns a whole 128-column row of the accumulator), S never exists in HBM
    GemmEpilogue e = store_epi(probs, 1, Bp, scale);
    e.mode = EPI_SOFTMAX;
    e.n_valid = B;
    const int rc0 = gemm_bf16_batched_launch(qkv_view(qkv, B, L, D, H, 0, 0, pos),
                                             qkv_view(qkv, B, L, D, H, 1, 0, pos), G, B, Bp, hd, e, 1, 0,
                                             static_cast<long long>(B) * Bp, st);
    if (rc0) return rc0;
    return gemm_bf16_batched_launch(sq_view(probs, G, B, Bp, 0), qkv_view(qkv, B, L, D, H, 2, 1, pos), G, B,
                                    hd, B, store_epi(out, 1, pos ? D : static_cast<long long>(L) * D, 1.0f), H,
                                    hd, pos ? static_cast<long long>(B) * D : D, st);
  }
  int rc = gemm_bf16_batched_launch(qkv_view(qkv, B, L, D, H, 0, 0, pos), qkv_view(qkv, B, L, D, H, 1, 0, pos),
                                    G, B, Bp, hd, store_epi(scores, 0, Bp, scale), 1, 0,
                                    static_cast<long long>(B) * Bp, st);
  if (rc) return rc;
  const int rows = G * B;
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(scores, static_cast<__nv_bfloat16*>(probs),
                                                      rows, B, Bp);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return gemm_bf16_batched_launch(sq_view(probs, G, B, Bp, 0), qkv_view(qkv, B, L, D, H, 2, 1, pos), G, B,
                                  hd, B, store_epi(out, 1, pos ? D : static_cast<long long>(L) * D, 1.0f), H,
                                  hd, pos ? static_cast<long long>(B) * D : D, st);
}

int bwd(const void* qkv, const void* dout, const void* probs, float* scores, void* dprobs,
        void* dqkv, int B, int L, int D, int H, int pos, cudaStream_t st) {
  const int hd = D / H, G = L * H, Bp = (B + 7) / 8 * 8;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  // dP = dO V^T
  int rc = gemm_bf16_batched_launch(act_view(dout, B, L, D, H, 0, pos), qkv_view(qkv, B, L, D, H, 2, 0, pos),
                                    G, B, Bp, hd, store_epi(scores, 0, Bp, 1.0f), 1, 0,
                                    static_cast<long long>(B) * Bp, st);
  if (rc) return rc;
  const int rows = G * B;
  softmax_bwd_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(
      scores, static_cast<const __nv_bfloat16*>(probs), static_cast<__nv_bfloat16*>(dprobs), rows, B,
      Bp, scale);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
  const long long ld = pos ? 3LL * D : 3LL * D * L;          // row (sample) pitch of dqkv
  const long long lmid = pos ? 3LL * D * B : 3LL * D;        // position pitch
  // dV = P^T dO
  rc = gemm_bf16_batched_launch(sq_view(probs, G, B, Bp, 1), act_view(dout, B, L, D, H, 1, pos), G, B, hd,
                                B, store_epi(dq + 2 * D, 1, ld, 1.0f), H, hd, lmid, st);
  if (rc) return rc;
  // dK = dS^T Q
  rc = gemm_bf16_batched_launch(sq_view(dprobs, G, B, Bp, 1), qkv_view(qkv, B, L, D, H, 0, 1, pos), G, B,
                                hd, B, store_epi(dq + D, 1, ld, 1.0f), H, hd, lmid, st);
  if (rc) return rc;
  // dQ = dS K
  return gemm_bf16_batched_launch(sq_view(dprobs, G, B, Bp, 0), qkv_view(qkv, B, L, D, H, 1, 1, pos), G, B,
                                  hd, B, store_epi(dq, 1, ld, 1.0f), H, hd, lmid, st);
}


// ---------------------------------------------------------------- sequence-axis attention
// BERT self-attention of the MMBT path (reference call site src/mmbt.py:124-128; arithmetic of
// pytorch_pretrained_bert's BertSelfAttention): one S x S problem per (sample b, head h) over
// rows (b, s), with the additive key mask (1 - m)(-10000) of src/mmbt.py:103-107.  Same batched
// GEMMs as above with the roles of the batch and token axes swapped in the tensor maps.
__global__ void __launch_bounds__(256)
softmax_mask_rows_kernel(const float* __restrict__ S, const float* __restrict__ addmask,
                         __nv_bfloat16* __restrict__ P, int rows, int n, int np, int rows_per_sample,
                         __nv_bfloat16* __restrict__ Pd, const dropout::Site drop) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* s = S + static_cast<size_t>(row) * np;
  const float* am = addmask + static_cast<size_t>(row / rows_per_sample) * n;
  float m = -INFINITY;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, s[c] + am[c]);
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < n; c += 32) sum += __expf(s[c] + am[c] - m);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* p = P + static_cast<size_t>(row) * np;
  __nv_bfloat16* pd = Pd != nullptr ? Pd + static_cast<size_t>(row) * np : nullptr;
  const unsigned int e0 = static_cast<unsigned int>(row) * static_cast<unsigned int>(n);
  for (int c = lane; c < np; c += 32) {
    const float v = c < n ? __expf(s[c] + am[c] - m) * inv : 0.f;
    p[c] = __float2bfloat16_rn(v);
    // the dropped copy is derived from the ROUNDED probability, as the backward regenerates it
    if (pd != nullptr) pd[c] = __float2bfloat16_rn(c < n ? __bfloat162float(p[c]) * drop.mult(e0 + c) : 0.f);
  }
}

BatchedOperand qkv_view_seq(const void* qkv, int B, int S, int D, int H, int third, int mn) {
  BatchedOperand o{};
  o.base = qkv;
  o.inner = 3LL * D; o.mid = B; o.outer = S;
  o.mid_stride = 3LL * D * S; o.outer_stride = 3LL * D;
  o.mn_major = mn; o.hdiv = H; o.hstride = D / H; o.col0 = third * D;
  return o;
}
BatchedOperand act_view_seq(const void* x, int B, int S, int D, int H, int mn) {
  BatchedOperand o{};
  o.base = x;
  o.inner = D; o.mid = B; o.outer = S;
  o.mid_stride = static_cast<long long>(D) * S; o.outer_stride = D;
  o.mn_major = mn; o.hdiv = H; o.hstride = D / H; o.col0 = 0;
  return o;
}

int seq_fwd(const void* qkv, const float* addmask, void* out, void* probs, float* scores, int B, int S,
            int D, int H, cudaStream_t st, dropout::Site drop, void* pdrop) {
  const int hd = D / H, G = B * H, Sp = (S + 7) / 8 * 8;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  int rc = gemm_bf16_batched_launch(qkv_view_seq(qkv, B, S, D, H, 0, 0), qkv_view_seq(qkv, B, S, D, H, 1, 0),
                                    G, S, Sp, hd, store_epi(scores, 0, Sp, scale), 1, 0,
                                    static_cast<long long>(S) * Sp, st);
  if (rc) return rc;
  const int rows = G * S;
  const bool dropped = drop.on() && pdrop != nullptr;
  softmax_mask_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(
      scores, addmask, static_cast<__nv_bfloat16*>(probs), rows, S, Sp, H * S,
      dropped ? static_cast<__nv_bfloat16*>(pdrop) : nullptr, drop);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return gemm_bf16_batched_launch(sq_view(dropped ? pdrop : probs, G, S, Sp, 0), qkv_view_seq(qkv, B, S, D, H, 2, 1), G, S,
                                  hd, S, store_epi(out, 1, D, 1.0f), H, hd,
                                  static_cast<long long>(S) * D, st);
}

// With attention-probability dropout: Pd is regenerated into `dprobs` for dV = Pd^T dO FIRST, then
// dPd = dO V^T, dS = P o (dPd * mask / (1 - p) - delta) / sqrt(hd) overwrites `dprobs` -- by the
// fused kernel (dPd stays in TMEM, the mask is regenerated there, dQ = dS K comes with it) when it
// applies, else by the dP GEMM + row kernel.
int seq_bwd_dropout(const void* qkv, const void* dout, const void* probs, float* scores, void* dprobs,
                    void* dqkv, int B, int S, int D, int H, cudaStream_t st, dropout::Site drop,
                    const void* pdrop_saved) {
  const int hd = D / H, G = B * H, Sp = (S + 7) / 8 * 8;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  const int rows = G * S;
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
  const long long ld = 3LL * D, mid = 3LL * D * S;
  const void* pd = pdrop_saved;
  if (pd == nullptr) {  // the forward's dropped copy was not kept: regenerate it
    dropout_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(probs),
                                                       static_cast<__nv_bfloat16*>(dprobs), rows, S, Sp, drop);
    if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
    count_launch();
    pd = dprobs;
  }
  int rc = gemm_bf16_batched_launch(sq_view(pd, G, S, Sp, 1), act_view_seq(dout, B, S, D, H, 1), G, S, hd,
                                    S, store_epi(dq + 2 * D, 1, ld, 1.0f), H, hd, mid, st);
  if (rc) return rc;
  rc = fused_seq_attention_bwd_ds(qkv, dout, probs, dprobs, dqkv, B, S, D, H, st, drop);
  if (rc < 0) return rc;
  const bool dq_done = rc == 0;  // the fused kernel also produced dQ = dS K
  if (rc > 0) {
    rc = gemm_bf16_batched_launch(act_view_seq(dout, B, S, D, H, 0), qkv_view_seq(qkv, B, S, D, H, 2, 0),
                                  G, S, Sp, hd, store_epi(scores, 0, Sp, 1.0f), 1, 0,
                                  static_cast<long long>(S) * Sp, st);
    if (rc) return rc;
    softmax_bwd_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(
        scores, static_cast<const __nv_bfloat16*>(probs), static_cast<__nv_bfloat16*>(dprobs), rows, S, Sp,
        scale, drop);
    if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
    count_launch();
  }
  rc = gemm_bf16_batched_launch(sq_view(dprobs, G, S, Sp, 1), qkv_view_seq(qkv, B, S, D, H, 0, 1), G, S,
                                hd, S, store_epi(dq + D, 1, ld, 1.0f), H, hd, mid, st);
  if (rc || dq_done) return rc;
  return gemm_bf16_batched_launch(sq_view(dprobs, G, S, Sp, 0), qkv_view_seq(qkv, B, S, D, H, 1, 1), G, S,
                                  hd, S, store_epi(dq, 1, ld, 1.0f), H, hd, mid, st);
}

int seq_bwd(const void* qkv, const void* dout, const void* probs, float* scores, void* dprobs,
            void* dqkv, int B, int S, int D, int H, cudaStream_t st) {
  const int hd = D / H, G = B * H, Sp = (S + 7) / 8 * 8;
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  // dS = P o (dP - delta) / sqrt(hd), dP = dO V^T: one fused kernel (dP stays in TMEM) when it
  // applies, else the dP GEMM (fp32 to HBM) + the row kernel
  int rc = fused_seq_attention_bwd_ds(qkv, dout, probs, dprobs, dqkv, B, S, D, H, st);
  if (rc < 0) return rc;
  const bool dq_done = rc == 0;  // the fused kernel also produced dQ = dS K
  if (rc > 0) {
    rc = gemm_bf16_batched_launch(act_view_seq(dout, B, S, D, H, 0), qkv_view_seq(qkv, B, S, D, H, 2, 0),
                                  G, S, Sp, hd, store_epi(scores, 0, Sp, 1.0f), 1, 0,
                                  static_cast<long long>(S) * Sp, st);
    if (rc) return rc;
    const int rows = G * S;
    softmax_bwd_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(
        scores, static_cast<const __nv_bfloat16*>(probs), static_cast<__nv_bfloat16*>(dprobs), rows, S,
        Sp, scale);
    if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
    count_launch();
  }
  __nv_bfloat16* dq = static_cast<__nv_bfloat16*>(dqkv);
  const long long ld = 3LL * D, mid = 3LL * D * S;
  // dV = P^T dO ; dK = dS^T Q ; dQ = dS K
  rc = gemm_bf16_batched_launch(sq_view(probs, G, S, Sp, 1), act_view_seq(dout, B, S, D, H, 1), G, S, hd,
                                S, store_epi(dq + 2 * D, 1, ld, 1.0f), H, hd, mid, st);
  if (rc) return rc;
  rc = gemm_bf16_batched_launch(sq_view(dprobs, G, S, Sp, 1), qkv_view_seq(qkv, B, S, D, H, 0, 1), G, S,
                                hd, S, store_epi(dq + D, 1, ld, 1.0f), H, hd, mid, st);
  if (rc || dq_done) return rc;
  return gemm_bf16_batched_launch(sq_view(dprobs, G, S, Sp, 0), qkv_view_seq(qkv, B, S, D, H, 1, 1), G, S,
                                  hd, S, store_epi(dq, 1, ld, 1.0f), H, hd, mid, st);
}

}  // namespace tc

// fp32 parity path of the sequence-axis attention: one generic strided batched product
//   C[g](m, n) = alpha * sum_k A[g](m, k) * B[g](n, k),   g -> (g / H, g % H)
// (16 x 16 shared-memory tiles) covers S = Q K^T, O = P V and the four backward products.
namespace seq32 {
struct Strided {
  const float* base;
  long long s0, s1, sm, sk;  // element strides: g / H, g % H, row (m or n), reduction index
};
__global__ void __launch_bounds__(256)
bgemm_kernel(Strided A, Strided Bm, float* __restrict__ C, long long c0, long long c1, long long cm,
             int M, int N, int K, int H, float alpha) {
  __shared__ float As[16][17], Bs[16][17];
  const int g = blockIdx.z;
  const float* a = A.base + (g / H) * A.s0 + (g % H) * A.s1;
  const float* b = Bm.base + (g / H) * Bm.s0 + (g % H) * Bm.s1;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 16, n0 = blockIdx.x * 16;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    const int ka = k0 + tx;
    As[ty][tx] = (m0 + ty < M && ka < K) ? a[(m0 + ty) * A.sm + ka * A.sk] : 0.f;
    Bs[ty][tx] = (n0 + ty < N && ka < K) ? b[(n0 + ty) * Bm.sm + ka * Bm.sk] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(As[ty][k], Bs[tx][k], acc);
    __syncthreads();
  }
  if (m0 + ty < M && n0 + tx < N)
    C[(g / H) * c0 + (g % H) * c1 + (m0 + ty) * cm + n0 + tx] = alpha * acc;
}
int bgemm(const Strided& A, const Strided& B, float* C, long long c0, long long c1, long long cm, int G,
          int M, int N, int K, int H, float alpha, cudaStream_t st) {
  dim3 grid((N + 15) / 16, (M + 15) / 16, G);
  bgemm_kernel<<<grid, 256, 0, st>>>(A, B, C, c0, c1, cm, M, N, K, H, alpha);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}
__global__ void __launch_bounds__(256)
softmax_mask_kernel(float* __restrict__ P, const float* __restrict__ addmask, int rows, int n,
                    int rows_per_sample) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* s = P + static_cast<size_t>(row) * n;
  const float* am = addmask + static_cast<size_t>(row / rows_per_sample) * n;
  float m = -INFINITY;
  for (int c = lane; c < n; c += 32) m = fmaxf(m, s[c] + am[c]);
  m = warp_max(m);
  float sum = 0.f;
  for (int c = lane; c < n; c += 32) sum += expf(s[c] + am[c] - m);
  sum = warp_sum(sum);
  for (int c = lane; c < n; c += 32) s[c] = expf(s[c] + am[c] - m) / sum;
}
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(float* __restrict__ dP, const float* __restrict__ P, int rows, int n, float scale,
                   const dropout::Site drop = dropout::Site{0u, 0u, 0u, 1.0f}) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* g = dP + static_cast<size_t>(row) * n;
  const float* p = P + static_cast<size_t>(row) * n;
  const unsigned int e0 = static_cast<unsigned int>(row) * static_cast<unsigned int>(n);
  float delta = 0.f;
  for (int c = lane; c < n; c += 32) delta += g[c] * (drop.on() ? drop.mult(e0 + c) : 1.0f) * p[c];
  delta = warp_sum(delta);
  for (int c = lane; c < n; c += 32)
    g[c] = p[c] * (g[c] * (drop.on() ? drop.mult(e0 + c) : 1.0f) - delta) * scale;
}
// Pd = P * mask / (1 - p) (fp32 parity path)
__global__ void __launch_bounds__(256)
dropout_rows_kernel(const float* __restrict__ P, float* __restrict__ Pd, int rows, int n,
                    const dropout::Site drop) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const unsigned int e0 = static_cast<unsigned int>(row) * static_cast<unsigned int>(n);
  for (int c = lane; c < n; c += 32)
    Pd[static_cast<size_t>(row) * n + c] = P[static_cast<size_t>(row) * n + c] * drop.mult(e0 + c);
}
}  // namespace seq32

}  // namespace attn

int seq_attention_fwd(const void* qkv, const float* addmask, void* out, void* probs, float* scores,
                      int dtype, int B, int S, int D, int H, cudaStream_t stream, int keep_probs,
                      int allow_fused, dropout::Site drop, void* pdrop) {
  using namespace attn;
  if (qkv == nullptr || addmask == nullptr || out == nullptr || probs == nullptr) return MMU_ERR_ARG;
  if (B < 1 || S < 1 || H < 1 || D % H != 0) return MMU_ERR_SHAPE;
  if (drop.on() && pdrop == nullptr) return MMU_ERR_ARG;
  if (dtype == DT_BF16) {
    if ((D / H) % 64 != 0 || scores == nullptr) return MMU_ERR_SHAPE;
    if (allow_fused) {  // one fused kernel when it applies (head_dim 64, S <= 512)
      const int rc = fused_seq_attention_fwd(qkv, addmask, out, keep_probs ? probs : nullptr, B, S, D, H, stream,
                                             drop, pdrop);
      if (rc <= 0) return rc;
    }
    return tc::seq_fwd(qkv, addmask, out, probs, scores, B, S, D, H, stream, drop, pdrop);
  }
  using seq32::Strided;
  const int hd = D / H, G = B * H;
  const float* q = static_cast<const float*>(qkv);
  float* P = static_cast<float*>(probs);
  const long long SS = static_cast<long long>(S) * S;
  const Strided Q{q, 3LL * D * S, hd, 3LL * D, 1}, Kk{q + D, 3LL * D * S, hd, 3LL * D, 1};
  if (int rc = seq32::bgemm(Q, Kk, P, H * SS, SS, S, G, S, S, hd, H, 1.0f / sqrtf(static_cast<float>(hd)), stream))
    return rc;
  const int rows = G * S;
  seq32::softmax_mask_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(P, addmask, rows, S, H * S);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  const float* Puse = P;
  if (drop.on()) {  // the dropped copy feeds P V; the undropped P is what the backward keeps
    seq32::dropout_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(P, static_cast<float*>(pdrop), rows, S, drop);
    if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
    count_launch();
    Puse = static_cast<const float*>(pdrop);
  }
  const Strided Pm{Puse, H * SS, SS, S, 1}, V{q + 2 * D, 3LL * D * S, hd, 1, 3LL * D};
  return seq32::bgemm(Pm, V, static_cast<float*>(out), static_cast<long long>(S) * D, hd, D, G, S, hd, S, H,
                      1.0f, stream);
}

int seq_attention_bwd(const void* qkv, const void* dout, const void* probs, float* scores, void* dprobs,
                      void* dqkv, int dtype, int B, int S, int D, int H, cudaStream_t stream,
                      dropout::Site drop, const void* pdrop_saved) {
  using namespace attn;
  if (qkv == nullptr || dout == nullptr || probs == nullptr || scores == nullptr || dqkv == nullptr)
    return MMU_ERR_ARG;
  if (dtype == DT_BF16) {
    if ((D / H) % 64 != 0 || dprobs == nullptr) return MMU_ERR_SHAPE;
    if (drop.on())
      return tc::seq_bwd_dropout(qkv, dout, probs, scores, dprobs, dqkv, B, S, D, H, stream, drop, pdrop_saved);
    return tc::seq_bwd(qkv, dout, probs, scores, dprobs, dqkv, B, S, D, H, stream);
  }
  using seq32::Strided;
  const int hd = D / H, G = B * H;
  const float* q = static_cast<const float*>(qkv);
  const float* P = static_cast<const float*>(probs);
  const float* dO = static_cast<const float*>(dout);
  float* dq = static_cast<float*>(dqkv);
  float* dS = scores;  // fp32 [G][S][S] scratch
  const long long SS = static_cast<long long>(S) * S, q0 = 3LL * D * S;
  const int rows = G * S;
  const Strided dOt{dO, static_cast<long long>(S) * D, hd, 1, D};
  if (drop.on()) {  // dV = Pd^T dO first (Pd regenerated into the scratch), then the scratch is reused
    seq32::dropout_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(P, dS, rows, S, drop);
    if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
    count_launch();
    const Strided Pdt{dS, H * SS, SS, 1, S};
    if (int rc = seq32::bgemm(Pdt, dOt, dq + 2 * D, q0, hd, 3LL * D, G, S, hd, S, H, 1.0f, stream)) return rc;
  }
  const Strided dOv{dO, static_cast<long long>(S) * D, hd, D, 1}, Vn{q + 2 * D, q0, hd, 3LL * D, 1};
  if (int rc = seq32::bgemm(dOv, Vn, dS, H * SS, SS, S, G, S, S, hd, H, 1.0f, stream)) return rc;
  seq32::softmax_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(dS, P, rows, S,
                                                               1.0f / sqrtf(static_cast<float>(hd)), drop);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  if (!drop.on()) {
    const Strided Pt{P, H * SS, SS, 1, S};
    if (int rc = seq32::bgemm(Pt, dOt, dq + 2 * D, q0, hd, 3LL * D, G, S, hd, S, H, 1.0f, stream)) return rc;
  }
  const Strided dSt{dS, H * SS, SS, 1, S}, Qt{q, q0, hd, 1, 3LL * D};
  if (int rc = seq32::bgemm(dSt, Qt, dq + D, q0, hd, 3LL * D, G, S, hd, S, H, 1.0f, stream)) return rc;
  const Strided dSn{dS, H * SS, SS, S, 1}, Kt{q + D, q0, hd, 1, 3LL * D};
  return seq32::bgemm(dSn, Kt, dq, q0, hd, 3LL * D, G, S, hd, S, H, 1.0f, stream);
}

int attention_fwd(const void* qkv, void* out, float* lse, void* probs, float* scores, int dtype,
                  int B, int L, int D, int H, cudaStream_t stream, int pos_major, int keep_probs) {
  using namespace attn;
  if (dtype == DT_BF16 && qkv != nullptr && out != nullptr && (!keep_probs || probs != nullptr)) {
    const int rc = fused_batch_attention_fwd(qkv, out, keep_probs ? probs : nullptr, B, L, D, H, pos_major, stream);
    if (rc <= 0) return rc;
  }
  if (tc::eligible(dtype, B, D, H, probs, scores))
    return tc::fwd(qkv, out, probs, scores, B, L, D, H, pos_major, stream);
  if (lse == nullptr) return MMU_ERR_ARG;
  if (int rc = check(B, D, H)) return rc;
  const size_t smem = smem_bytes(B);
  dim3 grid(L * H, (B + TQ - 1) / TQ);
  if (dtype == DT_BF16) {
    if (set_smem(attn_fwd_kernel<__nv_bfloat16>, smem)) return MMU_ERR_CUDA;
    attn_fwd_kernel<__nv_bfloat16><<<grid, THREADS, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, B, L, D, H, pos_major);
  } else {
    if (set_smem(attn_fwd_kernel<float>, smem)) return MMU_ERR_CUDA;
    attn_fwd_kernel<float><<<grid, THREADS, smem, stream>>>(static_cast<const float*>(qkv),
                                                            static_cast<float*>(out), lse, B, L, D, H, pos_major);
  }
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}

int attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                  float* delta_ws, const void* probs, float* scores, void* dprobs, void* dqkv,
                  int dtype, int B, int L, int D, int H, cudaStream_t stream, int pos_major) {
  using namespace attn;
  if (tc::eligible(dtype, B, D, H, probs, scores) && dprobs != nullptr)
    return tc::bwd(qkv, dout, probs, scores, dprobs, dqkv, B, L, D, H, pos_major, stream);
  if (lse == nullptr || delta_ws == nullptr) return MMU_ERR_ARG;
  if (int rc = check(B, D, H)) return rc;
  const size_t smem = smem_bytes(B);
  dim3 grid(L * H, (B + TQ - 1) / TQ);
  if (dtype == DT_BF16) {
    using T = __nv_bfloat16;
    if (set_smem(attn_bwd_dq_kernel<T>, smem) || set_smem(attn_bwd_dkv_kernel<T>, smem))
      return MMU_ERR_CUDA;
    attn_bwd_dq_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse,
        delta_ws, static_cast<T*>(dqkv), B, L, D, H, pos_major);
    attn_bwd_dkv_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(dout), lse, delta_ws,
        static_cast<T*>(dqkv), B, L, D, H, pos_major);
  } else {
    using T = float;
    if (set_smem(attn_bwd_dq_kernel<T>, smem) || set_smem(attn_bwd_dkv_kernel<T>, smem))
      return MMU_ERR_CUDA;
    attn_bwd_dq_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse,
        delta_ws, static_cast<T*>(dqkv), B, L, D, H, pos_major);
    attn_bwd_dkv_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(dout), lse, delta_ws,
        static_cast<T*>(dqkv), B, L, D, H, pos_major);
  }
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch(2);
  return 0;
}

}  // namespace mmu
