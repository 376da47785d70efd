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
attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int B, int L,
                int D, int H) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, B);
  const int lh = blockIdx.x, l = lh / H, h = lh % H;
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
                   int B, int L, int D, int H) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, B);
  const int lh = blockIdx.x, l = lh / H, h = lh % H;
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
                    T* __restrict__ dqkv, int B, int L, int D, int H) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, B);
  const int lh = blockIdx.x, l = lh / H, h = lh % H;
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

}  // namespace attn

int attention_fwd(const void* qkv, void* out, float* lse, int dtype, int B, int L, int D, int H,
                  cudaStream_t stream) {
  using namespace attn;
  if (int rc = check(B, D, H)) return rc;
  const size_t smem = smem_bytes(B);
  dim3 grid(L * H, (B + TQ - 1) / TQ);
  if (dtype == DT_BF16) {
    if (set_smem(attn_fwd_kernel<__nv_bfloat16>, smem)) return MMU_ERR_CUDA;
    attn_fwd_kernel<__nv_bfloat16><<<grid, THREADS, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, B, L, D, H);
  } else {
    if (set_smem(attn_fwd_kernel<float>, smem)) return MMU_ERR_CUDA;
    attn_fwd_kernel<float><<<grid, THREADS, smem, stream>>>(static_cast<const float*>(qkv),
                                                            static_cast<float*>(out), lse, B, L, D, H);
  }
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}

int attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                  float* delta_ws, void* dqkv, int dtype, int B, int L, int D, int H,
                  cudaStream_t stream) {
  using namespace attn;
  if (int rc = check(B, D, H)) return rc;
  const size_t smem = smem_bytes(B);
  dim3 grid(L * H, (B + TQ - 1) / TQ);
  if (dtype == DT_BF16) {
    using T = __nv_bfloat16;
    if (set_smem(attn_bwd_dq_kernel<T>, smem) || set_smem(attn_bwd_dkv_kernel<T>, smem))
      return MMU_ERR_CUDA;
    attn_bwd_dq_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse,
        delta_ws, static_cast<T*>(dqkv), B, L, D, H);
    attn_bwd_dkv_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(dout), lse, delta_ws,
        static_cast<T*>(dqkv), B, L, D, H);
  } else {
    using T = float;
    if (set_smem(attn_bwd_dq_kernel<T>, smem) || set_smem(attn_bwd_dkv_kernel<T>, smem))
      return MMU_ERR_CUDA;
    attn_bwd_dq_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse,
        delta_ws, static_cast<T*>(dqkv), B, L, D, H);
    attn_bwd_dkv_kernel<T><<<grid, THREADS, smem, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(dout), lse, delta_ws,
        static_cast<T*>(dqkv), B, L, D, H);
  }
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch(2);
  return 0;
}

}  // namespace mmu
