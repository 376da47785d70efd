// Four-view FashionMNIST ResNet engine: forward / backward of the reference's MIMOResNet
// (src/model.py:17-100: conv3x3(4->64) + BN + ReLU, two BasicBlocks(64) at 14x14, two
// BasicBlocks(128) at 7x7 with a strided 1x1 downsample, AvgPool2d(4), MultiHeadFC) and
// BasicBlock (src/layers.py:7-38), BatchNorm in batch-statistics mode when training.
//
// Layout: activations are NHWC fp32, i.e. row (b, y, x) of a [B*H*W, C] matrix, so that
//   * a convolution is im2col (columns ordered (ci, ky, kx) = the reference's OIHW weight rows,
//     so checkpoints need no re-layout) followed by ONE GEMM  out[M, Co] = cols[M, Ci*k*k] W^T,
//     its input gradient one GEMM + a gather (col2im), its weight gradient one split-K GEMM;
//   * BatchNorm statistics are column statistics of that matrix.
// The GEMMs are the library's own: the fp32 FFMA path reproduces the reference's arithmetic
// (this model is its fp32 configuration); given a bf16 copy of the
// parameters, every layer whose K = ci*k*k keeps 16-byte operand rows (all but the 4-channel
// stem) runs on the tcgen05 kernel with bf16 columns / gradients and fp32 accumulation.
// No allocation, no synchronisation: caller-owned flat parameter / gradient / statistics buffers
// and workspace, everything enqueued on the caller's stream.
#include "resnet.h"
#include "kernels.h"

#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "gemm_api.h"

namespace mmu {

namespace {

#define RN_CHECK_LAUNCH()                                                        \
  do {                                                                          \
    const cudaError_t err__ = cudaGetLastError();                               \
    if (err__ != cudaSuccess) {                                                 \
      fprintf(stderr, "mmu: launch failed at %s:%d: %s\n", __FILE__, __LINE__, \
              cudaGetErrorString(err__));                                       \
      return MMU_ERR_CUDA;                                                      \
    }                                                                           \
    count_launch();                                                             \
  } while (0)

#define RN_TRY(x)               \
  do {                          \
    const int rc__ = (x);       \
    if (rc__ != 0) return rc__; \
  } while (0)

int blocks_for(size_t n, int per_block) {
  size_t b = (n + per_block - 1) / per_block;
  const size_t cap = static_cast<size_t>(sm_count()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

__device__ __forceinline__ void put(float* p, float v) { *p = v; }
__device__ __forceinline__ void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ float get(const float* p) { return *p; }
__device__ __forceinline__ float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 pk = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&lo);
  pk.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = pk;
}

// ------------------------------------------------------------------------- im2col / col2im
// cols[(b, yo, xo)][ci*k*k + ky*k + kx] = in(b, yo*stride + ky - pad, xo*stride + kx - pad, ci)
// Threads run over (row, tap, ci) with ci fastest: NHWC reads are coalesced.
template <typename TC>
__global__ void im2col_kernel(const float* __restrict__ in, int nchw, int B, int H, int W, int Ci,
                              int k, int stride, int pad, int Ho, int Wo, TC* __restrict__ cols) {
  const int kk = k * k, K = Ci * kk;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * K;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Ci);
    const int tap = static_cast<int>((i / Ci) % kk);
    const size_t row = i / K;
    const int xo = static_cast<int>(row % Wo), yo = static_cast<int>((row / Wo) % Ho);
    const int b = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    const int y = yo * stride + tap / k - pad, x = xo * stride + tap % k - pad;
    float v = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W)
      v = nchw ? in[((static_cast<size_t>(b) * Ci + ci) * H + y) * W + x]
               : in[((static_cast<size_t>(b) * H + y) * W + x) * Ci + ci];
    put(cols + row * K + static_cast<size_t>(ci) * kk + tap, v);
  }
}

// The same columns for a 3x3 convolution over an NHWC fp32 input, staged through shared memory: a
// block takes TR output rows, loads their 9 x Ci neighbourhood vectors with coalesced reads (ci
// fastest) into sm[r][tap][ci] (tap stride Ci + 1: conflict-free transposed reads) and writes the
// rows out in the reference's (ci, tap) column order with coalesced stores.  The scalar kernel
// above scatters 4-byte stores 36 bytes apart: 145 us per FashionMNIST layer against 35 here
// (116 MB of columns).  Needs Ci <= 256 and 256 % Ci == 0 (a thread keeps one channel).
constexpr int IM2COL_TR = 8;
template <typename TC>
__global__ void __launch_bounds__(256)
im2col3x3_tiled_kernel(const float* __restrict__ in, int B, int H, int W, int Ci, int stride, int pad,
                       int Ho, int Wo, TC* __restrict__ cols) {
  extern __shared__ float sm[];  // [TR][9][Ci + 1]
  __shared__ int rb[IM2COL_TR], ry[IM2COL_TR], rx[IM2COL_TR];
  const int CP = Ci + 1, K = Ci * 9;
  const size_t rows = static_cast<size_t>(B) * Ho * Wo;
  const int ci = threadIdx.x % Ci, t0 = threadIdx.x / Ci, tstep = 256 / Ci;
  for (size_t r0 = static_cast<size_t>(blockIdx.x) * IM2COL_TR; r0 < rows;
       r0 += static_cast<size_t>(gridDim.x) * IM2COL_TR) {
    if (threadIdx.x < IM2COL_TR) {
      const size_t row = r0 + threadIdx.x;
      const int xo = static_cast<int>(row % Wo), yo = static_cast<int>((row / Wo) % Ho);
      rb[threadIdx.x] = row < rows ? static_cast<int>(row / (static_cast<size_t>(Wo) * Ho)) : -1;
      ry[threadIdx.x] = yo * stride - pad;
      rx[threadIdx.x] = xo * stride - pad;
    }
    __syncthreads();
    for (int t = t0; t < IM2COL_TR * 9; t += tstep) {  // t = r * 9 + tap
      const int r = t / 9, tap = t - 9 * r;
      const int ky = tap / 3, kx = tap - 3 * ky;
      const int b = rb[r], y = ry[r] + ky, x = rx[r] + kx;
      float v = 0.f;
      if (b >= 0 && y >= 0 && y < H && x >= 0 && x < W)
        v = in[((static_cast<size_t>(b) * H + y) * W + x) * Ci + ci];
      sm[t * CP + ci] = v;
    }
    __syncthreads();
    for (int r = 0; r < IM2COL_TR; ++r) {
      if (rb[r] < 0) break;
      TC* dst = cols + (r0 + r) * K;
      for (int c = threadIdx.x; c < K; c += 256) {
        const int cc = c / 9, tap = c - 9 * cc;
        put(dst + c, sm[(r * 9 + tap) * CP + cc]);
      }
    }
    __syncthreads();
  }
}

// bf16 columns, NHWC input, K % 8 == 0: one thread writes 8 consecutive columns of a row as one
// 16-byte store (the scalar kernel above scatters 2-byte stores 2*k*k bytes apart); the reads are
// 3x3 neighbourhoods that overlap between rows and stay in L1 / L2.
template <typename TA>
__global__ void __launch_bounds__(256)
im2col_bf16x8_kernel(const TA* __restrict__ in, int B, int H, int W, int Ci, int k, int stride, int pad,
                     int Ho, int Wo, __nv_bfloat16* __restrict__ cols) {
  const int kk = k * k, K = Ci * kk, K8 = K >> 3;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * K8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % K8);
    const size_t row = i / K8;
    const int xo = static_cast<int>(row % Wo), yo = static_cast<int>((row / Wo) % Ho);
    const int b = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    const TA* base = in + static_cast<size_t>(b) * H * W * Ci;
    int ci = (c8 * 8) / kk, tap = (c8 * 8) % kk;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ky = tap / k, kx = tap - ky * k;
      const int y = yo * stride + ky - pad, x = xo * stride + kx - pad;
      v[j] = (y >= 0 && y < H && x >= 0 && x < W) ? get(base + (static_cast<size_t>(y) * W + x) * Ci + ci) : 0.f;
      if (++tap == kk) { tap = 0; ++ci; }
    }
    uint4 pk;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
    pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
    *reinterpret_cast<uint4*>(cols + row * K + static_cast<size_t>(c8) * 8) = pk;
  }
}

// Padded-K variant for an NCHW input (the image encoder's 7x7 stem): columns [K, Kp) are zero.
__global__ void __launch_bounds__(256)
im2col_nchw_pad_bf16x8_kernel(const float* __restrict__ in, int B, int H, int W, int Ci, int k, int stride,
                              int pad, int Ho, int Wo, int Kp, __nv_bfloat16* __restrict__ cols) {
  const int kk = k * k, K = Ci * kk, K8 = Kp >> 3;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * K8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % K8);
    const size_t row = i / K8;
    const int xo = static_cast<int>(row % Wo), yo = static_cast<int>((row / Wo) % Ho);
    const int b = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    const float* base = in + static_cast<size_t>(b) * Ci * H * W;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      float val = 0.f;
      if (c < K) {
        const int ci = c / kk, tap = c - ci * kk;
        const int ky = tap / k, kx = tap - ky * k;
        const int y = yo * stride + ky - pad, x = xo * stride + kx - pad;
        if (y >= 0 && y < H && x >= 0 && x < W) val = base[(static_cast<size_t>(ci) * H + y) * W + x];
      }
      v[j] = val;
    }
    uint4 pk;
    __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
    pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
    *reinterpret_cast<uint4*>(cols + row * Kp + static_cast<size_t>(c8) * 8) = pk;
  }
}
__global__ void pad_weights_bf16_kernel(const float* __restrict__ w, int co, int K, int Kp,
                                        __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= co * Kp) return;
  const int r = i / Kp, c = i % Kp;
  out[i] = __float2bfloat16_rn(c < K ? w[r * K + c] : 0.f);
}
__global__ void unpad_add_kernel(const float* __restrict__ dwp, int co, int K, int Kp, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= co * K) return;
  dw[i] += dwp[(i / K) * Kp + i % K];
}

// ---- tap-major columns: cols[row][tap * Ci + ci].  A thread moves 8 consecutive channels of one
// tap: one 16-byte load from the bf16 NHWC input (two from an fp32 input, rounded here), one
// 16-byte store (all fully coalesced).
// I: index type of the flat work index -- unsigned when it fits (every shape of the two networks):
// the decode is six divisions per 16 bytes moved, and a 64-bit division is ~5x a 32-bit one (ncu:
// these kernels are issue bound on exactly that, not on memory).
template <typename TA, typename I>
__global__ void __launch_bounds__(256)
im2col_tm_kernel(const TA* __restrict__ in, int B, int H, int W, int Ci, int k, int stride, int pad,
                 int Ho, int Wo, __nv_bfloat16* __restrict__ cols) {
  const int kk = k * k, C8 = Ci >> 3;
  const size_t K = static_cast<size_t>(Ci) * kk;
  const I total = static_cast<I>(B) * Ho * Wo * kk * C8;
  for (I i = blockIdx.x * static_cast<I>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<I>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % static_cast<I>(C8));
    const I t = i / static_cast<I>(C8);
    const int tap = static_cast<int>(t % static_cast<I>(kk));
    const I row = t / static_cast<I>(kk);
    const I rowy = row / static_cast<I>(Wo);
    const int xo = static_cast<int>(row - rowy * Wo), yo = static_cast<int>(rowy % static_cast<I>(Ho));
    const int b = static_cast<int>(rowy / static_cast<I>(Ho));
    const int ky = tap / k, kx = tap - ky * k;
    const int y = yo * stride + ky - pad, x = xo * stride + kx - pad;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const TA* src = in + ((static_cast<size_t>(b) * H + y) * W + x) * Ci + c8 * 8;
      if constexpr (sizeof(TA) == 2) {
        v = *reinterpret_cast<const uint4*>(src);
      } else {
        const float4 lo = *reinterpret_cast<const float4*>(src), hi = *reinterpret_cast<const float4*>(src + 4);
        __nv_bfloat162 t0 = __floats2bfloat162_rn(lo.x, lo.y), t1 = __floats2bfloat162_rn(lo.z, lo.w);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(hi.x, hi.y), t3 = __floats2bfloat162_rn(hi.z, hi.w);
        v.x = *reinterpret_cast<uint32_t*>(&t0); v.y = *reinterpret_cast<uint32_t*>(&t1);
        v.z = *reinterpret_cast<uint32_t*>(&t2); v.w = *reinterpret_cast<uint32_t*>(&t3);
      }
    }
    *reinterpret_cast<uint4*>(cols + static_cast<size_t>(row) * K + static_cast<size_t>(tap) * Ci + c8 * 8) = v;
  }
}
// din[pixel][ci .. ci+8) (+)= sum over the taps that reach the pixel of dcols[row][tap * Ci + ci ..]
template <typename I>
__global__ void __launch_bounds__(256)
col2im_tm_kernel(const __nv_bfloat16* __restrict__ dcols, int B, int H, int W, int Ci, int k, int stride,
                 int pad, int Ho, int Wo, float* __restrict__ dx, int accumulate) {
  const int kk = k * k, C8 = Ci >> 3;
  const size_t K = static_cast<size_t>(Ci) * kk;
  const I total = static_cast<I>(B) * H * W * C8;
  for (I i = blockIdx.x * static_cast<I>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<I>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % static_cast<I>(C8));
    const I pix = i / static_cast<I>(C8);
    const I pixy = pix / static_cast<I>(W);
    const int x = static_cast<int>(pix - pixy * W), y = static_cast<int>(pixy % static_cast<I>(H));
    const int b = static_cast<int>(pixy / static_cast<I>(H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ky = 0; ky < k; ++ky) {
      const int ty = y + pad - ky;
      if (ty < 0 || ty % stride != 0) continue;
      const int yo = ty / stride;
      if (yo >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = x + pad - kx;
        if (tx < 0 || tx % stride != 0) continue;
        const int xo = tx / stride;
        if (xo >= Wo) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(
            dcols + ((static_cast<size_t>(b) * Ho + yo) * Wo + xo) * K + static_cast<size_t>(ky * k + kx) * Ci + c8 * 8);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(w[j] << 16);
          acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
        }
      }
    }
    float* o = dx + static_cast<size_t>(pix) * Ci + c8 * 8;
    float4 lo = make_float4(acc[0], acc[1], acc[2], acc[3]), hi = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (accumulate) {
      const float4 a = *reinterpret_cast<const float4*>(o), c = *reinterpret_cast<const float4*>(o + 4);
      lo.x += a.x; lo.y += a.y; lo.z += a.z; lo.w += a.w;
      hi.x += c.x; hi.y += c.y; hi.z += c.z; hi.w += c.w;
    }
    *reinterpret_cast<float4*>(o) = lo;
    *reinterpret_cast<float4*>(o + 4) = hi;
  }
}
// wperm[co][tap * Ci + ci] = bf16(w[(co * Ci + ci) * kk + tap])   (OIHW -> tap-major rows)
__global__ void permute_weights_kernel(const float* __restrict__ w, int co, int Ci, int kk,
                                       __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = Ci * kk;
  if (i >= co * K) return;
  const int o = i / K, r = i % K, tap = r / Ci, ci = r % Ci;
  out[i] = __float2bfloat16_rn(w[(static_cast<size_t>(o) * Ci + ci) * kk + tap]);
}
// dW[(co * Ci + ci) * kk + tap] += dwp[co][tap * Ci + ci]
__global__ void unpermute_add_kernel(const float* __restrict__ dwp, int co, int Ci, int kk, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = Ci * kk;
  if (i >= co * K) return;
  const int o = i / K, r = i % K, ci = r / kk, tap = r % kk;
  dw[i] += dwp[static_cast<size_t>(o) * K + static_cast<size_t>(tap) * Ci + ci];
}

// dx(b, y, x, ci) (+)= sum over taps of dcols[(b, yo, xo)][ci*k*k + tap] with y = yo*stride+ky-pad
template <typename TC>
__global__ void col2im_kernel(const TC* __restrict__ dcols, int B, int H, int W, int Ci, int k,
                              int stride, int pad, int Ho, int Wo, float* __restrict__ dx,
                              int accumulate) {
  const int kk = k * k, K = Ci * kk;
  const size_t total = static_cast<size_t>(B) * H * W * Ci;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % Ci);
    const size_t pix = i / Ci;
    const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    float s = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int ty = y + pad - ky;
      if (ty < 0 || ty % stride != 0) continue;
      const int yo = ty / stride;
      if (yo >= Ho) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int tx = x + pad - kx;
        if (tx < 0 || tx % stride != 0) continue;
        const int xo = tx / stride;
        if (xo >= Wo) continue;
        s += get(dcols + ((static_cast<size_t>(b) * Ho + yo) * Wo + xo) * K +
                 static_cast<size_t>(ci) * kk + ky * k + kx);
      }
    }
    dx[i] = accumulate ? dx[i] + s : s;
  }
}

// ------------------------------------------------------------------------------- BatchNorm
// Column sums of t and t^2 over the M rows (fp64 partials: E[x^2]-E[x]^2 is then safe).
// Block = 32 columns x 8 row lanes; grid.y slabs of rows.
// Geometry of the column-statistics kernels: a thread owns 4 adjacent columns (one 16- / 8-byte
// load per row), TX threads across the columns (a power of two, <= 32, i.e. up to 128 columns per
// block) and 256 / TX row lanes; per-thread fp32 partial sums over short row runs are folded into
// fp64 accumulators, reduced through shared memory, then added atomically (fp64).
struct ColGeo { int tx, ty, gx; };
inline int stat_rows_per_block(int M, const ColGeo& g);
inline ColGeo col_geo(int C) {
  int tx = 1;
  while (tx < 32 && tx * 2 <= C / 4) tx *= 2;
  return ColGeo{tx, 256 / tx, (C / 4 + tx - 1) / tx};
}

// rows per block: enough blocks to fill the machine a few times over, at least 4 row steps each
inline int stat_rows_per_block(int M, const ColGeo& g) {
  const int target_blocks = 148 * 6;
  int gy = (target_blocks + g.gx - 1) / g.gx;
  int rpb = (M + gy - 1) / gy;
  const int min_rows = 8 * g.ty;
  if (rpb < min_rows) rpb = min_rows;
  return (rpb + g.ty - 1) / g.ty * g.ty;
}

template <typename TA>
__global__ void __launch_bounds__(256)
bn_sums_kernel(const TA* __restrict__ t, int M, int C, int rows_per_block, int TX, double* __restrict__ sums) {
  __shared__ double red[256][8];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const int c = (blockIdx.x * TX + tx) * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < C) {
    int r = r0 + ty;
    for (; r + 3 * TY < r1; r += 4 * TY) {  // four independent vector loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld4(t + static_cast<size_t>(r + u * TY) * C + c);
      float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s1.x += v[u].x; s1.y += v[u].y; s1.z += v[u].z; s1.w += v[u].w;
        s2.x = fmaf(v[u].x, v[u].x, s2.x); s2.y = fmaf(v[u].y, v[u].y, s2.y);
        s2.z = fmaf(v[u].z, v[u].z, s2.z); s2.w = fmaf(v[u].w, v[u].w, s2.w);
      }
      a[0] += s1.x; a[1] += s1.y; a[2] += s1.z; a[3] += s1.w;
      b[0] += s2.x; b[1] += s2.y; b[2] += s2.z; b[3] += s2.w;
    }
    for (; r < r1; r += TY) {
      const float4 v = ld4(t + static_cast<size_t>(r) * C + c);
      a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
      b[0] += static_cast<double>(v.x) * v.x; b[1] += static_cast<double>(v.y) * v.y;
      b[2] += static_cast<double>(v.z) * v.z; b[3] += static_cast<double>(v.w) * v.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = a[j]; red[threadIdx.x][4 + j] = b[j]; }
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int y = 1; y < TY; ++y)
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += red[y * TX + tx][j]; b[j] += red[y * TX + tx][4 + j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sums[c + j], a[j]);
      atomicAdd(&sums[C + c + j], b[j]);
    }
  }
}

// mean / rstd of the batch (biased variance, eps 1e-5); running statistics move by `momentum`
// towards the batch mean and the UNBIASED batch variance (torch.nn.BatchNorm2d).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int M, int C, float momentum,
                                   float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = sums[c] / M;
  double var = sums[C + c] / M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = static_cast<float>(mu);
  rstd[c] = static_cast<float>(1.0 / sqrt(var + 1e-5));
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.f - momentum) * running_var[c] +
                     momentum * static_cast<float>(var * M / (M > 1 ? M - 1 : 1));
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, int C,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = running_mean[c];
  rstd[c] = 1.0f / sqrtf(running_var[c] + 1e-5f);
}

// out = [relu]( (t - mean) * rstd * gamma + beta [+ residual] )
template <typename TA>
__global__ void bn_apply_kernel(const TA* __restrict__ t, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const TA* __restrict__ residual,
                                int relu, TA* __restrict__ out, size_t n4, int C) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // channel of the vector: a 32-bit modulo when the element index fits (a 64-bit one costs more
    // instructions than the rest of the iteration)
    const int c = n4 < (1ull << 30) ? static_cast<int>((static_cast<unsigned>(i) * 4u) % static_cast<unsigned>(C))
                                    : static_cast<int>((i * 4) % C);
    const float4 v = ld4(t + 4 * i);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    float4 o;
    o.x = (v.x - mu.x) * rs.x * g.x + be.x;
    o.y = (v.y - mu.y) * rs.y * g.y + be.y;
    o.z = (v.z - mu.z) * rs.z * g.z + be.z;
    o.w = (v.w - mu.w) * rs.w * g.w + be.w;
    if (residual != nullptr) {
      const float4 r = ld4(residual + 4 * i);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    st4(out + 4 * i, o);
  }
}

// Backward, pass 1: dyb = dout * [out > 0] (ReLU mask from the saved output; out == nullptr: no
// ReLU), column sums of dyb and dyb * xhat.
template <typename TA>
__global__ void __launch_bounds__(256)
bn_bwd_sums_kernel(const float* __restrict__ dout, const float* dout2, const TA* __restrict__ out,
                   const TA* __restrict__ t, const float* __restrict__ mean,
                   const float* __restrict__ rstd, int M, int C, int rows_per_block, int TX,
                   float* dyb, double* __restrict__ sums) {
  // dout2 (nullable): a second gradient term added on the fly -- the identity-shortcut gradient of
  // the NEXT residual block, which would otherwise cost a separate add pass.  It may alias dyb
  // (element i is read before element i is written, by the same thread).
  __shared__ double red[256][8];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX, TY = 256 / TX;
  const int c = (blockIdx.x * TX + tx) * 4;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  double a[4] = {0.0, 0.0, 0.0, 0.0}, b[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < C) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    int r = r0 + ty;
    for (; r + TY < r1; r += 2 * TY) {  // two rows per iteration: six independent vector loads
      float4 d[2], tv[2], ov[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const size_t i = static_cast<size_t>(r + u * TY) * C + c;
        d[u] = *reinterpret_cast<const float4*>(dout + i);
        if (dout2 != nullptr) {
          const float4 e = *reinterpret_cast<const float4*>(dout2 + i);
          d[u].x += e.x; d[u].y += e.y; d[u].z += e.z; d[u].w += e.w;
        }
        tv[u] = ld4(t + i);
        ov[u] = out != nullptr ? ld4(out + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      }
      float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!(ov[u].x > 0.f)) d[u].x = 0.f;
        if (!(ov[u].y > 0.f)) d[u].y = 0.f;
        if (!(ov[u].z > 0.f)) d[u].z = 0.f;
        if (!(ov[u].w > 0.f)) d[u].w = 0.f;
        *reinterpret_cast<float4*>(dyb + static_cast<size_t>(r + u * TY) * C + c) = d[u];
        s1.x += d[u].x; s1.y += d[u].y; s1.z += d[u].z; s1.w += d[u].w;
        s2.x = fmaf(d[u].x, (tv[u].x - mu.x) * rs.x, s2.x); s2.y = fmaf(d[u].y, (tv[u].y - mu.y) * rs.y, s2.y);
        s2.z = fmaf(d[u].z, (tv[u].z - mu.z) * rs.z, s2.z); s2.w = fmaf(d[u].w, (tv[u].w - mu.w) * rs.w, s2.w);
      }
      a[0] += s1.x; a[1] += s1.y; a[2] += s1.z; a[3] += s1.w;
      b[0] += s2.x; b[1] += s2.y; b[2] += s2.z; b[3] += s2.w;
    }
    for (; r < r1; r += TY) {
      const size_t i = static_cast<size_t>(r) * C + c;
      float4 d = *reinterpret_cast<const float4*>(dout + i);
      if (dout2 != nullptr) {
        const float4 e = *reinterpret_cast<const float4*>(dout2 + i);
        d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
      }
      const float4 tv = ld4(t + i);
      if (out != nullptr) {
        const float4 ov = ld4(out + i);
        if (!(ov.x > 0.f)) d.x = 0.f;
        if (!(ov.y > 0.f)) d.y = 0.f;
        if (!(ov.z > 0.f)) d.z = 0.f;
        if (!(ov.w > 0.f)) d.w = 0.f;
      }
      *reinterpret_cast<float4*>(dyb + i) = d;
      a[0] += d.x; a[1] += d.y; a[2] += d.z; a[3] += d.w;
      b[0] += static_cast<double>(d.x) * ((tv.x - mu.x) * rs.x); b[1] += static_cast<double>(d.y) * ((tv.y - mu.y) * rs.y);
      b[2] += static_cast<double>(d.z) * ((tv.z - mu.z) * rs.z); b[3] += static_cast<double>(d.w) * ((tv.w - mu.w) * rs.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = a[j]; red[threadIdx.x][4 + j] = b[j]; }
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int y = 1; y < TY; ++y)
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += red[y * TX + tx][j]; b[j] += red[y * TX + tx][4 + j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&sums[c + j], a[j]);
      atomicAdd(&sums[C + c + j], b[j]);
    }
  }
}

// Backward, pass 2: dt = gamma * rstd * (dyb - mean(dyb) - xhat * mean(dyb * xhat));
// thread 0..C-1 of block 0 also accumulate dgamma / dbeta.
template <typename TD, typename TA>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dyb, const TA* __restrict__ t,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const double* __restrict__ sums,
                                    int M, int C, TD* __restrict__ dt, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
  const size_t n4 = static_cast<size_t>(M) * C / 4;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] += static_cast<float>(sums[c]);
      dgamma[c] += static_cast<float>(sums[C + c]);
    }
  const double invM = 1.0 / M;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // channel of the vector: a 32-bit modulo when the element index fits (a 64-bit one costs more
    // instructions than the rest of the iteration)
    const int c = n4 < (1ull << 30) ? static_cast<int>((static_cast<unsigned>(i) * 4u) % static_cast<unsigned>(C))
                                    : static_cast<int>((i * 4) % C);
    const float4 d = *reinterpret_cast<const float4*>(dyb + 4 * i);
    const float4 tv = ld4(t + 4 * i);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const double2 sa = *reinterpret_cast<const double2*>(sums + c), sb = *reinterpret_cast<const double2*>(sums + c + 2);
    const double2 qa = *reinterpret_cast<const double2*>(sums + C + c), qb = *reinterpret_cast<const double2*>(sums + C + c + 2);
    const float m1[4] = {static_cast<float>(sa.x * invM), static_cast<float>(sa.y * invM),
                         static_cast<float>(sb.x * invM), static_cast<float>(sb.y * invM)};
    const float m2[4] = {static_cast<float>(qa.x * invM), static_cast<float>(qa.y * invM),
                         static_cast<float>(qb.x * invM), static_cast<float>(qb.y * invM)};
    float4 o;
    o.x = g.x * rs.x * (d.x - m1[0] - (tv.x - mu.x) * rs.x * m2[0]);
    o.y = g.y * rs.y * (d.y - m1[1] - (tv.y - mu.y) * rs.y * m2[1]);
    o.z = g.z * rs.z * (d.z - m1[2] - (tv.z - mu.z) * rs.z * m2[2]);
    o.w = g.w * rs.w * (d.w - m1[3] - (tv.w - mu.w) * rs.w * m2[3]);
    st4(dt + 4 * i, o);
  }
}

// ------------------------------------------------------------ pooling / elementwise helpers
// AvgPool2d(4) on the 7x7 map keeps ONE window (rows/cols 0..3): p[b, c] = mean of 16 pixels.
__global__ void pool_fwd_kernel(const float* __restrict__ a, int B, int H, int W, int C, int win,
                                float* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, b = i / C;
  float s = 0.f;
  for (int y = 0; y < win; ++y)
    for (int x = 0; x < win; ++x) s += a[((static_cast<size_t>(b) * H + y) * W + x) * C + c];
  p[i] = s / (win * win);
}
__global__ void pool_bwd_kernel(const float* __restrict__ dp, int B, int H, int W, int C, int win,
                                float* __restrict__ da) {
  const size_t total = static_cast<size_t>(B) * H * W * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    da[i] = (y < win && x < win) ? dp[static_cast<size_t>(b) * C + c] / (win * win) : 0.f;
  }
}
__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    a[i] += b[i];
}
__global__ void colsum_rows_kernel(const float* __restrict__ x, int M, int N, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int r = 0; r < M; ++r) s += x[static_cast<size_t>(r) * N + c];
  out[c] += s;
}

// ------------------------------------------------------------------------------- layout
struct ConvBn {
  int ci, co, k, stride, pad, hin, hout;
  long long w, g, b;  // offsets into the flat parameter buffer: conv weight, BN gain, BN bias
  long long stat;     // offset into the flat statistics buffer: running_mean[co] | running_var[co]
  long long wp = -1;  // offset (elements) of this layer's tap-major bf16 weight copy, or -1
};
struct Net {
  ConvBn stem;
  ConvBn c1[4], c2[4], ds;  // four BasicBlocks; ds belongs to block 2 (layer2.0)
  long long fc_w, fc_b;
  long long n_params, n_stats;
  long long n_wperm = 0;  // elements of the tap-major bf16 weight copies (3x3 layers with ci % 8 == 0)
  long long max_wk = 0;   // largest co * K among them (weight-gradient scratch)
};

struct Tab {
  ParamEntry* out;
  int max, n;
  long long cursor;
  long long add(const char* name, int rows, int cols) {
    const long long numel = static_cast<long long>(rows) * (cols > 0 ? cols : 1);
    const long long off = cursor;
    if (out != nullptr && n < max) {
      ParamEntry& e = out[n];
      std::memset(&e, 0, sizeof(e));
      std::snprintf(e.name, sizeof(e.name), "%s", name);
      e.offset = off; e.numel = numel; e.rows = rows; e.cols = cols; e.stage = 0;
    }
    ++n;
    cursor = (cursor + numel + 63) / 64 * 64;
    return off;
  }
};

// Parameter order = the reference's named_parameters() order (src/model.py:21-31, layers.py:13-19:
// a BasicBlock registers bn1, conv1, bn2, conv2, downsample in that order).
int build(const ResNetConfig& c, Net* net, ParamEntry* ptab, int pmax, ParamEntry* stab, int smax,
          int* n_ptab, int* n_stab) {
  if (c.B < 1 || c.cin < 1 || c.H != 14 || c.W != 14 || c.E < 1 || c.E > 16 || c.C < 1) return MMU_ERR_SHAPE;
  Tab p{ptab, pmax, 0, 0}, s{stab, smax, 0, 0};
  char nm[96];
  auto bn = [&](const char* prefix, int ch, ConvBn* cb) {
    std::snprintf(nm, sizeof(nm), "%s.weight", prefix); cb->g = p.add(nm, ch, 0);
    std::snprintf(nm, sizeof(nm), "%s.bias", prefix);   cb->b = p.add(nm, ch, 0);
    std::snprintf(nm, sizeof(nm), "%s.running_mean", prefix); cb->stat = s.add(nm, ch, 0);
    std::snprintf(nm, sizeof(nm), "%s.running_var", prefix);  s.add(nm, ch, 0);
  };
  auto conv = [&](const char* name, int ci, int co, int k, int stride, int hin, ConvBn* cb) {
    cb->ci = ci; cb->co = co; cb->k = k; cb->stride = stride; cb->pad = k / 2; cb->hin = hin;
    cb->hout = (hin + 2 * cb->pad - k) / stride + 1;
    cb->w = p.add(name, co, ci * k * k);
  };
  conv("conv1.weight", c.cin, 64, 3, 1, 14, &net->stem);
  bn("bn1", 64, &net->stem);
  const int planes[4] = {64, 64, 128, 128}, inpl[4] = {64, 64, 64, 128}, strides[4] = {1, 1, 2, 1};
  const int hin[4] = {14, 14, 14, 7};
  const char* names[4] = {"layer1.0", "layer1.1", "layer2.0", "layer2.1"};
  char pre[64];
  for (int i = 0; i < 4; ++i) {
    std::snprintf(pre, sizeof(pre), "%s.bn1", names[i]);
    ConvBn tmp1{}, tmp2{};
    bn(pre, planes[i], &tmp1);
    std::snprintf(pre, sizeof(pre), "%s.conv1.weight", names[i]);
    conv(pre, inpl[i], planes[i], 3, strides[i], hin[i], &net->c1[i]);
    net->c1[i].g = tmp1.g; net->c1[i].b = tmp1.b; net->c1[i].stat = tmp1.stat;
    std::snprintf(pre, sizeof(pre), "%s.bn2", names[i]);
    bn(pre, planes[i], &tmp2);
    std::snprintf(pre, sizeof(pre), "%s.conv2.weight", names[i]);
    conv(pre, planes[i], planes[i], 3, 1, net->c1[i].hout, &net->c2[i]);
    net->c2[i].g = tmp2.g; net->c2[i].b = tmp2.b; net->c2[i].stat = tmp2.stat;
    if (i == 2) {
      std::snprintf(pre, sizeof(pre), "%s.downsample.0.weight", names[i]);
      conv(pre, inpl[i], planes[i], 1, strides[i], hin[i], &net->ds);
      net->ds.pad = 0;
      net->ds.hout = (hin[i] - 1) / strides[i] + 1;
      std::snprintf(pre, sizeof(pre), "%s.downsample.1", names[i]);
      ConvBn tmpd{};
      bn(pre, planes[i], &tmpd);
      net->ds.g = tmpd.g; net->ds.b = tmpd.b; net->ds.stat = tmpd.stat;
    }
  }
  // tap-major copies for the tensor-core mode (every 3x3 layer but the 4-channel stem)
  net->n_wperm = 0;
  net->max_wk = 0;
  for (int i = 0; i < 4; ++i) {
    for (ConvBn* l : {&net->c1[i], &net->c2[i]}) {
      if (l->ci % 8 != 0) continue;
      const long long wk = static_cast<long long>(l->co) * l->ci * l->k * l->k;
      l->wp = net->n_wperm;
      net->n_wperm += (wk + 63) / 64 * 64;
      if (wk > net->max_wk) net->max_wk = wk;
    }
  }
  net->fc_w = p.add("output_layer.fc.weight", c.E * c.C, 128);
  net->fc_b = p.add("output_layer.fc.bias", c.E * c.C, 0);
  net->n_params = p.cursor;
  net->n_stats = s.cursor;
  if (n_ptab) *n_ptab = p.n;
  if (n_stab) *n_stab = s.n;
  return 0;
}

// ------------------------------------------------------------------------------ workspace
struct CbWs {
  float* t;     // conv output, pre-BN          [Mout, co]
  float* mean;  // batch (or running) mean      [co]
  float* rstd;
  float* out;   // post BN (+residual) (+ReLU)  [Mout, co]
};
struct Ws {
  CbWs stem, c1[4], c2[4], ds;
  float* cols;     // im2col scratch (largest layer)
  float* pooled;   // [B, 128]
  double* sums;    // [2 * 128] BN reduction scratch
  // backward
  float* dcols;
  float* dyb;      // masked upstream gradient of the BN being differentiated
  float* dt;       // gradient w.r.t. the conv output
  float* dact[2];  // gradients w.r.t. activations (ping-pong), largest activation size
  float* dres;     // gradient flowing along the residual connection
  float* dpooled;
  // tensor-core path of a convolution whose K = ci*k*k is not a multiple of 8 (the 7x7x3 stem of
  // the MMBT image encoder, K = 147): K padded with zero columns to Kp = 152 -- a padded bf16
  // copy of the weights and an fp32 [co x Kp] scratch for the weight gradient.  Null: disabled.
  void* pad_w;
  float* pad_dw;
  // tap-major 3x3 convolutions (image encoder, bf16 activations): columns ordered (tap, ci) so that
  // im2col / col2im move 16-byte vectors; the weights are re-ordered into `wperm` (bf16, all 3x3
  // layers) by the forward, the weight gradient goes through the fp32 scratch `dwperm`.
  void* wperm;
  float* dwperm;
  long long bytes;
};
struct Bump {
  char* base;
  long long off;
  template <typename T>
  T* take(long long bytes) {
    const long long o = off;
    off = (off + bytes + 255) / 256 * 256;
    return base != nullptr ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};
long long rows_of(const ResNetConfig& c, int h) { return static_cast<long long>(c.B) * h * h; }

void carve(const ResNetConfig& c, const Net& n, int training, void* base, Ws* w) {
  Bump b{static_cast<char*>(base), 0};
  auto cb = [&](const ConvBn& l, CbWs* o) {
    const long long m = rows_of(c, l.hout);
    o->t = b.take<float>(m * l.co * 4);
    o->mean = b.take<float>(l.co * 4);
    o->rstd = b.take<float>(l.co * 4);
    o->out = b.take<float>(m * l.co * 4);
  };
  cb(n.stem, &w->stem);
  long long max_cols = rows_of(c, n.stem.hout) * n.stem.ci * 9;
  for (int i = 0; i < 4; ++i) {
    cb(n.c1[i], &w->c1[i]);
    cb(n.c2[i], &w->c2[i]);
    const long long a = rows_of(c, n.c1[i].hout) * n.c1[i].ci * 9, d = rows_of(c, n.c2[i].hout) * n.c2[i].ci * 9;
    if (a > max_cols) max_cols = a;
    if (d > max_cols) max_cols = d;
  }
  cb(n.ds, &w->ds);
  w->cols = b.take<float>(max_cols * 4);
  w->pooled = b.take<float>(static_cast<long long>(c.B) * 128 * 4);
  w->sums = b.take<double>(2 * 128 * 8);
  if (training) {
    const long long max_act = rows_of(c, 14) * 64;  // == rows_of(7) * 128 * 2: the largest activation
    w->dcols = b.take<float>(max_cols * 4);
    w->dyb = b.take<float>(max_act * 4);
    w->dt = b.take<float>(max_act * 4);
    w->dact[0] = b.take<float>(max_act * 4);
    w->dact[1] = b.take<float>(max_act * 4);
    w->dres = b.take<float>(max_act * 4);
    w->dpooled = b.take<float>(static_cast<long long>(c.B) * 128 * 4);
  } else {
    w->dcols = w->dyb = w->dt = w->dact[0] = w->dact[1] = w->dres = w->dpooled = nullptr;
  }
  // the stem (K = 9 * cin, 36 for the four-view model) runs on the tensor cores with K padded to
  // a multiple of 8 when a bf16 shadow is given -- its fp32 weight gradient reduces 50 176 rows
  // into a 64 x 36 tile and was 14 % of the bf16 step on the FFMA kernel
  const long long stem_k = static_cast<long long>(n.stem.ci) * n.stem.k * n.stem.k;
  const long long stem_kp = (stem_k + 7) / 8 * 8;
  w->pad_w = stem_k % 8 != 0 ? b.take<void>(n.stem.co * stem_kp * 2) : nullptr;
  w->pad_dw = (training && stem_k % 8 != 0) ? b.take<float>(n.stem.co * stem_kp * 4) : nullptr;
  w->wperm = n.n_wperm > 0 ? b.take<void>(n.n_wperm * 2) : nullptr;
  w->dwperm = (training && n.max_wk > 0) ? b.take<float>(n.max_wk * 4) : nullptr;
  w->bytes = b.off;
}

GemmEpilogue store_epi(float* out, long long ld, const float* bias) {
  GemmEpilogue e{};
  e.mode = EPI_STORE; e.out_bf16 = 0; e.out = out; e.ld_out = ld; e.bias = bias; e.alpha = 1.0f;
  return e;
}

// ------------------------------------------------------------------------- layer helpers
struct Ctx {
  const ResNetConfig& c;
  const void* shadow;  // bf16 copy of params (tensor-core path) or null (fp32 path)
  const float* params;
  float* stats;      // running statistics (updated in training mode), may be null in eval? no: required
  float* grads;
  int training;
  Ws& w;
  cudaStream_t st;
  // 1: the activations (conv outputs t, BatchNorm outputs, pooled maps) are bf16 -- the MMBT image
  // encoder's tensor-core mode; the buffers are still declared float* and reinterpreted.  An NCHW
  // input (the images) is always fp32.  0: fp32 activations (FashionMNIST engine, parity paths).
  int act16 = 0;
};
using bf16_t = __nv_bfloat16;
inline const bf16_t* as16(const float* p) { return reinterpret_cast<const bf16_t*>(p); }
inline bf16_t* as16(float* p) { return reinterpret_cast<bf16_t*>(p); }

// A layer runs on the tcgen05 path when a bf16 shadow is given and its GEMM K (= ci*k*k) keeps
// operand rows 16-byte aligned (every layer but the 4-channel stem, K = 36).
// fp32 columns of a layer (parity path): the shared-memory-tiled kernel for NHWC 3x3 layers
int f32_columns(const Ctx& x, const ConvBn& l, const float* in, int in_nchw, float* cols) {
  const size_t M = static_cast<size_t>(rows_of(x.c, l.hout));
  const size_t ncols = M * l.ci * l.k * l.k;
  if (!in_nchw && l.k == 3 && l.ci <= 256 && 256 % l.ci == 0) {
    const size_t smem = static_cast<size_t>(IM2COL_TR) * 9 * (l.ci + 1) * sizeof(float);
    if (smem <= 48 * 1024) {
      im2col3x3_tiled_kernel<float><<<blocks_for(M, IM2COL_TR), 256, smem, x.st>>>(
          in, x.c.B, l.hin, l.hin, l.ci, l.stride, l.pad, l.hout, l.hout, cols);
      RN_CHECK_LAUNCH();
      return 0;
    }
  }
  im2col_kernel<float><<<blocks_for(ncols, 256), 256, 0, x.st>>>(
      in, in_nchw, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
  RN_CHECK_LAUNCH();
  return 0;
}

bool use_tc(const Ctx& x, const ConvBn& l) { return x.shadow != nullptr && (l.ci * l.k * l.k) % 8 == 0; }
// ... or, with K padded to a multiple of 8, when the padded scratch exists (image-encoder stem)
bool use_tc_padded(const Ctx& x, const ConvBn& l, int in_nchw) {
  return x.shadow != nullptr && (l.ci * l.k * l.k) % 8 != 0 && in_nchw && x.w.pad_w != nullptr;
}
// tap-major layer: 3x3 (or larger) convolution on the tensor-core path with a re-ordered weight
// copy (bf16 activations in the image encoder, fp32 activations in the FashionMNIST engine)
bool use_tm(const Ctx& x, const ConvBn& l) {
  return x.shadow != nullptr && l.k > 1 && l.wp >= 0 && x.w.wperm != nullptr && l.ci % 8 == 0;
}
const __nv_bfloat16* wbf(const Ctx& x, const ConvBn& l) {
  if (use_tm(x, l)) return static_cast<const __nv_bfloat16*>(x.w.wperm) + l.wp;
  return static_cast<const __nv_bfloat16*>(x.shadow) + l.w;
}

// bf16 columns of a layer for the tensor-core path (*cols_out = the GEMM operand): a 1x1 stride-1
// convolution's columns ARE its NHWC input -- used in place when the activations are bf16, one
// vectorised cast otherwise; other shapes: 8 columns per thread.
int tc_columns(const Ctx& x, const ConvBn& l, const float* in, int in_nchw, __nv_bfloat16* cols,
               const __nv_bfloat16** cols_out) {
  const size_t M = static_cast<size_t>(rows_of(x.c, l.hout));
  const int K = l.ci * l.k * l.k;
  *cols_out = cols;
  if (l.k == 1 && l.stride == 1 && !in_nchw) {
    if (x.act16) {
      *cols_out = as16(in);
      return 0;
    }
    return cast_f32_to_bf16(in, cols, M * K, x.st);
  }
  if (!in_nchw && use_tm(x, l)) {
    const bool small = M * (K / 8) + 256ull * 16 * sm_count() < (1ull << 32);  // grid-stride overshoot included
    const int nb = blocks_for(M * (K / 8), 256);
    if (x.act16 && small)
      im2col_tm_kernel<bf16_t, unsigned><<<nb, 256, 0, x.st>>>(
          as16(in), x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
    else if (x.act16)
      im2col_tm_kernel<bf16_t, size_t><<<nb, 256, 0, x.st>>>(
          as16(in), x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
    else if (small)
      im2col_tm_kernel<float, unsigned><<<nb, 256, 0, x.st>>>(
          in, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
    else
      im2col_tm_kernel<float, size_t><<<nb, 256, 0, x.st>>>(
          in, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
    RN_CHECK_LAUNCH();
    return 0;
  }
  if (!in_nchw && K % 8 == 0) {
    if (x.act16)
      im2col_bf16x8_kernel<bf16_t><<<blocks_for(M * (K / 8), 256), 256, 0, x.st>>>(
          as16(in), x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
    else
      im2col_bf16x8_kernel<float><<<blocks_for(M * (K / 8), 256), 256, 0, x.st>>>(
          in, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
    RN_CHECK_LAUNCH();
    return 0;
  }
  if (x.act16 && !in_nchw) return MMU_ERR_SHAPE;  // bf16 activations need K % 8 == 0 everywhere
  im2col_kernel<__nv_bfloat16><<<blocks_for(M * K, 256), 256, 0, x.st>>>(
      in, in_nchw, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, cols);
  RN_CHECK_LAUNCH();
  return 0;
}

int conv_bn_fwd(const Ctx& x, const ConvBn& l, const CbWs& o, const float* in, int in_nchw,
                const float* residual, int relu) {
  const int M = static_cast<int>(rows_of(x.c, l.hout)), K = l.ci * l.k * l.k;
  GemmProblem p{M, l.co, K, 0, 0, 1};
  if (use_tc_padded(x, l, in_nchw)) {
    const int Kp = (K + 7) / 8 * 8;
    __nv_bfloat16* cols = reinterpret_cast<__nv_bfloat16*>(x.w.cols);
    __nv_bfloat16* wp = static_cast<__nv_bfloat16*>(x.w.pad_w);
    pad_weights_bf16_kernel<<<(l.co * Kp + 255) / 256, 256, 0, x.st>>>(x.params + l.w, l.co, K, Kp, wp);
    RN_CHECK_LAUNCH();
    im2col_nchw_pad_bf16x8_kernel<<<blocks_for(static_cast<size_t>(M) * (Kp / 8), 256), 256, 0, x.st>>>(
        in, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, Kp, cols);
    RN_CHECK_LAUNCH();
    GemmProblem pp{M, l.co, Kp, 0, 0, 1};
    GemmEpilogue e = store_epi(o.t, l.co, nullptr);
    e.out_bf16 = x.act16;
    RN_TRY(gemm_bf16_launch(cols, Kp, wp, Kp, pp, e, x.st));
  } else if (use_tc(x, l)) {
    if (use_tm(x, l)) {  // tap-major weight copy (kept for the backward of the same step)
      permute_weights_kernel<<<(l.co * K + 255) / 256, 256, 0, x.st>>>(
          x.params + l.w, l.co, l.ci, l.k * l.k, static_cast<__nv_bfloat16*>(x.w.wperm) + l.wp);
      RN_CHECK_LAUNCH();
    }
    const __nv_bfloat16* cols = nullptr;
    RN_TRY(tc_columns(x, l, in, in_nchw, reinterpret_cast<__nv_bfloat16*>(x.w.cols), &cols));
    GemmEpilogue e = store_epi(o.t, l.co, nullptr);
    e.out_bf16 = x.act16;
    RN_TRY(gemm_bf16_launch(cols, K, wbf(x, l), K, p, e, x.st));
  } else {
    if (x.act16) return MMU_ERR_SHAPE;
    RN_TRY(f32_columns(x, l, in, in_nchw, x.w.cols));
    RN_TRY(gemm_f32_launch(x.w.cols, K, x.params + l.w, K, p, store_epi(o.t, l.co, nullptr), x.st));
  }
  if (x.training) {
    if (cudaMemsetAsync(x.w.sums, 0, 2 * l.co * sizeof(double), x.st) != cudaSuccess) return MMU_ERR_CUDA;
    const ColGeo cg = col_geo(l.co);
    const int rpb = stat_rows_per_block(M, cg);
    const int gy = (M + rpb - 1) / rpb;
    if (x.act16) bn_sums_kernel<bf16_t><<<dim3(cg.gx, gy), 256, 0, x.st>>>(as16(o.t), M, l.co, rpb, cg.tx, x.w.sums);
    else bn_sums_kernel<float><<<dim3(cg.gx, gy), 256, 0, x.st>>>(o.t, M, l.co, rpb, cg.tx, x.w.sums);
    RN_CHECK_LAUNCH();
    bn_finalize_kernel<<<(l.co + 127) / 128, 128, 0, x.st>>>(x.w.sums, M, l.co, 0.1f, o.mean, o.rstd,
                                                              x.stats + l.stat, x.stats + l.stat + (l.co + 63) / 64 * 64);
    RN_CHECK_LAUNCH();
  } else {
    bn_eval_stats_kernel<<<(l.co + 127) / 128, 128, 0, x.st>>>(x.stats + l.stat, x.stats + l.stat + (l.co + 63) / 64 * 64,
                                                                l.co, o.mean, o.rstd);
    RN_CHECK_LAUNCH();
  }
  const size_t n4 = static_cast<size_t>(M) * l.co / 4;
  if (x.act16)
    bn_apply_kernel<bf16_t><<<blocks_for(n4, 256), 256, 0, x.st>>>(
        as16(o.t), o.mean, o.rstd, x.params + l.g, x.params + l.b, as16(residual), relu, as16(o.out), n4, l.co);
  else
    bn_apply_kernel<float><<<blocks_for(n4, 256), 256, 0, x.st>>>(o.t, o.mean, o.rstd, x.params + l.g,
                                                                  x.params + l.b, residual, relu, o.out, n4, l.co);
  RN_CHECK_LAUNCH();
  return 0;
}

// dout: gradient w.r.t. this layer's output `o.out`.  relu: the forward applied ReLU.
// Produces (or accumulates into) din, the gradient w.r.t. the layer input (NHWC); if
// dres_out != nullptr the masked gradient is also copied there (the residual branch).
int conv_bn_bwd(const Ctx& x, const ConvBn& l, const CbWs& o, const float* in, int in_nchw,
                const float* dout, int relu, float* din, int din_accumulate, float* dres_out,
                const float* dout2 = nullptr) {
  const int M = static_cast<int>(rows_of(x.c, l.hout)), K = l.ci * l.k * l.k;
  if (cudaMemsetAsync(x.w.sums, 0, 2 * l.co * sizeof(double), x.st) != cudaSuccess) return MMU_ERR_CUDA;
  float* dyb = dres_out != nullptr ? dres_out : x.w.dyb;
  const ColGeo cg = col_geo(l.co);
  const int rpb = stat_rows_per_block(M, cg);
  const int gy = (M + rpb - 1) / rpb;
  if (x.act16)
    bn_bwd_sums_kernel<bf16_t><<<dim3(cg.gx, gy), 256, 0, x.st>>>(
        dout, dout2, relu ? as16(o.out) : nullptr, as16(o.t), o.mean, o.rstd, M, l.co, rpb, cg.tx, dyb, x.w.sums);
  else
    bn_bwd_sums_kernel<float><<<dim3(cg.gx, gy), 256, 0, x.st>>>(
        dout, dout2, relu ? o.out : nullptr, o.t, o.mean, o.rstd, M, l.co, rpb, cg.tx, dyb, x.w.sums);
  RN_CHECK_LAUNCH();
  const size_t n = static_cast<size_t>(M) * l.co;
  const size_t nin = static_cast<size_t>(x.c.B) * l.hin * l.hin * l.ci;
  GemmEpilogue wg{};
  wg.mode = EPI_ATOMIC; wg.out = x.grads + l.w; wg.ld_out = K; wg.alpha = 1.0f;
  if (use_tc_padded(x, l, in_nchw) && din == nullptr) {
    // padded-K tensor-core weight gradient (7x7x3 stem): dW_p[co, Kp] in a scratch, then += into dW
    const int Kp = (K + 7) / 8 * 8;
    __nv_bfloat16* dt = reinterpret_cast<__nv_bfloat16*>(x.w.dt);
    __nv_bfloat16* cols = reinterpret_cast<__nv_bfloat16*>(x.w.cols);
    if (x.act16)
      bn_bwd_apply_kernel<bf16_t, bf16_t><<<blocks_for(n / 4, 256), 256, 0, x.st>>>(
          dyb, as16(o.t), o.mean, o.rstd, x.params + l.g, x.w.sums, M, l.co, dt, x.grads + l.g, x.grads + l.b);
    else
      bn_bwd_apply_kernel<bf16_t, float><<<blocks_for(n / 4, 256), 256, 0, x.st>>>(
          dyb, o.t, o.mean, o.rstd, x.params + l.g, x.w.sums, M, l.co, dt, x.grads + l.g, x.grads + l.b);
    RN_CHECK_LAUNCH();
    im2col_nchw_pad_bf16x8_kernel<<<blocks_for(static_cast<size_t>(M) * (Kp / 8), 256), 256, 0, x.st>>>(
        in, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, Kp, cols);
    RN_CHECK_LAUNCH();
    if (cudaMemsetAsync(x.w.pad_dw, 0, static_cast<size_t>(l.co) * Kp * 4, x.st) != cudaSuccess) return MMU_ERR_CUDA;
    GemmEpilogue wgp{};
    wgp.mode = EPI_ATOMIC; wgp.out = x.w.pad_dw; wgp.ld_out = Kp; wgp.alpha = 1.0f;
    int splits = sm_count() / ((Kp + 255) / 256);
    if (const char* e = getenv("MMU_STEM_SPLITS")) splits = atoi(e);
    const int kb = (M + 63) / 64;
    if (splits > kb / 4) splits = kb / 4;
    if (splits < 1) splits = 1;
    GemmProblem p{l.co, Kp, M, 1, 1, splits};
    RN_TRY(gemm_bf16_launch(dt, l.co, cols, Kp, p, wgp, x.st));
    unpad_add_kernel<<<(l.co * K + 255) / 256, 256, 0, x.st>>>(x.w.pad_dw, l.co, K, Kp, x.grads + l.w);
    RN_CHECK_LAUNCH();
    return 0;
  }
  if (use_tc(x, l)) {
    // tensor-core path: dt and the recomputed columns in bf16, fp32 accumulation / outputs
    __nv_bfloat16* dt = reinterpret_cast<__nv_bfloat16*>(x.w.dt);
    __nv_bfloat16* cols = reinterpret_cast<__nv_bfloat16*>(x.w.cols);
    __nv_bfloat16* dcols = reinterpret_cast<__nv_bfloat16*>(x.w.dcols);
    if (x.act16)
      bn_bwd_apply_kernel<bf16_t, bf16_t><<<blocks_for(n / 4, 256), 256, 0, x.st>>>(
          dyb, as16(o.t), o.mean, o.rstd, x.params + l.g, x.w.sums, M, l.co, dt, x.grads + l.g, x.grads + l.b);
    else
      bn_bwd_apply_kernel<bf16_t, float><<<blocks_for(n / 4, 256), 256, 0, x.st>>>(
          dyb, o.t, o.mean, o.rstd, x.params + l.g, x.w.sums, M, l.co, dt, x.grads + l.g, x.grads + l.b);
    RN_CHECK_LAUNCH();
    const __nv_bfloat16* wcols = nullptr;
    RN_TRY(tc_columns(x, l, in, in_nchw, cols, &wcols));
    // Split-K so that the few [co x K] output tiles fill the machine: a layer3 convolution has 18
    // tiles and 98 k-blocks (an eighth of the SMs busy without a split).  Wave-aware choice as in
    // engine.cu: the tile count x split lands just under a whole number of waves.
    const bool pair = l.co >= 512 && K > 128;
    const int tiles = ((l.co + (pair ? 255 : 127)) / (pair ? 256 : 128)) * ((K + 255) / 256);
    const int units = pair ? sm_count() / 2 : sm_count();
    int smax = ((M + 63) / 64) / 4;
    if (smax > 48) smax = 48;
    if (smax < 1) smax = 1;
    int splits = 1;
    double best = 0.0;
    for (int sp = 1; sp <= smax; ++sp) {
      const int t = tiles * sp;
      const double eff = static_cast<double>(t) / (static_cast<double>((t + units - 1) / units) * units);
      if (eff > best + 0.02) { best = eff; splits = sp; }
    }
    GemmProblem p{l.co, K, M, 1, 1, splits};
    if (use_tm(x, l)) {  // dW in tap-major order into the scratch, then += into the OIHW gradient
      if (cudaMemsetAsync(x.w.dwperm, 0, static_cast<size_t>(l.co) * K * 4, x.st) != cudaSuccess) return MMU_ERR_CUDA;
      GemmEpilogue wgp = wg;
      wgp.out = x.w.dwperm;
      RN_TRY(gemm_bf16_launch(dt, l.co, wcols, K, p, wgp, x.st));
      unpermute_add_kernel<<<(l.co * K + 255) / 256, 256, 0, x.st>>>(x.w.dwperm, l.co, l.ci, l.k * l.k,
                                                                      x.grads + l.w);
      RN_CHECK_LAUNCH();
    } else {
      RN_TRY(gemm_bf16_launch(dt, l.co, wcols, K, p, wg, x.st));
    }
    if (din != nullptr && l.k == 1 && l.stride == 1) {
      // 1x1 stride-1: the input gradient IS dt W -- written (or TMA-reduce-added) straight to din
      GemmProblem q{M, K, l.co, 0, 1, 1};
      GemmEpilogue e = store_epi(din, K, nullptr);
      if (din_accumulate) e.mode = EPI_ATOMIC;
      RN_TRY(gemm_bf16_launch(dt, l.co, wbf(x, l), K, q, e, x.st));
    } else if (din != nullptr) {
      GemmProblem q{M, K, l.co, 0, 1, 1};
      GemmEpilogue e = store_epi(reinterpret_cast<float*>(dcols), K, nullptr);
      e.out_bf16 = 1;
      RN_TRY(gemm_bf16_launch(dt, l.co, wbf(x, l), K, q, e, x.st));
      if (use_tm(x, l) && nin / 8 + 256ull * 16 * sm_count() < (1ull << 32))
        col2im_tm_kernel<unsigned><<<blocks_for(nin / 8, 256), 256, 0, x.st>>>(
            dcols, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, din, din_accumulate);
      else if (use_tm(x, l))
        col2im_tm_kernel<size_t><<<blocks_for(nin / 8, 256), 256, 0, x.st>>>(
            dcols, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, din, din_accumulate);
      else
        col2im_kernel<__nv_bfloat16><<<blocks_for(nin, 256), 256, 0, x.st>>>(
            dcols, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, din, din_accumulate);
      RN_CHECK_LAUNCH();
    }
    return 0;
  }
  if (x.act16) return MMU_ERR_SHAPE;
  bn_bwd_apply_kernel<float, float><<<blocks_for(n / 4, 256), 256, 0, x.st>>>(
      dyb, o.t, o.mean, o.rstd, x.params + l.g, x.w.sums, M, l.co, x.w.dt, x.grads + l.g, x.grads + l.b);
  RN_CHECK_LAUNCH();
  // weight gradient: dW[co, K] += dt^T cols   (cols recomputed: 9x cheaper than keeping them)
  RN_TRY(f32_columns(x, l, in, in_nchw, x.w.cols));
  {
    int splits = M / 2048;
    if (splits < 1) splits = 1;
    if (splits > 32) splits = 32;
    GemmProblem p{l.co, K, M, 1, 1, splits};
    RN_TRY(gemm_f32_launch(x.w.dt, l.co, x.w.cols, K, p, wg, x.st));
  }
  if (din != nullptr) {
    // input gradient: dcols[M, K] = dt W, then gather back to pixels
    GemmProblem p{M, K, l.co, 0, 1, 1};
    RN_TRY(gemm_f32_launch(x.w.dt, l.co, x.params + l.w, K, p, store_epi(x.w.dcols, K, nullptr), x.st));
    col2im_kernel<float><<<blocks_for(nin, 256), 256, 0, x.st>>>(
        x.w.dcols, x.c.B, l.hin, l.hin, l.ci, l.k, l.stride, l.pad, l.hout, l.hout, din, din_accumulate);
    RN_CHECK_LAUNCH();
  }
  return 0;
}

}  // namespace

// =========================================================================== public
int resnet_param_table(const ResNetConfig& c, ParamEntry* out, int max_entries) {
  Net n;
  int np = 0;
  RN_TRY(build(c, &n, out, max_entries, nullptr, 0, &np, nullptr));
  return np;
}
int resnet_stat_table(const ResNetConfig& c, ParamEntry* out, int max_entries) {
  Net n;
  int ns = 0;
  RN_TRY(build(c, &n, nullptr, 0, out, max_entries, nullptr, &ns));
  return ns;
}
long long resnet_param_count(const ResNetConfig& c) {
  Net n;
  if (build(c, &n, nullptr, 0, nullptr, 0, nullptr, nullptr) != 0) return MMU_ERR_SHAPE;
  return n.n_params;
}
long long resnet_stat_count(const ResNetConfig& c) {
  Net n;
  if (build(c, &n, nullptr, 0, nullptr, 0, nullptr, nullptr) != 0) return MMU_ERR_SHAPE;
  return n.n_stats;
}
long long resnet_workspace_bytes(const ResNetConfig& c, int training) {
  Net n;
  if (build(c, &n, nullptr, 0, nullptr, 0, nullptr, nullptr) != 0) return MMU_ERR_SHAPE;
  Ws w;
  carve(c, n, training, nullptr, &w);
  return w.bytes;
}

int resnet_forward(const ResNetConfig& c, const float* params, const void* params_bf16, float* stats,
                   const float* x_nchw, void* ws, long long ws_bytes, int training, float* logits,
                   cudaStream_t stream) {
  if (params == nullptr || stats == nullptr || x_nchw == nullptr || ws == nullptr || logits == nullptr)
    return MMU_ERR_ARG;
  Net n;
  RN_TRY(build(c, &n, nullptr, 0, nullptr, 0, nullptr, nullptr));
  Ws w;
  carve(c, n, training, ws, &w);
  if (w.bytes > ws_bytes) return MMU_ERR_WORKSPACE;
  const Ctx x{c, params_bf16, params, stats, nullptr, training, w, stream};
  RN_TRY(conv_bn_fwd(x, n.stem, w.stem, x_nchw, 1, nullptr, 1));
  const float* a = w.stem.out;
  for (int i = 0; i < 4; ++i) {
    RN_TRY(conv_bn_fwd(x, n.c1[i], w.c1[i], a, 0, nullptr, 1));
    const float* residual = a;
    if (i == 2) {
      RN_TRY(conv_bn_fwd(x, n.ds, w.ds, a, 0, nullptr, 0));
      residual = w.ds.out;
    }
    RN_TRY(conv_bn_fwd(x, n.c2[i], w.c2[i], w.c1[i].out, 0, residual, 1));
    a = w.c2[i].out;
  }
  pool_fwd_kernel<<<(c.B * 128 + 255) / 256, 256, 0, stream>>>(a, c.B, 7, 7, 128, 4, w.pooled);
  RN_CHECK_LAUNCH();
  GemmProblem p{c.B, c.E * c.C, 128, 0, 0, 1};
  if ((c.E * c.C) % 4 != 0) return MMU_ERR_SHAPE;
  RN_TRY(gemm_f32_launch(w.pooled, 128, params + n.fc_w, 128, p,
                         store_epi(logits, c.E * c.C, params + n.fc_b), stream));
  return 0;
}

int resnet_backward(const ResNetConfig& c, const float* params, const void* params_bf16, float* stats,
                    const float* x_nchw, void* ws, long long ws_bytes, const float* dlogits,
                    float* grads, cudaStream_t stream) {
  if (params == nullptr || x_nchw == nullptr || ws == nullptr || dlogits == nullptr || grads == nullptr)
    return MMU_ERR_ARG;
  Net n;
  RN_TRY(build(c, &n, nullptr, 0, nullptr, 0, nullptr, nullptr));
  Ws w;
  carve(c, n, 1, ws, &w);
  if (w.bytes > ws_bytes) return MMU_ERR_WORKSPACE;
  const Ctx x{c, params_bf16, params, stats, grads, 1, w, stream};
  const int EC = c.E * c.C;
  // ---- MultiHeadFC: dWfc += dlogits^T pooled ; dbfc += colsum ; dpooled = dlogits Wfc
  {
    GemmProblem p{EC, 128, c.B, 1, 1, 1};
    GemmEpilogue e{};
    e.mode = EPI_ATOMIC; e.out = grads + n.fc_w; e.ld_out = 128; e.alpha = 1.0f;
    RN_TRY(gemm_f32_launch(dlogits, EC, w.pooled, 128, p, e, stream));
    colsum_rows_kernel<<<(EC + 127) / 128, 128, 0, stream>>>(dlogits, c.B, EC, grads + n.fc_b);
    RN_CHECK_LAUNCH();
    GemmProblem q{c.B, 128, EC, 0, 1, 1};
    RN_TRY(gemm_f32_launch(dlogits, EC, params + n.fc_w, 128, q, store_epi(w.dpooled, 128, nullptr), stream));
  }
  float* dcur = w.dact[0];
  float* dnext = w.dact[1];
  {
    const size_t total = static_cast<size_t>(c.B) * 49 * 128;
    pool_bwd_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(w.dpooled, c.B, 7, 7, 128, 4, dcur);
    RN_CHECK_LAUNCH();
  }
  // ---- BasicBlocks in reverse.  dcur = gradient w.r.t. the block output.
  for (int i = 3; i >= 0; --i) {
    const float* blk_in = i == 0 ? w.stem.out : w.c2[i - 1].out;
    // conv2 + bn2 (+residual) + relu: masked gradient also flows along the residual (w.dres)
    RN_TRY(conv_bn_bwd(x, n.c2[i], w.c2[i], w.c1[i].out, 0, dcur, 1, dnext, 0, w.dres));
    // dnext = gradient w.r.t. c1 output (post-ReLU); conv1 + bn1 + relu -> gradient w.r.t. block input
    RN_TRY(conv_bn_bwd(x, n.c1[i], w.c1[i], blk_in, 0, dnext, 1, dcur, 0, nullptr));
    if (i == 2) {
      // downsample branch: conv1x1 + bn, no ReLU; its input gradient adds to dcur
      RN_TRY(conv_bn_bwd(x, n.ds, w.ds, blk_in, 0, w.dres, 0, dcur, 1, nullptr));
    } else {
      const size_t nel = static_cast<size_t>(rows_of(c, n.c1[i].hin)) * n.c1[i].ci;
      add_inplace_kernel<<<blocks_for(nel, 256), 256, 0, stream>>>(dcur, w.dres, nel);
      RN_CHECK_LAUNCH();
    }
  }
  // ---- stem: conv1 + bn1 + relu (no input gradient needed)
  RN_TRY(conv_bn_bwd(x, n.stem, w.stem, x_nchw, 1, dcur, 1, nullptr, 0, nullptr));
  return 0;
}


// ===========================================================================================
// MMBT image encoder (reference src/mmbt.py:15-45 ImageEncoder): torchvision's Bottleneck ResNet
// trunk (resnet152 = layers [3, 8, 36, 3]; children()[:-2], i.e. up to layer4) followed by an
// adaptive average / max pool to num_image_embeds cells, flattened to (B, N, 2048) tokens.
// Built from the same conv + BatchNorm helpers as the FashionMNIST ResNet above (NHWC fp32
// activations, im2col + GEMM; tcgen05 for every conv but the 3-channel stem when a bf16 shadow of
// the parameters is given).
namespace {

constexpr int IE_MAX_CONVS = 208;   // 1 stem + 3 per block + 4 downsamples (resnet152: 155)
constexpr int IE_MAX_BLOCKS = 64;

struct IeBlock { int c1, c2, c3, ds; };  // indices into IeNet::conv (ds = -1: identity shortcut)
struct IeNet {
  ConvBn conv[IE_MAX_CONVS];
  int n_conv;
  IeBlock block[IE_MAX_BLOCKS];
  int n_block;
  long long n_wperm;  // elements of the tap-major weight copies (all 3x3 layers)
  long long max_wk;   // largest co * K among them (weight-gradient scratch)
  int h_pool_in;   // spatial size after the stem's max pool
  int h_out;       // spatial size of the layer4 output
  int c_out;       // 2048
  long long n_params, n_stats;
};

int ie_build(const ImgEncConfig& c, IeNet* net, ParamEntry* ptab, int pmax, ParamEntry* stab, int smax,
             int* n_ptab, int* n_stab) {
  if (c.B < 1 || c.H < 32 || c.pool_h < 1 || c.pool_w < 1 || c.width_per_group < 8 ||
      c.width_per_group % 8 != 0)
    return MMU_ERR_SHAPE;
  int nb = 0;
  for (int i = 0; i < 4; ++i) {
    if (c.layers[i] < 1) return MMU_ERR_SHAPE;
    nb += c.layers[i];
  }
  if (nb > IE_MAX_BLOCKS || 1 + 3 * nb + 4 > IE_MAX_CONVS) return MMU_ERR_SHAPE;
  Tab p{ptab, pmax, 0, 0}, s{stab, smax, 0, 0};
  char nm[96];
  auto bn = [&](const char* prefix, int ch, ConvBn* cb) {
    std::snprintf(nm, sizeof(nm), "%s.weight", prefix); cb->g = p.add(nm, ch, 0);
    std::snprintf(nm, sizeof(nm), "%s.bias", prefix);   cb->b = p.add(nm, ch, 0);
    std::snprintf(nm, sizeof(nm), "%s.running_mean", prefix); cb->stat = s.add(nm, ch, 0);
    std::snprintf(nm, sizeof(nm), "%s.running_var", prefix);  s.add(nm, ch, 0);
  };
  auto conv = [&](const char* name, int ci, int co, int k, int stride, int hin, ConvBn* cb) {
    cb->ci = ci; cb->co = co; cb->k = k; cb->stride = stride; cb->pad = k / 2; cb->hin = hin;
    cb->hout = (hin + 2 * cb->pad - k) / stride + 1;
    cb->w = p.add(name, co, ci * k * k);
  };
  int nc = 0;
  // nn.Sequential(conv1, bn1, relu, maxpool, layer1..layer4): keys model.0, model.1, model.4..7
  conv("model.0.weight", 3, 64, 7, 2, c.H, &net->conv[nc]);
  bn("model.1", 64, &net->conv[nc]);
  ++nc;
  const int h_stem = net->conv[0].hout;
  net->h_pool_in = (h_stem + 2 - 3) / 2 + 1;  // MaxPool2d(3, stride 2, padding 1)
  int h = net->h_pool_in, inpl = 64, blk = 0;
  char pre[80], key[96];
  for (int li = 0; li < 4; ++li) {
    const int planes = 64 << li, width = planes * c.width_per_group / 64, outp = planes * 4;
    for (int bi = 0; bi < c.layers[li]; ++bi) {
      const int stride = (bi == 0 && li > 0) ? 2 : 1;
      std::snprintf(pre, sizeof(pre), "model.%d.%d", 4 + li, bi);
      IeBlock& B = net->block[blk++];
      B.c1 = nc;
      std::snprintf(key, sizeof(key), "%s.conv1.weight", pre); conv(key, inpl, width, 1, 1, h, &net->conv[nc]);
      std::snprintf(key, sizeof(key), "%s.bn1", pre);          bn(key, width, &net->conv[nc]);
      ++nc;
      B.c2 = nc;
      std::snprintf(key, sizeof(key), "%s.conv2.weight", pre); conv(key, width, width, 3, stride, h, &net->conv[nc]);
      std::snprintf(key, sizeof(key), "%s.bn2", pre);          bn(key, width, &net->conv[nc]);
      const int h2 = net->conv[nc].hout;
      ++nc;
      B.c3 = nc;
      std::snprintf(key, sizeof(key), "%s.conv3.weight", pre); conv(key, width, outp, 1, 1, h2, &net->conv[nc]);
      std::snprintf(key, sizeof(key), "%s.bn3", pre);          bn(key, outp, &net->conv[nc]);
      ++nc;
      B.ds = -1;
      if (bi == 0 && (stride != 1 || inpl != outp)) {
        B.ds = nc;
        std::snprintf(key, sizeof(key), "%s.downsample.0.weight", pre);
        conv(key, inpl, outp, 1, stride, h, &net->conv[nc]);
        std::snprintf(key, sizeof(key), "%s.downsample.1", pre);
        bn(key, outp, &net->conv[nc]);
        ++nc;
      }
      inpl = outp;
      h = h2;
    }
  }
  net->n_conv = nc;
  net->n_block = blk;
  net->n_wperm = 0;
  net->max_wk = 0;
  for (int i = 1; i < nc; ++i) {
    ConvBn& l = net->conv[i];
    if (l.k > 1) {
      const long long wk = static_cast<long long>(l.co) * l.ci * l.k * l.k;
      l.wp = net->n_wperm;
      net->n_wperm += (wk + 63) / 64 * 64;
      if (wk > net->max_wk) net->max_wk = wk;
    }
  }
  net->h_out = h;
  net->c_out = inpl;
  // (an adaptive pool grid larger than the map is legal: cells then repeat pixels, as in torch)
  net->n_params = p.cursor;
  net->n_stats = s.cursor;
  if (n_ptab) *n_ptab = p.n;
  if (n_stab) *n_stab = s.n;
  return 0;
}

struct IeWs {
  CbWs cb[IE_MAX_CONVS];
  float* mp;            // max-pooled stem output [B*hp*hp, 64]
  unsigned char* mp_idx;  // window position (0..8) of each maximum
  int* pool_idx;        // adaptive max pool: flat pixel index of each maximum
  float* dact[2];
  Ws shared;            // cols / sums / dcols / dyb / dt / dres used by conv_bn_fwd / conv_bn_bwd
  long long bytes;
};

void ie_carve(const ImgEncConfig& c, const IeNet& n, int training, void* base, IeWs* w) {
  Bump b{static_cast<char*>(base), 0};
  long long max_cols = 0, max_act = 0;
  int max_co = 64;
  for (int i = 0; i < n.n_conv; ++i) {
    const ConvBn& l = n.conv[i];
    const long long m = static_cast<long long>(c.B) * l.hout * l.hout;
    CbWs& o = w->cb[i];
    o.t = b.take<float>(m * l.co * 4);
    o.mean = b.take<float>(l.co * 4);
    o.rstd = b.take<float>(l.co * 4);
    o.out = b.take<float>(m * l.co * 4);
    const long long cols = m * l.ci * l.k * l.k;
    if (cols > max_cols) max_cols = cols;
    if (m * l.co > max_act) max_act = m * l.co;
    const long long min_ = static_cast<long long>(c.B) * l.hin * l.hin * l.ci;
    if (min_ > max_act) max_act = min_;
    if (l.co > max_co) max_co = l.co;
  }
  const long long mp_el = static_cast<long long>(c.B) * n.h_pool_in * n.h_pool_in * 64;
  w->mp = b.take<float>(mp_el * 4);
  w->mp_idx = b.take<unsigned char>(mp_el);
  w->pool_idx = b.take<int>(static_cast<long long>(c.B) * c.pool_h * c.pool_w * n.c_out * 4);
  std::memset(&w->shared, 0, sizeof(w->shared));
  {  // padded stem operands (K = 147 -> 152)
    const ConvBn& st = n.conv[0];
    const long long Kp = (st.ci * st.k * st.k + 7) / 8 * 8;
    w->shared.pad_w = b.take<void>(st.co * Kp * 2);
    w->shared.pad_dw = b.take<float>(st.co * Kp * 4);
    const long long m0 = static_cast<long long>(c.B) * st.hout * st.hout;
    if (m0 * Kp > max_cols) max_cols = m0 * Kp;
  }
  w->shared.cols = b.take<float>(max_cols * 4);
  w->shared.sums = b.take<double>(2LL * max_co * 8);
  w->shared.wperm = b.take<void>(n.n_wperm * 2);
  w->shared.dwperm = training ? b.take<float>(n.max_wk * 4) : nullptr;
  if (training) {
    w->shared.dcols = b.take<float>(max_cols * 4);
    w->shared.dyb = b.take<float>(max_act * 4);
    w->shared.dt = b.take<float>(max_act * 4);
    w->shared.dres = b.take<float>(max_act * 4);
    w->dact[0] = b.take<float>(max_act * 4);
    w->dact[1] = b.take<float>(max_act * 4);
  } else {
    w->dact[0] = w->dact[1] = nullptr;
  }
  w->bytes = b.off;
}

// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC; idx = window tap (ky*3+kx) of the maximum
// (first maximum in scan order, like torch's CPU kernel).
template <typename TA>
__global__ void maxpool_fwd_kernel(const TA* __restrict__ a, int B, int H, int Ho, int C,
                                   TA* __restrict__ out, unsigned char* __restrict__ idx) {
  const size_t total = static_cast<size_t>(B) * Ho * Ho * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int xo = static_cast<int>(pix % Ho), yo = static_cast<int>((pix / Ho) % Ho);
    const int b = static_cast<int>(pix / (static_cast<size_t>(Ho) * Ho));
    float best = -INFINITY;
    int tap = 0;
    for (int ky = 0; ky < 3; ++ky) {
      const int y = yo * 2 + ky - 1;
      if (y < 0 || y >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int x = xo * 2 + kx - 1;
        if (x < 0 || x >= H) continue;
        const float v = get(a + ((static_cast<size_t>(b) * H + y) * H + x) * C + c);
        if (v > best) { best = v; tap = ky * 3 + kx; }
      }
    }
    put(out + i, best);
    idx[i] = static_cast<unsigned char>(tap);
  }
}
// gather form (no atomics): an input pixel collects from the <= 4 windows that contain it
__global__ void maxpool_bwd_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ idx,
                                   int B, int H, int Ho, int C, float* __restrict__ da) {
  const size_t total = static_cast<size_t>(B) * H * H * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int x = static_cast<int>(pix % H), y = static_cast<int>((pix / H) % H);
    const int b = static_cast<int>(pix / (static_cast<size_t>(H) * H));
    float s = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1)) continue;
      const int yo = ty >> 1;
      if (yo >= Ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1)) continue;
        const int xo = tx >> 1;
        if (xo >= Ho) continue;
        const size_t o = ((static_cast<size_t>(b) * Ho + yo) * Ho + xo) * C + c;
        if (idx[o] == ky * 3 + kx) s += dout[o];
      }
    }
    da[i] = s;
  }
}
// AdaptiveAvgPool2d / AdaptiveMaxPool2d((ph, pw)) on the (B, H, H, C) NHWC map, written as tokens
// (B, ph*pw, C): cell (i, j) covers rows [floor(i H / ph), ceil((i+1) H / ph)), likewise columns.
template <typename TA>
__global__ void adaptive_pool_fwd_kernel(const TA* __restrict__ a, int B, int H, int C, int ph, int pw,
                                         int is_max, float* __restrict__ tok, int* __restrict__ arg) {
  const size_t total = static_cast<size_t>(B) * ph * pw * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int cell = static_cast<int>((i / C) % (ph * pw));
    const int b = static_cast<int>(i / (static_cast<size_t>(C) * ph * pw));
    const int ci = cell / pw, cj = cell % pw;
    const int y0 = (ci * H) / ph, y1 = ((ci + 1) * H + ph - 1) / ph;
    const int x0 = (cj * H) / pw, x1 = ((cj + 1) * H + pw - 1) / pw;
    float s = is_max ? -INFINITY : 0.f;
    int best = 0;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const float v = get(a + ((static_cast<size_t>(b) * H + y) * H + x) * C + c);
        if (is_max) { if (v > s) { s = v; best = y * H + x; } }
        else s += v;
      }
    tok[i] = is_max ? s : s / ((y1 - y0) * (x1 - x0));
    if (is_max) arg[i] = best;
  }
}
__global__ void adaptive_pool_bwd_kernel(const float* __restrict__ dtok, const int* __restrict__ arg, int B,
                                         int H, int C, int ph, int pw, int is_max, float* __restrict__ da) {
  const size_t total = static_cast<size_t>(B) * H * H * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int x = static_cast<int>(pix % H), y = static_cast<int>((pix / H) % H);
    const int b = static_cast<int>(pix / (static_cast<size_t>(H) * H));
    float s = 0.f;
    for (int ci = 0; ci < ph; ++ci) {
      const int y0 = (ci * H) / ph, y1 = ((ci + 1) * H + ph - 1) / ph;
      if (y < y0 || y >= y1) continue;
      for (int cj = 0; cj < pw; ++cj) {
        const int x0 = (cj * H) / pw, x1 = ((cj + 1) * H + pw - 1) / pw;
        if (x < x0 || x >= x1) continue;
        const size_t o = ((static_cast<size_t>(b) * ph + ci) * pw + cj) * C + c;
        if (is_max) { if (arg[o] == y * H + x) s += dtok[o]; }
        else s += dtok[o] / ((y1 - y0) * (x1 - x0));
      }
    }
    da[i] = s;
  }
}

ResNetConfig ie_rc(const ImgEncConfig& c) { return ResNetConfig{c.B, 3, c.H, c.H, 1, 1}; }

}  // namespace

int imgenc_param_table(const ImgEncConfig& c, ParamEntry* out, int max_entries) {
  IeNet* n = new IeNet;
  int np = 0;
  const int rc = ie_build(c, n, out, max_entries, nullptr, 0, &np, nullptr);
  delete n;
  return rc != 0 ? rc : np;
}
int imgenc_stat_table(const ImgEncConfig& c, ParamEntry* out, int max_entries) {
  IeNet* n = new IeNet;
  int ns = 0;
  const int rc = ie_build(c, n, nullptr, 0, out, max_entries, nullptr, &ns);
  delete n;
  return rc != 0 ? rc : ns;
}
long long imgenc_param_count(const ImgEncConfig& c) {
  IeNet* n = new IeNet;
  const int rc = ie_build(c, n, nullptr, 0, nullptr, 0, nullptr, nullptr);
  const long long r = rc != 0 ? rc : n->n_params;
  delete n;
  return r;
}
long long imgenc_stat_count(const ImgEncConfig& c) {
  IeNet* n = new IeNet;
  const int rc = ie_build(c, n, nullptr, 0, nullptr, 0, nullptr, nullptr);
  const long long r = rc != 0 ? rc : n->n_stats;
  delete n;
  return r;
}
long long imgenc_workspace_bytes(const ImgEncConfig& c, int training) {
  IeNet* n = new IeNet;
  IeWs* w = new IeWs;
  long long r = ie_build(c, n, nullptr, 0, nullptr, 0, nullptr, nullptr);
  if (r == 0) {
    ie_carve(c, *n, training, nullptr, w);
    r = w->bytes;
  }
  delete n;
  delete w;
  return r;
}

namespace {
int ie_forward(const ImgEncConfig& c, const IeNet& n, IeWs& w, const float* params, const void* shadow,
               float* stats, const float* x_nchw, int training, float* tokens, cudaStream_t stream) {
  const ResNetConfig rc = ie_rc(c);
  // tensor-core mode keeps every activation in bf16 (halves the BatchNorm / column traffic, and a
  // 1x1 convolution reads its input in place)
  const int act16 = shadow != nullptr;
  const Ctx x{rc, shadow, params, stats, nullptr, training, w.shared, stream, act16};
  RN_TRY(conv_bn_fwd(x, n.conv[0], w.cb[0], x_nchw, 1, nullptr, 1));
  {
    const size_t total = static_cast<size_t>(c.B) * n.h_pool_in * n.h_pool_in * 64;
    if (act16)
      maxpool_fwd_kernel<bf16_t><<<blocks_for(total, 256), 256, 0, stream>>>(
          as16(w.cb[0].out), c.B, n.conv[0].hout, n.h_pool_in, 64, as16(w.mp), w.mp_idx);
    else
      maxpool_fwd_kernel<float><<<blocks_for(total, 256), 256, 0, stream>>>(w.cb[0].out, c.B, n.conv[0].hout,
                                                                           n.h_pool_in, 64, w.mp, w.mp_idx);
    RN_CHECK_LAUNCH();
  }
  const float* a = w.mp;
  for (int bi = 0; bi < n.n_block; ++bi) {
    const IeBlock& B = n.block[bi];
    RN_TRY(conv_bn_fwd(x, n.conv[B.c1], w.cb[B.c1], a, 0, nullptr, 1));
    RN_TRY(conv_bn_fwd(x, n.conv[B.c2], w.cb[B.c2], w.cb[B.c1].out, 0, nullptr, 1));
    const float* residual = a;
    if (B.ds >= 0) {
      RN_TRY(conv_bn_fwd(x, n.conv[B.ds], w.cb[B.ds], a, 0, nullptr, 0));
      residual = w.cb[B.ds].out;
    }
    RN_TRY(conv_bn_fwd(x, n.conv[B.c3], w.cb[B.c3], w.cb[B.c2].out, 0, residual, 1));
    a = w.cb[B.c3].out;
  }
  const size_t total = static_cast<size_t>(c.B) * c.pool_h * c.pool_w * n.c_out;
  if (act16)
    adaptive_pool_fwd_kernel<bf16_t><<<blocks_for(total, 256), 256, 0, stream>>>(
        as16(a), c.B, n.h_out, n.c_out, c.pool_h, c.pool_w, c.pool_max, tokens, w.pool_idx);
  else
    adaptive_pool_fwd_kernel<float><<<blocks_for(total, 256), 256, 0, stream>>>(
        a, c.B, n.h_out, n.c_out, c.pool_h, c.pool_w, c.pool_max, tokens, w.pool_idx);
  RN_CHECK_LAUNCH();
  return 0;
}

int ie_backward(const ImgEncConfig& c, const IeNet& n, IeWs& w, const float* params, const void* shadow,
                float* stats, const float* x_nchw, const float* dtokens, float* grads, cudaStream_t stream) {
  const ResNetConfig rc = ie_rc(c);
  const Ctx x{rc, shadow, params, stats, grads, 1, w.shared, stream, shadow != nullptr};
  float* dcur = w.dact[0];
  float* dnext = w.dact[1];
  {
    const size_t total = static_cast<size_t>(c.B) * n.h_out * n.h_out * n.c_out;
    adaptive_pool_bwd_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(
        dtokens, w.pool_idx, c.B, n.h_out, n.c_out, c.pool_h, c.pool_w, c.pool_max, dcur);
    RN_CHECK_LAUNCH();
  }
  // `pending` = identity-shortcut gradient of the block processed last, still to be added to dcur:
  // it is folded into the next conv3's BatchNorm-backward pass instead of a separate add.
  const float* pending = nullptr;
  for (int bi = n.n_block - 1; bi >= 0; --bi) {
    const IeBlock& B = n.block[bi];
    const float* blk_in = bi == 0 ? w.mp : w.cb[n.block[bi - 1].c3].out;
    // conv3 + bn3 + shortcut + relu: the masked gradient also flows along the shortcut (dres)
    RN_TRY(conv_bn_bwd(x, n.conv[B.c3], w.cb[B.c3], w.cb[B.c2].out, 0, dcur, 1, dnext, 0, w.shared.dres, pending));
    pending = nullptr;
    RN_TRY(conv_bn_bwd(x, n.conv[B.c2], w.cb[B.c2], w.cb[B.c1].out, 0, dnext, 1, dcur, 0, nullptr));
    RN_TRY(conv_bn_bwd(x, n.conv[B.c1], w.cb[B.c1], blk_in, 0, dcur, 1, dnext, 0, nullptr));
    if (B.ds >= 0) {
      RN_TRY(conv_bn_bwd(x, n.conv[B.ds], w.cb[B.ds], blk_in, 0, w.shared.dres, 0, dnext, 1, nullptr));
    } else {
      pending = w.shared.dres;
    }
    float* t = dcur; dcur = dnext; dnext = t;
  }
  if (pending != nullptr) {  // only if the very first block had an identity shortcut
    const ConvBn& l = n.conv[n.block[0].c1];
    const size_t nel = static_cast<size_t>(c.B) * l.hin * l.hin * l.ci;
    add_inplace_kernel<<<blocks_for(nel, 256), 256, 0, stream>>>(dcur, pending, nel);
    RN_CHECK_LAUNCH();
  }
  {
    const int hs = n.conv[0].hout;
    const size_t total = static_cast<size_t>(c.B) * hs * hs * 64;
    maxpool_bwd_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(dcur, w.mp_idx, c.B, hs, n.h_pool_in, 64,
                                                                    dnext);
    RN_CHECK_LAUNCH();
  }
  return conv_bn_bwd(x, n.conv[0], w.cb[0], x_nchw, 1, dnext, 1, nullptr, 0, nullptr);
}
}  // namespace

int imgenc_forward(const ImgEncConfig& c, const float* params, const void* params_bf16, float* stats,
                   const float* x_nchw, void* ws, long long ws_bytes, int training, float* tokens,
                   cudaStream_t stream) {
  if (params == nullptr || stats == nullptr || x_nchw == nullptr || ws == nullptr || tokens == nullptr)
    return MMU_ERR_ARG;
  IeNet* n = new IeNet;
  IeWs* w = new IeWs;
  int rc = ie_build(c, n, nullptr, 0, nullptr, 0, nullptr, nullptr);
  if (rc == 0) {
    ie_carve(c, *n, training, ws, w);
    rc = w->bytes > ws_bytes ? MMU_ERR_WORKSPACE
                             : ie_forward(c, *n, *w, params, params_bf16, stats, x_nchw, training, tokens, stream);
  }
  delete n;
  delete w;
  return rc;
}

int imgenc_backward(const ImgEncConfig& c, const float* params, const void* params_bf16, float* stats,
                    const float* x_nchw, void* ws, long long ws_bytes, const float* dtokens, float* grads,
                    cudaStream_t stream) {
  if (params == nullptr || x_nchw == nullptr || ws == nullptr || dtokens == nullptr || grads == nullptr)
    return MMU_ERR_ARG;
  IeNet* n = new IeNet;
  IeWs* w = new IeWs;
  int rc = ie_build(c, n, nullptr, 0, nullptr, 0, nullptr, nullptr);
  if (rc == 0) {
    ie_carve(c, *n, 1, ws, w);
    rc = w->bytes > ws_bytes ? MMU_ERR_WORKSPACE
                             : ie_backward(c, *n, *w, params, params_bf16, stats, x_nchw, dtokens, grads, stream);
  }
  delete n;
  delete w;
  return rc;
}

}  // namespace mmu
