// Fused batch-axis attention forward (eval): ONE kernel per layer instead of the softmax-epilogue
// GEMM P = softmax(Q K^T / sqrt(hd)) followed by the batched GEMM O = P V -- the bf16 probabilities
// (2 * B * Bp bytes per (position, head) problem written and read back: 64 KB of the 320 KB the two
// kernels move at B = 128) never exist in HBM.  Reference call site: src/model.py:193,205-207
// (nn.MultiheadAttention with batch_first=False fed (B, L, D): attention ACROSS the mini-batch, one
// B x B problem per token position and head).
//
// head_dim 256, B <= 128.  A persistent CTA (one per SM) walks the (position l, head h) problems:
//   TMA : Q, K [B x 256] as four K-major [128 x 64] SWIZZLE_128B boxes each (rows >= B zero-fill),
//         V [B x 256] as two 64-key blocks of four MN-major [64 x 64] boxes -> 192 KB
//   MMA1: S = Q K^T  (tcgen05.mma 128 x 128 x 16, 16 steps) -> TMEM columns 0..127
//   softmax: 8 warps; warp w owns TMEM lane quadrant w % 4 (a thread = one query row) and 64 of the
//         128 key columns; one TMEM pass (the 64 scores stay in registers), (max, sum) exchanged with
//         the warp holding the other half through shared memory, p -> bf16 -> the K-major
//         SWIZZLE_128B A-operand tile of MMA2 in shared memory (32 KB)
//   MMA2: O = P V    (128 x 256 x 16, 8 steps) -> TMEM columns 128..383
//   epilogue: O -> bf16 -> [32 x 64 B] SWIZZLE_64B staging boxes (they reuse the P tile, dead once
//         MMA2 has retired) -> TMA tensor stores into out[(l, b), h * 256 ...] (rows >= B clipped)
// The loads of problem i + 1 start as soon as the MMAs of problem i have released their operands
// (Q, K after MMA1; V after MMA2), so the kernel streams: it is HBM bound (256 KB per problem
// against ~4 400 clocks of tensor + softmax work).
#include <cuda_bf16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace mmu {
namespace battn {

constexpr int HD = 256;
constexpr int BM = 128;           // queries = keys = samples of the mini-batch (padded)
constexpr int THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 softmax / epilogue
constexpr int KM_BOX = BM * 64 * 2;               // K-major [128 x 64] bf16 box, 16 KB
constexpr int MN_BOX = 64 * 64 * 2;               // MN-major [64 keys x 64 cols] box, 8 KB
constexpr int OFF_Q = 0;                          // 4 boxes
constexpr int OFF_K = OFF_Q + 4 * KM_BOX;         // 4 boxes
constexpr int OFF_V = OFF_K + 4 * KM_BOX;         // 2 key blocks x 4 boxes
constexpr int OFF_P = OFF_V + 8 * MN_BOX;         // 2 K-major boxes (64 keys each) = 16 staging boxes
constexpr int OFF_XCH = OFF_P + 2 * KM_BOX;       // float2 [2][128]
constexpr int OFF_BARS = OFF_XCH + 2 * BM * 8;
constexpr int SMEM_USED = OFF_BARS + 8 * 8 + 16;
constexpr int SMEM_BYTES = 227 * 1024;
static_assert(SMEM_USED + 512 <= SMEM_BYTES, "shared memory budget exceeded");
constexpr int STG_BOX = 32 * 64;                  // [32 rows x 64 B] output staging box

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void nbar(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// SAVE_P (training): the staged probabilities are also TMA-stored to probs[L*H][B][Bp] (what the
// backward reads) straight from the MMA2 operand tile.
template <bool SAVE_P>
__global__ void __launch_bounds__(THREADS, 1)
battn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qk, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_p, int B,
                 int L, int D, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* qk_full = bars + 0;    // TMA -> MMA: Q and K landed
  uint64_t* qk_empty = bars + 1;   // MMA1 retired: Q / K may be overwritten
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;    // MMA2 retired
  uint64_t* s_full = bars + 4;     // scores complete in TMEM
  uint64_t* p_full = bars + 5;     // probabilities staged (8 warp arrivals)
  uint64_t* o_full = bars + 6;     // output complete in TMEM (and the P tile is dead)
  uint64_t* o_empty = bars + 7;    // output drained from TMEM (8 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float2* xch = reinterpret_cast<float2*>(smem + OFF_XCH);
  if (threadIdx.x == 0 && (smem - smem_raw) + SMEM_USED > SMEM_BYTES) {
    printf("mmu: battn dynamic shared memory window is not 1024-byte aligned (offset %d)\n",
           static_cast<int>(smem - smem_raw));
    __trap();
  }
  ptx::pdl_trigger();  // a following tensor-core GEMM may start its prologue on SMs this grid has left
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = L * H;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_qk);
    ptx::prefetch_tmap(&tm_v);
    ptx::prefetch_tmap(&tm_o);
    if (SAVE_P) ptx::prefetch_tmap(&tm_p);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&bars[i], (i == 5 || i == 7) ? 8 : 1);
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem, tmem_o = tmem + 128;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t ph = 0;
      for (int g = blockIdx.x; g < G; g += gridDim.x, ph ^= 1) {
        const int l = g / H, h = g % H;
        ptx::mbar_wait(qk_empty, ph ^ 1);
        ptx::mbar_arrive_expect_tx(qk_full, 8 * KM_BOX);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          ptx::tma_load_3d(smem + OFF_Q + kb * KM_BOX, &tm_qk, qk_full, h * HD + kb * 64, l, 0);
          ptx::tma_load_3d(smem + OFF_K + kb * KM_BOX, &tm_qk, qk_full, D + h * HD + kb * 64, l, 0);
        }
        ptx::mbar_wait(v_empty, ph ^ 1);
        ptx::mbar_arrive_expect_tx(v_full, 8 * MN_BOX);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ptx::tma_load_3d(smem + OFF_V + (kb * 4 + j) * MN_BOX, &tm_v, v_full, 2 * D + h * HD + j * 64, l,
                             kb * 64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = ptx::make_idesc_bf16(BM, BM, 0, 0);
      const uint32_t idesc_o = ptx::make_idesc_bf16(BM, HD, 0, 1);
      const uint32_t sq = ptx::smem_u32(smem + OFF_Q), sk = ptx::smem_u32(smem + OFF_K);
      const uint32_t sv = ptx::smem_u32(smem + OFF_V), sp = ptx::smem_u32(smem + OFF_P);
      uint32_t ph = 0;
      for (int g = blockIdx.x; g < G; g += gridDim.x, ph ^= 1) {
        ptx::mbar_wait(qk_full, ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem_s, ptx::make_smem_desc_sw128(sq + kb * KM_BOX + k * 32, 16u, 1024u),
                           ptx::make_smem_desc_sw128(sk + kb * KM_BOX + k * 32, 16u, 1024u), idesc_s,
                           (kb > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit(qk_empty);
        ptx::umma_commit(s_full);
        ptx::mbar_wait(v_full, ph);
        ptx::mbar_wait(o_empty, ph ^ 1);
        ptx::mbar_wait(p_full, ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem_o, ptx::make_smem_desc_sw128(sp + kb * KM_BOX + k * 32, 16u, 1024u),
                           ptx::make_smem_desc_sw128(sv + kb * 4 * MN_BOX + k * 2048, 8192u, 1024u), idesc_o,
                           (kb > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit(v_empty);
        if constexpr (SAVE_P) {
          // the same tile goes to HBM for the backward; it becomes output staging once o_full
          // fires, so the store must have read it before the barrier is armed
          ptx::tma_store_3d(&tm_p, sp, 0, g, 0);
          if (B > 64) ptx::tma_store_3d(&tm_p, sp + KM_BOX, 64, g, 0);
          ptx::bulk_commit();
          ptx::bulk_wait_read<0>();
        }
        ptx::umma_commit(o_full);
      }
      if constexpr (SAVE_P) ptx::bulk_wait<0>();
    }
  } else {
    // --------------------------------------------------------- softmax + output epilogue
    const int we = warp - 2;       // 0..7
    const int q = warp & 3;        // TMEM lane quadrant (hardware: lanes 32 * (warp % 4) ..)
    const int hf = we >> 2;        // which 64 key columns of S / which 128 columns of O
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_s = tmem_s + lane_addr + static_cast<uint32_t>(hf * 64);
    const uint32_t t_o = tmem_o + lane_addr + static_cast<uint32_t>(hf * 128);
    const float sc = scale * 1.4426950408889634f;
    const uint32_t p_row = ptx::smem_u32(smem + OFF_P) + static_cast<uint32_t>(hf) * KM_BOX +
                           static_cast<uint32_t>(row) * 128u;
    const uint32_t p_sw = static_cast<uint32_t>(row & 7);
    const uint32_t stg0 = ptx::smem_u32(smem + OFF_P) + static_cast<uint32_t>(we) * 2u * STG_BOX;
    const uint32_t stg_row = static_cast<uint32_t>(lane) * 64u;
    const uint32_t stg_sw = static_cast<uint32_t>(lane >> 1) & 3u;
    const int nk = min(max(B - hf * 64, 0), 64);  // valid keys in this warp's column half
    uint32_t ph = 0;
    for (int g = blockIdx.x; g < G; g += gridDim.x, ph ^= 1) {
      const int l = g / H, h = g % H;
      // ---- scores -> registers
      uint32_t r[2][32];
      ptx::mbar_wait(s_full, ph);
      ptx::tc_fence_after();
      ptx::tmem_ld_32x32(t_s, r[0]);
      ptx::tmem_ld_32x32(t_s + 32, r[1]);
      ptx::tmem_ld_wait();
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i < nk) mx = fmaxf(mx, __uint_as_float(r[i >> 5][i & 31]));
      const float off = nk > 0 ? -mx * sc : 0.f;
      float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float e = i < nk ? ex2(fmaf(__uint_as_float(r[i >> 5][i & 31]), sc, off)) : 0.f;
        r[i >> 5][i & 31] = __float_as_uint(e);
        part[i & 3] += e;
      }
      const float sum = (part[0] + part[1]) + (part[2] + part[3]);
      // ---- (max, sum) of the other 64 columns of the same row
      xch[hf * BM + row] = make_float2(mx, sum);
      nbar(2 + q, 64);
      const float2 o = xch[(hf ^ 1) * BM + row];
      const float mall = fmaxf(mx, o.x);
      const float f_self = nk > 0 ? ex2((mx - mall) * sc) : 0.f;
      const float f_other = o.y > 0.f ? ex2((o.x - mall) * sc) : 0.f;
      const float f = f_self / (sum * f_self + o.y * f_other);
      // ---- the P tile doubles as this kernel's output staging: the previous problem's stores
      //      must have read it (every warp's), and the exchange slots may be rewritten after this
      if (lane == 0) ptx::bulk_wait_read<0>();
      nbar(1, 256);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t* e = &r[k >> 2][(k & 3) * 8];
        ptx::sts_v4u(p_row + ((static_cast<uint32_t>(k) ^ p_sw) << 4),
                     pack2(__uint_as_float(e[0]) * f, __uint_as_float(e[1]) * f),
                     pack2(__uint_as_float(e[2]) * f, __uint_as_float(e[3]) * f),
                     pack2(__uint_as_float(e[4]) * f, __uint_as_float(e[5]) * f),
                     pack2(__uint_as_float(e[6]) * f, __uint_as_float(e[7]) * f));
      }
      ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);

      // ---- O = P V -> bf16 -> staging boxes -> TMA stores
      ptx::mbar_wait(o_full, ph);
      ptx::tc_fence_after();
      ptx::tmem_ld_32x32(t_o, r[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::tmem_ld_wait();
        if (c + 1 < 4) {
          ptx::tmem_ld_32x32(t_o + (c + 1) * 32, r[(c + 1) & 1]);
        } else {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(o_empty);
        }
        const uint32_t box = stg0 + static_cast<uint32_t>(c & 1) * STG_BOX;
        if (c >= 2) {  // the store issued two chunks ago has read this box
          if (lane == 0) ptx::bulk_wait_read<1>();
          __syncwarp();
        }
        const uint32_t* v = r[c & 1];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::sts_v4u(box + stg_row + ((static_cast<uint32_t>(k) ^ stg_sw) << 4),
                       pack2(__uint_as_float(v[8 * k]), __uint_as_float(v[8 * k + 1])),
                       pack2(__uint_as_float(v[8 * k + 2]), __uint_as_float(v[8 * k + 3])),
                       pack2(__uint_as_float(v[8 * k + 4]), __uint_as_float(v[8 * k + 5])),
                       pack2(__uint_as_float(v[8 * k + 6]), __uint_as_float(v[8 * k + 7])));
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_4d(&tm_o, box, hf * 128 + c * 32, q * 32, h, l);
          ptx::bulk_commit();
        }
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();  // all stores of this warp have completed at exit
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace battn

// 0: launched; > 0: not applicable (the caller runs the two-kernel path); < 0: error.
// probs != nullptr: the bf16 probabilities [L*H][B][Bp] are also written (training forward).
int fused_batch_attention_fwd(const void* qkv, void* out, void* probs, int B, int L, int D, int H,
                              int pos_major, cudaStream_t stream) {
  using namespace battn;
  static const bool disabled = getenv("MMU_BATTN_UNFUSED") != nullptr;  // A/B switch
  if (disabled || H < 1 || D % H != 0 || D / H != HD || B < 1 || B > BM || L < 1) return 1;
  const long long ld = 3LL * D;
  const long long mid_stride = pos_major ? ld * B : ld;          // position pitch
  const long long outer_stride = pos_major ? ld : ld * L;        // sample pitch
  const int G = L * H, Bp = (B + 7) / 8 * 8;
  CUtensorMap tq, tv, to, tp;
  int rc = make_tmap_bf16_3d(&tq, qkv, ld, L, B, mid_stride, outer_stride, 64, BM);
  if (rc != 0) return rc;
  rc = make_tmap_bf16_3d(&tv, qkv, ld, L, B, mid_stride, outer_stride, 64, 64);
  if (rc != 0) return rc;
  rc = make_tmap_out_4d(&to, out, 1, HD, B, pos_major ? D : static_cast<long long>(L) * D, H, HD, L,
                        pos_major ? static_cast<long long>(B) * D : D);
  if (rc != 0) return rc;
  tp = tq;
  if (probs != nullptr) {
    rc = make_tmap_bf16_3d(&tp, probs, Bp, G, B, static_cast<long long>(B) * Bp, Bp, 64, BM);
    if (rc != 0) return rc;
  }
  static cudaError_t attr_err[2] = {
      cudaFuncSetAttribute(battn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES),
      cudaFuncSetAttribute(battn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)};
  if (attr_err[0] != cudaSuccess || attr_err[1] != cudaSuccess) return MMU_ERR_CUDA;
  const int grid = G < sm_count() ? G : sm_count();
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  if (probs != nullptr)
    battn_fwd_kernel<true><<<grid, THREADS, SMEM_BYTES, stream>>>(tq, tv, to, tp, B, L, D, H, scale);
  else
    battn_fwd_kernel<false><<<grid, THREADS, SMEM_BYTES, stream>>>(tq, tv, to, tp, B, L, D, H, scale);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}

}  // namespace mmu
