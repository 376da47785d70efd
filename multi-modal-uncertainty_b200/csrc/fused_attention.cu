// Fused sequence-axis attention forward for the BERT encoder of the MMBT path (head_dim 64,
// S <= 512): ONE kernel per layer instead of  scores GEMM -> masked softmax -> P V GEMM, and the
// fp32 score matrix (B*H*S*S*4 bytes = 402 MB per layer at B = 32, S = 512) never exists in HBM.
//
// One CTA per (sample b, head h, 128-query tile):
//   TMA : Q tile [128 x 64], K [S x 64] (K-major, 128-key boxes), V [S x 64] (MN-major, 64-key
//         boxes) from the packed qkv buffer -> shared memory (144 KB)
//   MMA1: S = Q K^T, one tcgen05.mma group per 128-key block -> ALL of TMEM (128 lanes x 512 fp32
//         columns hold the whole score tile)
//   softmax: 8 warps; warp w owns TMEM lane quadrant w % 4 (a thread = one query row) and one half
//         of the key columns; pass 1 row max, (pass 2 row sum when P is kept,) last pass
//         p = exp2(s * scale * log2e + mask - max) [* 1/sum] -> bf16 -> a [128 x 64] SWIZZLE_128B
//         K-major tile in shared memory (two buffers per column half); the two halves exchange
//         (max, sum) through shared memory
//   MMA2: O += P_chunk V_chunk as soon as a chunk is staged; the accumulator reuses TMEM columns
//         0..63, which chunk 0's last pass has already drained (chunk 0 is always issued first)
//   TMA store of every P chunk to the probs tensor when the backward needs it (training)
//   epilogue: O (x 1/sum in the 2-pass variant) -> bf16 -> out[b, s, h*64 ...]
// Reference: pytorch_pretrained_bert BertSelfAttention as called from src/mmbt.py:124-128, with the
// additive mask of src/mmbt.py:103-107.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "dropout.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace mmu {
namespace fattn {

constexpr int HD = 64;
constexpr int BQ = 128;          // queries per CTA
constexpr int SMAX = 512;        // keys: the whole row of scores lives in TMEM
constexpr int KB = 128;          // keys per MMA1 block / K box
constexpr int CH = 64;           // keys per P chunk / V box
constexpr int THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 softmax
constexpr int OFF_Q = 0;                          // 16 KB
constexpr int OFF_K = OFF_Q + BQ * HD * 2;        // 4 x 16 KB
constexpr int OFF_V = OFF_K + SMAX * HD * 2;      // 8 x 8 KB
constexpr int OFF_P = OFF_V + SMAX * HD * 2;      // 2 halves x 2 buffers x 16 KB
constexpr int P_BYTES = BQ * CH * 2;
constexpr int OFF_MASK = OFF_P + 4 * P_BYTES;     // 512 floats
constexpr int OFF_XCH = OFF_MASK + SMAX * 4;      // float2 [2][128]
constexpr int OFF_BARS = OFF_XCH + 2 * BQ * 8;
constexpr int SMEM_USED = OFF_BARS + 16 * 8 + 16;
constexpr int SMEM_BYTES = SMEM_USED + 1024;      // slack for the 1024-byte alignment

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void nbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// DROP (training with attention-probability dropout, pytorch_pretrained_bert BertSelfAttention):
// the UNDROPPED rounded probabilities go to the probs tensor (the backward needs them where the
// mask is 0 too), the dropped ones Pd = bf16(P * mask / (1 - p)) feed MMA2.  Each column half then
// uses its two P buffers as (store staging, MMA operand) instead of double-buffering one role;
// element counter = (g * S + query) * S + key, the mask function of csrc/dropout.cuh.
template <bool SAVE_P, bool DROP>
__global__ void __launch_bounds__(THREADS, 1)
fattn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qk, const __grid_constant__ CUtensorMap tm_v,
                 const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_pd,
                 const float* __restrict__ addmask, __nv_bfloat16* __restrict__ out, int B, int S, int D, int H,
                 float scale, int tiles_per_cta, const dropout::Site drop, int save_pd) {
  static_assert(!DROP || SAVE_P, "dropout exists in training only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* kv_full = bars;        // K + V of this (sample, head) landed (once per CTA)
  uint64_t* q_full = bars + 1;     // Q tile landed                       (once per query tile)
  uint64_t* s_full = bars + 2;     // scores complete in TMEM
  uint64_t* o_full = bars + 3;     // O complete in TMEM
  uint64_t* p_full = bars + 4;     // [2 halves][2 buffers]
  uint64_t* p_empty = bars + 8;    // [2][2]
  uint64_t* t_empty = bars + 12;   // TMEM drained by the 8 softmax warps (count 8)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* mask_s = reinterpret_cast<float*>(smem + OFF_MASK);
  float2* xch = reinterpret_cast<float2*>(smem + OFF_XCH);

  // A CTA owns `tiles_per_cta` consecutive 128-query tiles of one (sample b, head h): K and V are
  // loaded ONCE per CTA and shared by its tiles (a CTA per tile re-reads 128 KB of K/V for every
  // 16 KB of Q; a CTA per head leaves 2.6 waves on 148 SMs -- the host picks the split).
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int all_tiles = (S + BQ - 1) / BQ;
  const int splits = (all_tiles + tiles_per_cta - 1) / tiles_per_cta;
  const int g = blockIdx.x / splits;
  const int tile0 = (blockIdx.x % splits) * tiles_per_cta;
  const int q_tiles = min(tiles_per_cta, all_tiles - tile0);
  const int b = g / H, h = g % H;
  const int n_kb = (S + KB - 1) / KB;   // MMA1 blocks
  const int n_ch = (S + CH - 1) / CH;   // P / V chunks

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_qk);
    ptx::prefetch_tmap(&tm_v);
    if (SAVE_P) ptx::prefetch_tmap(&tm_p);
    if (DROP) ptx::prefetch_tmap(&tm_pd);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 12; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::mbar_init(t_empty, 8);
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // additive mask of this sample, pre-multiplied by log2(e); keys >= S are excluded outright
  for (int k = threadIdx.x; k < SMAX; k += THREADS)
    mask_s[k] = k < S ? addmask[static_cast<size_t>(b) * S + k] * 1.4426950408889634f : -INFINITY;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, BQ * HD * 2);
      ptx::tma_load_3d(smem + OFF_Q, &tm_qk, q_full, h * HD, b, tile0 * BQ);
      ptx::mbar_arrive_expect_tx(kv_full, n_kb * KB * HD * 2 + n_ch * CH * HD * 2);
      for (int kb = 0; kb < n_kb; ++kb)
        ptx::tma_load_3d(smem + OFF_K + kb * KB * HD * 2, &tm_qk, kv_full, D + h * HD, b, kb * KB);
      for (int c = 0; c < n_ch; ++c)
        ptx::tma_load_3d(smem + OFF_V + c * CH * HD * 2, &tm_v, kv_full, 2 * D + h * HD, b, c * CH);
      // the next Q tile is fetched as soon as MMA1 of the current one has retired
      for (int qt = 1; qt < q_tiles; ++qt) {
        ptx::mbar_wait(s_full, (qt - 1) & 1);
        ptx::mbar_arrive_expect_tx(q_full, BQ * HD * 2);
        ptx::tma_load_3d(smem + OFF_Q, &tm_qk, q_full, h * HD, b, (tile0 + qt) * BQ);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id1 = ptx::make_idesc_bf16(BQ, KB, 0, 0);
      const uint32_t id2 = ptx::make_idesc_bf16(BQ, HD, 0, 1);
      const uint32_t sq = ptx::smem_u32(smem + OFF_Q);
      ptx::mbar_wait(kv_full, 0);
      for (int qt = 0; qt < q_tiles; ++qt) {
        // ---- MMA1: S[128 x 128*n_kb] = Q K^T  (TMEM must have been drained by the previous tile)
        ptx::mbar_wait(q_full, qt & 1);
        if (qt > 0) ptx::mbar_wait(t_empty, (qt - 1) & 1);
        ptx::tc_fence_after();
        for (int kb = 0; kb < n_kb; ++kb) {
          const uint32_t sk = ptx::smem_u32(smem + OFF_K + kb * KB * HD * 2);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            ptx::umma_bf16(tmem + kb * KB, ptx::make_smem_desc_sw128(sq + k * 32, 16u, 1024u),
                           ptx::make_smem_desc_sw128(sk + k * 32, 16u, 1024u), id1, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(s_full);
        // ---- MMA2: O[128 x 64] += P_chunk V_chunk, chunk 0 first (its columns become the accumulator)
        bool first = true;
        for (int i = 0; i < 4; ++i) {
          for (int hh = 0; hh < 2; ++hh) {
            const int c = hh * 4 + i;
            if (c >= n_ch) continue;
            const int buf = DROP ? 1 : (i & 1);
            // global use index of this P buffer (a half with n chunks uses buffer 0 ceil(n/2) and
            // buffer 1 floor(n/2) times per query tile; DROP: buffer 1 for every chunk): the
            // barrier phase is its parity
            const int n_half = max(0, min(n_ch, hh * 4 + 4) - hh * 4);
            const int use = DROP ? qt * n_half + i : qt * ((n_half - buf + 1) / 2) + (i >> 1);
            ptx::mbar_wait(&p_full[hh * 2 + buf], use & 1);
            ptx::tc_fence_after();
            const uint32_t sp = ptx::smem_u32(smem + OFF_P + (hh * 2 + buf) * P_BYTES);
            const uint32_t sv = ptx::smem_u32(smem + OFF_V + c * CH * HD * 2);
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) {
              ptx::umma_bf16(tmem, ptx::make_smem_desc_sw128(sp + k * 32, 16u, 1024u),
                             ptx::make_smem_desc_sw128(sv + k * 2048, 8192u, 1024u), id2,
                             (first && k == 0) ? 0u : 1u);
            }
            first = false;
            ptx::umma_commit(&p_empty[hh * 2 + buf]);
          }
        }
        ptx::umma_commit(o_full);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps
    const int we = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant (hardware: lanes 32*(warp % 4)..)
    const int hh = we >> 2;            // which half of the key columns
    const int row = q * 32 + lane;     // query row within the tile
    const float sc = scale * 1.4426950408889634f;
    const int c_begin = hh * 4, c_end = min(n_ch, hh * 4 + 4);   // this half's chunks
    const int n_mine = c_end - c_begin;                           // 0..4 (uniform over the half)
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t prow = static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const bool issuer = (q == 2 && lane == 0);  // warp 2 / warp 6: first warp of each half
    uint32_t r[32];
    for (int qt = 0; qt < q_tiles; ++qt) {
      const int q0 = (tile0 + qt) * BQ;
      ptx::mbar_wait(s_full, qt & 1);
      ptx::tc_fence_after();
      // Passes 1 and 2 stream this half's columns 16 at a time through two register buffers: the
      // tcgen05.ld of step i + 1 is in flight while step i is reduced (TMEM load latency, not
      // arithmetic, bounded these passes).
      const int n16 = n_mine * (CH / 16);
      const uint32_t tcol0 = trow + c_begin * CH;
      const float* ms0 = mask_s + c_begin * CH;
      uint32_t ra[16], rb[16];
      auto stream16 = [&](auto&& body) {
        if (n16 > 0) ptx::tmem_ld_32x16(tcol0, ra);
        for (int i = 0; i < n16; i += 2) {
          ptx::tmem_ld_wait();
          if (i + 1 < n16) ptx::tmem_ld_32x16(tcol0 + (i + 1) * 16, rb);
          body(ra, ms0 + i * 16);
          if (i + 1 < n16) {
            ptx::tmem_ld_wait();
            if (i + 2 < n16) ptx::tmem_ld_32x16(tcol0 + (i + 2) * 16, ra);
            body(rb, ms0 + (i + 1) * 16);
          }
        }
      };
      // pass 1: row maximum over this half's columns
      float mx = -INFINITY;
      {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        stream16([&](const uint32_t (&v)[16], const float* ms) {
#pragma unroll
          for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], fmaf(__uint_as_float(v[i]), sc, ms[i]));
        });
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      }
      xch[hh * BQ + row] = make_float2(mx, 0.f);
      nbar(3 + q, 64);
      mx = fmaxf(mx, xch[(hh ^ 1) * BQ + row].x);
      nbar(3 + q, 64);
      float sum = 0.f;
      if (SAVE_P) {  // pass 2: row sum, so that the NORMALISED probabilities can be stored
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        stream16([&](const uint32_t (&v)[16], const float* ms) {
#pragma unroll
          for (int i = 0; i < 16; ++i) s4[i & 3] += ex2f(fmaf(__uint_as_float(v[i]), sc, ms[i]) - mx);
        });
        sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        xch[hh * BQ + row] = make_float2(mx, sum);
        nbar(3 + q, 64);
        sum += xch[(hh ^ 1) * BQ + row].y;
        nbar(3 + q, 64);
      }
      const float inv = SAVE_P ? 1.0f / sum : 1.0f;
      // last pass: probabilities -> bf16 -> swizzled K-major tile -> MMA2 (and the probs tensor)
      for (int c = c_begin; c < c_end; ++c) {
        const int i = c - c_begin, buf = DROP ? 1 : (i & 1);
        const uint32_t pbuf = ptx::smem_u32(smem + OFF_P + (hh * 2 + buf) * P_BYTES);   // MMA2 operand
        const uint32_t pst = DROP ? ptx::smem_u32(smem + OFF_P + (hh * 2) * P_BYTES) : pbuf;  // TMA-store source
        // buffer reuse: its previous tile (this query tile's use i - 2, or the previous query
        // tile's last use of this buffer) must have been consumed by MMA2 and by its TMA store.
        // Completions of p_empty[buf] come in use order, 2 (or 1) per query tile (DROP: one per chunk).
        const int uses_per_tile = DROP ? n_mine : (n_mine - buf + 1) / 2;   // uses of this buffer per tile
        const int use = qt * uses_per_tile + (DROP ? i : (i >> 1));         // global use index of this write
        if (use > 0) {
          ptx::mbar_wait(&p_empty[hh * 2 + buf], (use - 1) & 1);
          if (SAVE_P) {
            if (issuer) {
              if (DROP) ptx::bulk_wait_read<0>();   // one staging buffer: its last store has read it
              else ptx::bulk_wait_read<1>();
            }
            nbar(1 + hh, 128);
          }
        }
        {  // four 16-column steps, the next step's TMEM load in flight behind the current one
          auto emit = [&](const uint32_t (&tv)[16], int step) {
            const float* ms = mask_s + c * CH + step * 16;
            float v[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = ex2f(fmaf(__uint_as_float(tv[t]), sc, ms[t]) - mx) * inv;
            if (!SAVE_P) {
              float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int t = 0; t < 16; ++t) s4[t & 3] += v[t];
              sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint32_t piece = static_cast<uint32_t>(step * 2 + k);
              const uint32_t w[4] = {pack2(v[8 * k], v[8 * k + 1]), pack2(v[8 * k + 2], v[8 * k + 3]),
                                     pack2(v[8 * k + 4], v[8 * k + 5]), pack2(v[8 * k + 6], v[8 * k + 7])};
              ptx::sts_v4u(pst + prow + ((piece ^ sw) << 4), w[0], w[1], w[2], w[3]);
              if constexpr (DROP) {
                // the dropped copy is derived from the ROUNDED probability, as the backward regenerates it
                const unsigned int e0 = (static_cast<unsigned int>(g) * static_cast<unsigned int>(S) +
                                         static_cast<unsigned int>(q0 + row)) * static_cast<unsigned int>(S) +
                                        static_cast<unsigned int>(c * CH + step * 16 + 8 * k);
                uint32_t d[4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  d[t] = pack2(__uint_as_float(w[t] << 16) * drop.mult(e0 + 2 * t),
                               __uint_as_float(w[t] & 0xffff0000u) * drop.mult(e0 + 2 * t + 1));
                ptx::sts_v4u(pbuf + prow + ((piece ^ sw) << 4), d[0], d[1], d[2], d[3]);
              }
            }
          };
          const uint32_t tc = trow + c * CH;
          ptx::tmem_ld_32x16(tc, ra);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x16(tc + 16, rb);
          emit(ra, 0);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x16(tc + 32, ra);
          emit(rb, 1);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x16(tc + 48, rb);
          emit(ra, 2);
          ptx::tmem_ld_wait();
          emit(rb, 3);
        }
        if (c == 0) ptx::tc_fence_before();  // columns 0..63 are drained before MMA2 overwrites them
        ptx::fence_proxy_async();
        nbar(1 + hh, 128);
        if (issuer) {
          ptx::mbar_arrive(&p_full[hh * 2 + buf]);
          if (SAVE_P) {
            ptx::tma_store_3d(&tm_p, pst, c * CH, g, q0);
            // the dropped copy too (dV = Pd^T dO reads it): the MMA operand tile, read concurrently by MMA2
            if (DROP && save_pd) ptx::tma_store_3d(&tm_pd, pbuf, c * CH, g, q0);
            ptx::bulk_commit();
          }
        }
      }
      float inv_o = 1.0f;
      if (!SAVE_P) {
        xch[hh * BQ + row] = make_float2(mx, sum);
        nbar(3 + q, 64);
        sum += xch[(hh ^ 1) * BQ + row].y;
        nbar(3 + q, 64);
        inv_o = 1.0f / sum;
      }
      // ---- epilogue: this warp's 32 rows x 32 of the 64 output columns
      ptx::mbar_wait(o_full, qt & 1);
      ptx::tc_fence_after();
      ptx::tmem_ld_32x32(trow + hh * 32, r);
      ptx::tmem_ld_wait();
      // TMEM is free for the next tile's MMA1 once all 8 warps have their O values in registers
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(t_empty);
      if (q0 + row < S) {
        __nv_bfloat16* o = out + (static_cast<size_t>(b) * S + q0 + row) * D + h * HD + hh * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 pk;
          pk.x = pack2(__uint_as_float(r[8 * k]) * inv_o, __uint_as_float(r[8 * k + 1]) * inv_o);
          pk.y = pack2(__uint_as_float(r[8 * k + 2]) * inv_o, __uint_as_float(r[8 * k + 3]) * inv_o);
          pk.z = pack2(__uint_as_float(r[8 * k + 4]) * inv_o, __uint_as_float(r[8 * k + 5]) * inv_o);
          pk.w = pack2(__uint_as_float(r[8 * k + 6]) * inv_o, __uint_as_float(r[8 * k + 7]) * inv_o);
          *reinterpret_cast<uint4*>(o + 8 * k) = pk;
        }
      }
    }
    if (SAVE_P && issuer) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}


// ---------------------------------------------------------------------------------------------
// Backward, first half, fused: dS = P o (dP - delta) / sqrt(hd) with dP = dO V^T and
// delta_i = sum_j dP_ij P_ij -- replaces the dP GEMM (fp32 [G, S, S] written to HBM) + the row kernel
// that read it back.  Same CTA geometry as the forward: dP of a 128-query tile is accumulated
// straight into TMEM (512 columns), the saved probabilities of the tile arrive by TMA as eight
// [128 x 64] SWIZZLE_128B chunks, the eight softmax warps make two passes (delta, then dS written
// IN PLACE over P in shared memory) and every finished chunk leaves by one TMA store.  The staged
// dS chunks are K-major UMMA operands as they stand, so dQ = dS K is accumulated right here too
// (K reloaded MN-major over the V tile once MMA1 has retired; accumulator = TMEM columns 0..63,
// drained by chunk 0's second pass before the first MMA2) -- the separate dQ GEMM disappears.
constexpr int B_OFF_DO = 0;                           // 16 KB
constexpr int B_OFF_V = B_OFF_DO + BQ * HD * 2;       // 4 x 16 KB (K-major boxes of 128 keys)
constexpr int B_OFF_P = B_OFF_V + SMAX * HD * 2;      // 8 x 16 KB
constexpr int B_OFF_XCH = B_OFF_P + 8 * P_BYTES;      // float [2][128]
constexpr int B_OFF_BARS = B_OFF_XCH + 2 * BQ * 4;
constexpr int B_SMEM_BYTES = B_OFF_BARS + 16 * 8 + 16 + 1024;

// DROP: dP = dPd * mask / (1 - p) with the forward's mask regenerated from the element counter.
template <bool DROP>
__global__ void __launch_bounds__(THREADS, 1)
fattn_bwd_ds_kernel(const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_v,
                    const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_ds,
                    const __grid_constant__ CUtensorMap tm_k, __nv_bfloat16* __restrict__ dqkv, int S, int D,
                    int H, float scale, const dropout::Site drop) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BARS);
  uint64_t* in_full = bars;       // dO + V landed
  uint64_t* p_full = bars + 1;    // P tile landed
  uint64_t* s_full = bars + 2;    // dP complete in TMEM
  uint64_t* k_full = bars + 3;    // K (MN-major) landed over the V tile
  uint64_t* o_full = bars + 4;    // dQ complete in TMEM
  uint64_t* ds_full = bars + 8;   // [8] dS chunk staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* xch = reinterpret_cast<float*>(smem + B_OFF_XCH);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tiles = (S + BQ - 1) / BQ;
  const int g = blockIdx.x / q_tiles, qt = blockIdx.x % q_tiles;
  const int b = g / H, h = g % H;
  const int q0 = qt * BQ;
  const int n_kb = (S + KB - 1) / KB, n_ch = (S + CH - 1) / CH;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_do);
    ptx::prefetch_tmap(&tm_v);
    ptx::prefetch_tmap(&tm_p);
    ptx::prefetch_tmap(&tm_ds);
    ptx::prefetch_tmap(&tm_k);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 16; ++i) ptx::mbar_init(&bars[i], 1);
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

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(in_full, BQ * HD * 2 + n_kb * KB * HD * 2);
      ptx::tma_load_3d(smem + B_OFF_DO, &tm_do, in_full, h * HD, b, q0);
      for (int kb = 0; kb < n_kb; ++kb)
        ptx::tma_load_3d(smem + B_OFF_V + kb * KB * HD * 2, &tm_v, in_full, 2 * D + h * HD, b, kb * KB);
      ptx::mbar_arrive_expect_tx(p_full, n_ch * P_BYTES);
      for (int c = 0; c < n_ch; ++c)
        ptx::tma_load_3d(smem + B_OFF_P + c * P_BYTES, &tm_p, p_full, c * CH, g, q0);
      // MMA1 has retired -> the V tile is dead: K arrives over it, MN-major, for dQ = dS K
      ptx::mbar_wait(s_full, 0);
      ptx::mbar_arrive_expect_tx(k_full, n_ch * CH * HD * 2);
      for (int c = 0; c < n_ch; ++c)
        ptx::tma_load_3d(smem + B_OFF_V + c * CH * HD * 2, &tm_k, k_full, D + h * HD, b, c * CH);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_wait(in_full, 0);
      ptx::tc_fence_after();
      const uint32_t id1 = ptx::make_idesc_bf16(BQ, KB, 0, 0);
      const uint32_t sa = ptx::smem_u32(smem + B_OFF_DO);
      for (int kb = 0; kb < n_kb; ++kb) {
        const uint32_t sb = ptx::smem_u32(smem + B_OFF_V + kb * KB * HD * 2);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::umma_bf16(tmem + kb * KB, ptx::make_smem_desc_sw128(sa + k * 32, 16u, 1024u),
                         ptx::make_smem_desc_sw128(sb + k * 32, 16u, 1024u), id1, k > 0 ? 1u : 0u);
      }
      ptx::umma_commit(s_full);
      // ---- MMA2: dQ[128 x 64] += dS_chunk K_chunk, chunk 0 first (its columns become the accumulator)
      ptx::mbar_wait(k_full, 0);
      const uint32_t id2 = ptx::make_idesc_bf16(BQ, HD, 0, 1);
      bool first = true;
      for (int i = 0; i < 4; ++i) {
        for (int hh = 0; hh < 2; ++hh) {
          const int c = hh * 4 + i;
          if (c >= n_ch) continue;
          ptx::mbar_wait(&ds_full[c], 0);
          ptx::tc_fence_after();
          const uint32_t sp = ptx::smem_u32(smem + B_OFF_P + c * P_BYTES);
          const uint32_t sk = ptx::smem_u32(smem + B_OFF_V + c * CH * HD * 2);
#pragma unroll
          for (int k = 0; k < CH / 16; ++k)
            ptx::umma_bf16(tmem, ptx::make_smem_desc_sw128(sp + k * 32, 16u, 1024u),
                           ptx::make_smem_desc_sw128(sk + k * 2048, 8192u, 1024u), id2,
                           (first && k == 0) ? 0u : 1u);
          first = false;
        }
      }
      ptx::umma_commit(o_full);
    }
  } else {
    const int we = warp - 2;
    const int q = warp & 3;
    const int hh = we >> 2;
    const int row = q * 32 + lane;
    const int c_begin = hh * 4, c_end = min(n_ch, hh * 4 + 4);
    const uint32_t trow = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t prow = static_cast<uint32_t>(row >> 3) * 1024u + static_cast<uint32_t>(row & 7) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t pbase = ptx::smem_u32(smem + B_OFF_P);
    const bool issuer = (q == 2 && lane == 0);
    ptx::mbar_wait(p_full, 0);
    ptx::mbar_wait(s_full, 0);
    ptx::tc_fence_after();
    uint32_t r[32];
    // element counter of (this query row, key 0) in the forward's mask
    const unsigned int e_row = (static_cast<unsigned int>(g) * static_cast<unsigned int>(S) +
                                static_cast<unsigned int>(q0 + row)) * static_cast<unsigned int>(S);
    // dPd -> dP in place (registers); the mask function is evaluated in both passes (keeping one
    // bit per element from pass 1 in eight registers made the kernel 35 % SLOWER: the fully
    // unrolled passes lose the overlap of the TMEM loads with the arithmetic)
    auto undrop = [&](int c, int j) {
      if constexpr (DROP) {
        const unsigned int e0 = e_row + static_cast<unsigned int>(c * CH + j * 32);
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * drop.mult(e0 + i));
      }
    };
    // pass 1: delta = sum_j dP_ij P_ij
    float delta = 0.f;
    for (int c = c_begin; c < c_end; ++c) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        ptx::tmem_ld_32x32(trow + c * CH + j * 32, r);
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
        uint4 pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          pk[k] = ptx::lds_v4u(pbase + c * P_BYTES + prow + ((static_cast<uint32_t>(j * 4 + k) ^ sw) << 4));
        ptx::tmem_ld_wait();
        undrop(c, j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t w[4] = {pk[k].x, pk[k].y, pk[k].z, pk[k].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            d4[t] = fmaf(__uint_as_float(r[8 * k + 2 * t]), __uint_as_float(w[t] << 16), d4[t]);
            d4[t] = fmaf(__uint_as_float(r[8 * k + 2 * t + 1]), __uint_as_float(w[t] & 0xffff0000u), d4[t]);
          }
        }
        delta += (d4[0] + d4[1]) + (d4[2] + d4[3]);
      }
    }
    xch[hh * BQ + row] = delta;
    nbar(3 + q, 64);
    delta += xch[(hh ^ 1) * BQ + row];
    // pass 2: dS = P (dP - delta) * scale, in place over P, chunk by chunk -> TMA store
    for (int c = c_begin; c < c_end; ++c) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        ptx::tmem_ld_32x32(trow + c * CH + j * 32, r);
        uint4 pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          pk[k] = ptx::lds_v4u(pbase + c * P_BYTES + prow + ((static_cast<uint32_t>(j * 4 + k) ^ sw) << 4));
        ptx::tmem_ld_wait();
        undrop(c, j);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t w[4] = {pk[k].x, pk[k].y, pk[k].z, pk[k].w};
          uint32_t o[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float lo = __uint_as_float(w[t] << 16) * (__uint_as_float(r[8 * k + 2 * t]) - delta) * scale;
            const float hi = __uint_as_float(w[t] & 0xffff0000u) * (__uint_as_float(r[8 * k + 2 * t + 1]) - delta) * scale;
            o[t] = pack2(lo, hi);
          }
          ptx::sts_v4u(pbase + c * P_BYTES + prow + ((static_cast<uint32_t>(j * 4 + k) ^ sw) << 4), o[0], o[1],
                       o[2], o[3]);
        }
      }
      if (c == 0) ptx::tc_fence_before();  // columns 0..63 are drained before MMA2 overwrites them
      ptx::fence_proxy_async();
      nbar(1 + hh, 128);
      if (issuer) {
        ptx::mbar_arrive(&ds_full[c]);
        ptx::tma_store_3d(&tm_ds, pbase + c * P_BYTES, c * CH, g, q0);
        ptx::bulk_commit();
      }
    }
    // ---- dQ tile: this warp's 32 rows x 32 of the 64 columns -> dqkv[b, s, h*64 ...] (the q third)
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    ptx::tmem_ld_32x32(trow + hh * 32, r);
    ptx::tmem_ld_wait();
    if (q0 + row < S) {
      __nv_bfloat16* o = dqkv + (static_cast<size_t>(b) * S + q0 + row) * (3 * static_cast<size_t>(D)) + h * HD + hh * 32;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 pk4;
        pk4.x = pack2(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1]));
        pk4.y = pack2(__uint_as_float(r[8 * k + 2]), __uint_as_float(r[8 * k + 3]));
        pk4.z = pack2(__uint_as_float(r[8 * k + 4]), __uint_as_float(r[8 * k + 5]));
        pk4.w = pack2(__uint_as_float(r[8 * k + 6]), __uint_as_float(r[8 * k + 7]));
        *reinterpret_cast<uint4*>(o + 8 * k) = pk4;
      }
    }
    if (issuer) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace fattn

// Returns 1 when the fused kernel does not apply (caller falls back to the three-kernel path).
int fused_seq_attention_fwd(const void* qkv, const float* addmask, void* out, void* probs, int B, int S,
                            int D, int H, cudaStream_t stream, dropout::Site drop, void* pdrop) {
  using namespace fattn;
  static const bool disabled = getenv("MMU_ATTN_UNFUSED") != nullptr;  // A/B switch
  if (disabled || D / H != HD || D % H != 0 || S > SMAX || S < 1) return 1;
  if (drop.on() && probs == nullptr) return 1;   // dropout: training only
  const int Sp = (S + 7) / 8 * 8;
  CUtensorMap tq, tv, tp, tpd;
  int rc = make_tmap_bf16_3d(&tq, qkv, 3LL * D, B, S, 3LL * D * S, 3LL * D, HD, KB);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tv, qkv, 3LL * D, B, S, 3LL * D * S, 3LL * D, HD, CH);
  if (rc) return rc;
  tp = tq;
  if (probs != nullptr) {
    rc = make_tmap_bf16_3d(&tp, probs, Sp, static_cast<long long>(B) * H, S, static_cast<long long>(S) * Sp,
                           Sp, CH, BQ);
    if (rc) return rc;
  }
  tpd = tp;
  const int save_pd = (drop.on() && pdrop != nullptr) ? 1 : 0;
  if (save_pd) {
    rc = make_tmap_bf16_3d(&tpd, pdrop, Sp, static_cast<long long>(B) * H, S, static_cast<long long>(S) * Sp,
                           Sp, CH, BQ);
    if (rc) return rc;
  }
  // query tiles per CTA (K / V staged once per CTA).  Measured on B200 at B*H = 384, S = 512
  // (MMBT train step / eval forward, ms): 1 tile per CTA 15.6 / 4.45, 2 tiles 15.5 / 4.30, all 4
  // tiles (a CTA per head) 15.3 / 4.05 -- staging K/V once beats the 2.6-wave grid's quantisation
  static const int tpc_env_eval = getenv("MMU_FATTN_TPC_EVAL") ? atoi(getenv("MMU_FATTN_TPC_EVAL")) : 0;
  static const int tpc_env_train = getenv("MMU_FATTN_TPC_TRAIN") ? atoi(getenv("MMU_FATTN_TPC_TRAIN")) : 0;
  const int all_tiles = (S + BQ - 1) / BQ;
  int tpc = probs != nullptr ? (tpc_env_train > 0 ? tpc_env_train : all_tiles)
                             : (tpc_env_eval > 0 ? tpc_env_eval : all_tiles);
  if (tpc > all_tiles) tpc = all_tiles;
  const int grid = B * H * ((all_tiles + tpc - 1) / tpc);
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  auto launch = [&](auto kernel) -> int {
    static cudaError_t attr = cudaSuccess;
    attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (attr != cudaSuccess) return MMU_ERR_CUDA;
    kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tq, tv, tp, tpd, addmask, static_cast<__nv_bfloat16*>(out),
                                                  B, S, D, H, scale, tpc, drop, save_pd);
    if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
    count_launch();
    return 0;
  };
  if (drop.on()) return launch(fattn_fwd_kernel<true, true>);
  return probs != nullptr ? launch(fattn_fwd_kernel<true, false>) : launch(fattn_fwd_kernel<false, false>);
}

// dprobs (bf16 [G, S, Sp]) = dS and the q third of dqkv = dS K.  Returns 1 when the fused kernel
// does not apply.
int fused_seq_attention_bwd_ds(const void* qkv, const void* dout, const void* probs, void* dprobs, void* dqkv,
                               int B, int S, int D, int H, cudaStream_t stream, dropout::Site drop) {
  using namespace fattn;
  static const bool disabled = getenv("MMU_ATTN_UNFUSED") != nullptr;
  if (disabled || D / H != HD || D % H != 0 || S > SMAX || S < 1) return 1;
  const int Sp = (S + 7) / 8 * 8;
  const long long G = static_cast<long long>(B) * H;
  CUtensorMap tdo, tv, tp, tds, tk;
  int rc = make_tmap_bf16_3d(&tdo, dout, D, B, S, static_cast<long long>(D) * S, D, HD, BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tv, qkv, 3LL * D, B, S, 3LL * D * S, 3LL * D, HD, KB);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tp, probs, Sp, G, S, static_cast<long long>(S) * Sp, Sp, CH, BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tds, dprobs, Sp, G, S, static_cast<long long>(S) * Sp, Sp, CH, BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tk, qkv, 3LL * D, B, S, 3LL * D * S, 3LL * D, HD, CH);
  if (rc) return rc;
  static cudaError_t attr[2] = {
      cudaFuncSetAttribute(fattn_bwd_ds_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_BYTES),
      cudaFuncSetAttribute(fattn_bwd_ds_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM_BYTES)};
  if (attr[0] != cudaSuccess || attr[1] != cudaSuccess) return MMU_ERR_CUDA;
  const int grid = static_cast<int>(G) * ((S + BQ - 1) / BQ);
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  if (drop.on())
    fattn_bwd_ds_kernel<true><<<grid, THREADS, B_SMEM_BYTES, stream>>>(
        tdo, tv, tp, tds, tk, static_cast<__nv_bfloat16*>(dqkv), S, D, H, scale, drop);
  else
    fattn_bwd_ds_kernel<false><<<grid, THREADS, B_SMEM_BYTES, stream>>>(
        tdo, tv, tp, tds, tk, static_cast<__nv_bfloat16*>(dqkv), S, D, H, scale, drop);
  if (cudaGetLastError() != cudaSuccess) return MMU_ERR_CUDA;
  count_launch();
  return 0;
}

}  // namespace mmu
