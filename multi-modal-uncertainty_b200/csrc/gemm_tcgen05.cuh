// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring (4 stages of 48 KB)
//   -> tcgen05.mma (one elected thread, 128x256x16 UMMA, fp32 accumulators in TMEM,
//      two accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1)
//   -> tcgen05.ld -> fused epilogue in registers (bias / QuickGELU / dGELU) -> swizzled shared
//      staging boxes -> TMA tensor stores (cp.async.bulk.tensor store, or cp.reduce .add for the
//      split-K weight gradients).  The dGELU operand z is TMA-loaded into the same boxes.
//
// Computes  C[M,N] = epilogue( sum_k A(m,k) * B(n,k) ).
// Either operand may be K-major (stored [rows][K], K contiguous) or MN-major (stored
// [K][rows], rows contiguous): the latter is what the weight-gradient GEMMs need
// (dW = dY^T X reduces over the row axis of both operands) and what dgrad needs for W.
//
// This is the tensor-core path for every dense contraction of the fusion model
// (reference call sites: src/model.py:262,264 projections; :193 MHA in/out proj;
// :195-201 MLP c_fc/c_proj) and their backward passes.
//
// Epilogue geometry: 8 warps; warp w owns TMEM lanes 32*(w%4).. (hardware rule) and a 128-column
// half of the 256-column accumulator, i.e. a 32-row x 128-column slab, processed in chunks of 32
// (bf16 out) or 16 (fp32 out) columns with tcgen05.ld (lane = row, registers = consecutive
// columns).  A chunk is written as four 16-byte pieces per lane into a [32 rows x 64 B] staging
// box laid out exactly as TMA's SWIZZLE_64B expects (piece index XOR ((row >> 1) & 3)):
// conflict-free for the writers, and one TMA store per box then emits full 64-byte row segments
// -- no per-thread global stores and no edge predicates (the tensor map clips rows >= M and
// columns >= N).  Each warp alternates between two boxes, so writing chunk c overlaps the drain
// of chunk c-1; 4 operand stages (192 KB) + 32 KB of boxes fill the SM's shared memory.
#pragma once
#include "gemm_api.h"
#include "ptx.cuh"

namespace mmu {

namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
// NCTA = 1: one CTA per 128x256 tile, 4 stages of (A 16 KB + B 32 KB).
// NCTA = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256x256 tile: each CTA stages its
//           own 128 rows of A and HALF of B (128 of the 256 n-rows), 6 stages of (16 + 16) KB.
//           Operand bytes through shared memory per FLOP drop by a third -- the single-CTA
//           kernel is shared-memory-bandwidth bound (TMA writes + UMMA reads = 192 B/clk of the
//           SM's 128 B/clk at full tensor rate; ncu: tensor pipe 66 % active, all barriers idle).
// EPI_RESID_LN streams the fp32 residual tile through the epilogue (TMA load -> add in place -> TMA
// store, plus a bf16 copy): its warps own FOUR staging boxes (three fp32 boxes in a ring + one bf16
// box) instead of two, paid for with one operand stage.
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int TMEM_COLS = 2 * BN;                       // two accumulator stages (512 = all of TMEM)
constexpr int BOX_ROWS = 32;
constexpr int BOX_BYTES = BOX_ROWS * 64;                // 32 rows x 64 B (32 bf16 / 16 fp32 columns)
constexpr int MAX_STAGES = 6;
constexpr int LN_MAX_NT = 8;                            // partials per row of a folded LayerNorm
constexpr int AUX_BARS = 3;                             // per epilogue warp (operand boxes in flight)
constexpr int NUM_BARS = 2 * MAX_STAGES + 4 + AUX_BARS * NUM_EPI_WARPS;
// FOLD (LayerNorm folded into the epilogue, gemm_api.h) also gives up one operand stage: an extra
// warp stages the tile's folded-weight column sums and its rows' (-mean, rstd) in shared memory one
// tile ahead of the epilogue warps.
template <int NCTA, int MODE, bool FOLD = false> struct Geo {
  static constexpr bool RESID = MODE == EPI_RESID_LN;
  static constexpr int STAGES = NCTA == 2 ? ((RESID || FOLD) ? 5 : 6) : ((RESID || FOLD) ? 3 : 4);
  static constexpr int A_STAGE_BYTES = BM * BK * 2;            // 16 KiB
  static constexpr int B_ROWS = BN / NCTA;                     // n-rows of B staged by this CTA
  static constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;        // 32 / 16 KiB
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BM_TILE = BM * NCTA;                    // output rows per tile
  static constexpr int OPERAND_BYTES = STAGES * STAGE_BYTES;   // 192 KiB (RESID: 160 / 144 KiB)
  static constexpr int BOXES = RESID ? 4 : 2;                  // staging boxes per epilogue warp
  static constexpr int STG_WARP_BYTES = BOXES * BOX_BYTES;
  static constexpr int OFF_STG = OPERAND_BYTES;                // 1024-aligned
  static constexpr int OFF_BIAS = OFF_STG + NUM_EPI_WARPS * STG_WARP_BYTES;
  static constexpr int OFF_CW = OFF_BIAS + 2 * BN * 4;         // bias tile, double buffered
  static constexpr int OFF_ROWST = OFF_CW + (FOLD ? 2 * BN * 4 : 0);     // FOLD: column sums [2][BN]
  static constexpr int OFF_BARS = OFF_ROWST + (FOLD ? 2 * BM * 8 : 0);   // FOLD: (-mean, rstd) [2][BM]
  static constexpr int SMEM_USED = OFF_BARS + NUM_BARS * 8 + 16;  // + barriers and the TMEM slot
  static_assert(OFF_STG % 1024 == 0, "staging boxes must stay aligned for the swizzle pattern");
};
constexpr int SMEM_BYTES = 227 * 1024;                  // everything an SM has; the kernel checks
                                                        // that SMEM_USED fits behind the 1024-byte
                                                        // alignment of the dynamic window
static_assert(Geo<1, EPI_STORE>::OPERAND_BYTES == 192 * 1024 && Geo<2, EPI_STORE>::OPERAND_BYTES == 192 * 1024,
              "operand ring size");
static_assert(Geo<1, EPI_STORE>::SMEM_USED <= SMEM_BYTES && Geo<2, EPI_RESID_LN>::SMEM_USED <= SMEM_BYTES &&
              Geo<1, EPI_RESID_LN>::SMEM_USED <= SMEM_BYTES && Geo<1, EPI_STORE, true>::SMEM_USED <= SMEM_BYTES &&
              Geo<2, EPI_STORE, true>::SMEM_USED <= SMEM_BYTES, "shared memory budget exceeded");

// sigmoid(1.702 z) = 0.5 tanh(0.851 z) + 0.5: ONE MUFU op (tanh.approx, rel. error ~2^-11, far
// below bf16 resolution) instead of ex2 + rcp -- the GELU epilogues are MUFU-bound otherwise.
__device__ __forceinline__ float sigmoid_1702(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * z));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float quick_gelu(float z) { return z * sigmoid_1702(z); }
__device__ __forceinline__ float quick_gelu_grad(float z) {
  const float s = sigmoid_1702(z);
  return s * fmaf(1.702f * z, 1.0f - s, 1.0f);
}

// erf-GELU z*Phi(z) and its derivative Phi(z) + z*phi(z).  Phi through Abramowitz-Stegun 7.1.25
// (erf(x) = 1 - (a1 t + a2 t^2 + a3 t^3) exp(-x^2), t = 1/(1 + 0.47047 x), |error| <= 2.5e-5 --
// far below bf16 resolution): one rcp + one ex2 + ~10 FMAs, where erff() would make the epilogue
// of a K = 768 tile longer than its main loop.  exp(-x^2) = exp(-z^2/2) is shared with phi(z).
__device__ __forceinline__ void gelu_erf_terms(float z, float& cdf, float& pdf) {
  const float x = fabsf(z) * 0.70710678118654752f;
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.47047f, x, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * x * x));
  const float poly = t * fmaf(t, fmaf(t, 0.7478556f, -0.0958798f), 0.3480242f);
  const float half_erfc = 0.5f * poly * ex;             // 0.5 * erfc(|z| / sqrt 2)
  cdf = z >= 0.f ? 1.0f - half_erfc : half_erfc;
  pdf = 0.3989422804014327f * ex;
}
__device__ __forceinline__ float erf_gelu(float z) {
  float cdf, pdf;
  gelu_erf_terms(z, cdf, pdf);
  return z * cdf;
}
__device__ __forceinline__ float erf_gelu_grad(float z) {
  float cdf, pdf;
  gelu_erf_terms(z, cdf, pdf);
  return fmaf(z, pdf, cdf);
}

// Packed fp32 pairs: sm_100a issues FFMA2 / FMUL2 on a 64-bit register pair -- the IEEE result of
// the scalar op in each half for ONE issue slot.  The epilogue warps are instruction-issue limited
// (removing ~8 FP32 instructions per element from the dGELU epilogue made its launch 13 % shorter,
// profiles/r02_harness_act_grad_v1.log), so the per-element arithmetic below runs on pairs of
// adjacent columns; same operations in the same order as the scalar helpers above.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t bc2(float x) { return pk2(x, x); }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sigmoid_1702_2(uint64_t z) {
  float a0, a1, t0, t1;
  up2(mul2(z, bc2(0.851f)), a0, a1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  return fma2(pk2(t0, t1), bc2(0.5f), bc2(0.5f));
}
// (z0, z1) <- QuickGELU of the pair
__device__ __forceinline__ void quick_gelu_pair(float& z0, float& z1) {
  const uint64_t z = pk2(z0, z1);
  up2(mul2(z, sigmoid_1702_2(z)), z0, z1);
}
// (v0, v1) <- (v0, v1) * QuickGELU'(z0, z1)
__device__ __forceinline__ void quick_gelu_grad_pair(float& v0, float& v1, float z0, float z1) {
  const uint64_t z = pk2(z0, z1);
  const uint64_t sg = sigmoid_1702_2(z);
  const uint64_t oms = fma2(sg, bc2(-1.0f), bc2(1.0f));                  // 1 - s, one rounding
  const uint64_t g = mul2(sg, fma2(mul2(z, bc2(1.702f)), oms, bc2(1.0f)));
  up2(mul2(pk2(v0, v1), g), v0, v1);
}
// erf-GELU terms of a pair: gelu_erf_terms() operation by operation (|z| k == |z k| and
// (-c x) x is sign-symmetric, so working on the signed x = z / sqrt 2 changes no bit).
__device__ __forceinline__ void gelu_erf_terms2(uint64_t z, uint64_t& cdf, uint64_t& pdf) {
  const uint64_t xs = mul2(z, bc2(0.70710678118654752f));
  float x0, x1, d0, d1, t0, t1, m0, m1, e0, e1;
  up2(xs, x0, x1);
  up2(fma2(bc2(0.47047f), pk2(fabsf(x0), fabsf(x1)), bc2(1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  up2(mul2(mul2(xs, bc2(-1.4426950408889634f)), xs), m0, m1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(m0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(m1));
  const uint64_t t = pk2(t0, t1), ex = pk2(e0, e1);
  const uint64_t poly = mul2(t, fma2(t, fma2(t, bc2(0.7478556f), bc2(-0.0958798f)), bc2(0.3480242f)));
  const uint64_t he = mul2(mul2(bc2(0.5f), poly), ex);     // 0.5 * erfc(|z| / sqrt 2)
  float h0, h1, o0, o1, z0, z1;
  up2(he, h0, h1);
  up2(fma2(he, bc2(-1.0f), bc2(1.0f)), o0, o1);            // 1 - half_erfc, one rounding
  up2(z, z0, z1);
  cdf = pk2(z0 >= 0.f ? o0 : h0, z1 >= 0.f ? o1 : h1);
  pdf = mul2(bc2(0.3989422804014327f), ex);
}
__device__ __forceinline__ void erf_gelu_pair(float& z0, float& z1) {
  const uint64_t z = pk2(z0, z1);
  uint64_t cdf, pdf;
  gelu_erf_terms2(z, cdf, pdf);
  up2(mul2(z, cdf), z0, z1);
}
__device__ __forceinline__ void erf_gelu_grad_pair(float& v0, float& v1, float z0, float z1) {
  const uint64_t z = pk2(z0, z1);
  uint64_t cdf, pdf;
  gelu_erf_terms2(z, cdf, pdf);
  up2(mul2(pk2(v0, v1), fma2(z, pdf, cdf)), v0, v1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// MODE: GemmEpiMode; OBF: outputs (and the dGELU operand) are bf16, else fp32; NCTA: CTAs per
// tile (2 = cta_group::2 pair, launched as a cluster of 2; not used for batched problems).
// tma_c0: `out`; tma_c1: `out2` (QUICKGELU, RESID_LN) or `aux` (DGELU); tma_c2: the fp32 `aux` of
// RESID_LN.  All are 4-D maps (columns, rows, batch % out_hdiv, batch / out_hdiv) with a
// [32 x 64 B] box.
template <int MODE, bool OBF, int NCTA, bool FOLD = false>
__global__ void __launch_bounds__(NUM_THREADS + (FOLD ? 32 : 0), 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_c0,
                         const __grid_constant__ CUtensorMap tma_c1,
                         const __grid_constant__ CUtensorMap tma_c2, const GemmProblem p,
                         const GemmEpilogue e) {
  static_assert(!FOLD || ((MODE == EPI_STORE || MODE == EPI_QUICKGELU) && OBF), "FOLD: bf16 STORE / QUICKGELU");
  using G = Geo<NCTA, MODE, FOLD>;
  constexpr int OFF_STG = G::OFF_STG, OFF_BIAS = G::OFF_BIAS, OFF_BARS = G::OFF_BARS;
  constexpr int SMEM_USED = G::SMEM_USED, STG_WARP_BYTES = G::STG_WARP_BYTES;
  constexpr int STAGES = G::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles and staging boxes need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * G::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* full_bar = bars;                         // [STAGES]  TMA -> MMA (pair: the leader's)
  uint64_t* empty_bar = bars + MAX_STAGES;           // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;       // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;  // [2]       epilogue -> MMA (pair: the leader's)
  uint64_t* aux_bar = bars + 2 * MAX_STAGES + 4;     // [NUM_EPI_WARPS][AUX_BARS] operand box landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
  if (threadIdx.x == 0 && (smem - smem_raw) + SMEM_USED > SMEM_BYTES) {
    printf("mmu: dynamic shared memory window is not 1024-byte aligned (offset %d)\n",
           static_cast<int>(smem - smem_raw));
    __trap();
  }

  ptx::pdl_trigger();  // the next kernel may start its prologue on SMs this grid has left
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;  // CTA within the pair
  const int tile0 = NCTA == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = NCTA == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  const int m_tiles = (p.M + G::BM_TILE - 1) / G::BM_TILE;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  const int nbatch = p.batch > 0 ? p.batch : 1;
  const int mn_total = nbatch * m_tiles * n_tiles;
  const int num_tiles = mn_total * p.splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
    ptx::prefetch_tmap(&tma_c0);
    if (MODE == EPI_QUICKGELU || MODE == EPI_DGELU || MODE == EPI_RESID_LN) ptx::prefetch_tmap(&tma_c1);
    if (MODE == EPI_RESID_LN) ptx::prefetch_tmap(&tma_c2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], NCTA * NUM_EPI_WARPS);  // pair: both CTAs' epilogue warps
    }
    for (int i = 0; i < AUX_BARS * NUM_EPI_WARPS; ++i) ptx::mbar_init(&aux_bar[i], 1);
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    if constexpr (NCTA == 2) {
      ptx::tmem_alloc_pair(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (NCTA == 2) ptx::cluster_sync_all();  // the peer's barriers are initialised too
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the tail of
  // the previous kernel; its results are needed from here on.
  ptx::pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int split = tile / mn_total;  // K-split outermost: concurrent CTAs share operand panels
        const int mnb = tile % mn_total;
        const int g = mnb / (m_tiles * n_tiles);
        const int mn = mnb % (m_tiles * n_tiles);
        const int m0 = (mn / n_tiles) * G::BM_TILE + rank * BM;   // this CTA's 128 rows of A
        const int n0 = (mn % n_tiles) * BN + rank * G::B_ROWS;    // pair: this CTA's half of B
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        const int a_mid = p.batch > 0 ? g / p.a_hdiv : 0;
        const int a_off = p.batch > 0 ? (g % p.a_hdiv) * p.a_hstride + p.a_col0 : 0;
        const int b_mid = p.batch > 0 ? g / p.b_hdiv : 0;
        const int b_off = p.batch > 0 ? (g % p.b_hdiv) * p.b_hstride + p.b_col0 : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * G::A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * G::B_STAGE_BYTES;
          const int k0 = kb * BK;
          if constexpr (NCTA == 2) {
            // both CTAs' bytes are accounted on the LEADER's barrier (the MMA issuer waits there)
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * G::STAGE_BYTES);
            if (!p.a_mn_major) {
              ptx::tma_load_2d_pair(sa, &tma_a, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                ptx::tma_load_2d_pair(sa + j * 8192, &tma_a, &full_bar[stage], m0 + 64 * j, k0);
            }
            if (!p.b_mn_major) {
              ptx::tma_load_2d_pair(sb, &tma_b, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < G::B_ROWS / 64; ++j)
                ptx::tma_load_2d_pair(sb + j * 8192, &tma_b, &full_bar[stage], n0 + 64 * j, k0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          // N <= 128 problems stage (and multiply) only 128 rows of B
          const int b_boxes = p.half_n ? BN / 128 : BN / 64;
          ptx::mbar_arrive_expect_tx(&full_bar[stage],
                                     G::A_STAGE_BYTES + (p.half_n ? G::B_STAGE_BYTES / 2 : G::B_STAGE_BYTES));
          if (p.batch > 0) {
            if (!p.a_mn_major) {
              ptx::tma_load_3d(sa, &tma_a, &full_bar[stage], a_off + k0, a_mid, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                ptx::tma_load_3d(sa + j * 8192, &tma_a, &full_bar[stage], a_off + m0 + 64 * j, a_mid, k0);
            }
            if (!p.b_mn_major) {
              ptx::tma_load_3d(sb, &tma_b, &full_bar[stage], b_off + k0, b_mid, n0);
            } else {
              for (int j = 0; j < b_boxes; ++j)
                ptx::tma_load_3d(sb + j * 8192, &tma_b, &full_bar[stage], b_off + n0 + 64 * j, b_mid, k0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (!p.a_mn_major) {
            ptx::tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              ptx::tma_load_2d(sa + j * 8192, &tma_a, &full_bar[stage], m0 + 64 * j, k0);
          }
          if (!p.b_mn_major) {
            ptx::tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
          } else {
            for (int j = 0; j < b_boxes; ++j)
              ptx::tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n0 + 64 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0 && rank == 0) {  // pair: the leader issues for both SMs
      const uint32_t idesc = ptx::make_idesc_bf16(G::BM_TILE, (NCTA == 1 && p.half_n) ? BN / 2 : BN,
                                                  p.a_mn_major, p.b_mn_major);
      // K-major SW128: 8-row groups 1024 B apart, k advances 32 B inside the swizzle atom.
      // MN-major SW128: 64-element MN groups 8192 B apart (one TMA box), 8-k groups 1024 B apart.
      const uint32_t a_lbo = p.a_mn_major ? 8192u : 16u, a_kstep = p.a_mn_major ? 2048u : 32u;
      const uint32_t b_lbo = p.b_mn_major ? 8192u : 16u, b_kstep = p.b_mn_major ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int split = tile / mn_total;  // K-split outermost: concurrent CTAs share operand panels
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_a + stage * G::A_STAGE_BYTES);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * G::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024u);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024u);
            if constexpr (NCTA == 2) ptx::umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else ptx::umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if constexpr (NCTA == 2) ptx::umma_commit_pair(&empty_bar[stage]);
          else ptx::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (NCTA == 2) ptx::umma_commit_pair(&tfull_bar[as]);
        else ptx::umma_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (FOLD && warp == 2 + NUM_EPI_WARPS) {
    // ------------------------------------------- LN fold: statistics / column-sum staging warp
    // Runs one tile ahead of the epilogue warps and joins their per-tile barrier: buffer `as` of
    // tile i is written before barrier i; the loads of tile i + 1 overlap the epilogue of tile i.
    float* cw_s = reinterpret_cast<float*>(smem + G::OFF_CW);
    float2* rowst_s = reinterpret_cast<float2*>(smem + G::OFF_ROWST);
    int as = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int mn = (tile % mn_total) % (m_tiles * n_tiles);
      const int nt0 = (mn % n_tiles) * BN;
      const int mrow0 = (mn / n_tiles) * G::BM_TILE + rank * BM;
      float2 st[BM / 32];
#pragma unroll
      for (int rr = 0; rr < BM / 32; ++rr) {
        const int row = min(mrow0 + rr * 32 + lane, p.M - 1);
        const float2* part = reinterpret_cast<const float2*>(e.ln_stats) + static_cast<long long>(row) * e.ln_nt;
        float s1 = 0.f, s2 = 0.f;
        for (int j = 0; j < e.ln_nt; ++j) {
          const float2 t = __ldg(part + j);
          s1 += t.x;
          s2 += t.y;
        }
        const float mean = s1 * e.ln_inv_d;
        st[rr] = make_float2(-mean, rsqrtf(fmaxf(fmaf(s2, e.ln_inv_d, -mean * mean), 0.f) + e.ln_eps));
      }
      float cwv[BN / 32];
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) {
        const int col = nt0 + j * 32 + lane;
        cwv[j] = col < p.N ? __ldg(e.ln_cw + col) : 0.f;
      }
#pragma unroll
      for (int rr = 0; rr < BM / 32; ++rr) rowst_s[as * BM + rr * 32 + lane] = st[rr];
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) cw_s[as * BN + j * 32 + lane] = cwv[j];
      named_bar_sync(1, NUM_EPI_THREADS + 32);
      as ^= 1;
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int we = warp - 2;         // 0..7
    const int q = warp & 3;          // TMEM lane quadrant (hardware: lanes 32*(warp%4)..)
    const int h = we >> 2;           // which 128-column half of the accumulator
    const int et = threadIdx.x - 64; // 0..255
    const uint32_t box0 = ptx::smem_u32(smem + OFF_STG + we * STG_WARP_BYTES);
    const uint32_t row_off = static_cast<uint32_t>(lane) * 64u;   // this lane's row in a box
    const uint32_t sw = static_cast<uint32_t>(lane >> 1) & 3u;    // SWIZZLE_64B piece XOR
    auto piece = [&](int k) { return row_off + ((static_cast<uint32_t>(k) ^ sw) << 4); };
    float* bias_s = reinterpret_cast<float*>(smem + OFF_BIAS);
    const uint32_t bias_addr = ptx::smem_u32(bias_s);
    const bool has_bias = e.bias != nullptr;
    constexpr int NCOL = OBF ? 32 : 16;         // columns per chunk = per staging box
    constexpr int NCHUNK = BN / 2 / NCOL;
    uint64_t* my_aux = aux_bar + AUX_BARS * we;
    int as = 0;
    uint32_t aphase = 0, aux_phase[AUX_BARS] = {0, 0, 0};
    constexpr bool fold = FOLD;  // LayerNorm folded into this GEMM (gemm_api.h)
    const uint32_t cw_addr = ptx::smem_u32(smem + G::OFF_CW);
    const float2* rowst_s = reinterpret_cast<const float2*>(smem + G::OFF_ROWST);
    auto release_tmem = [&](int stage_idx) {  // one arrival per epilogue warp per tile
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (NCTA == 2) ptx::mbar_arrive_leader(&tempty_bar[stage_idx]);
        else ptx::mbar_arrive(&tempty_bar[stage_idx]);
      }
    };
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int mnb = tile % mn_total;
      const int g = mnb / (m_tiles * n_tiles);
      const int mn = mnb % (m_tiles * n_tiles);
      const int nt0 = (mn % n_tiles) * BN;
      const int m0 = (mn / n_tiles) * G::BM_TILE + rank * BM + q * 32;
      const int n0 = nt0 + h * (BN / 2);
      const int c2 = p.batch > 0 ? g % p.out_hdiv : 0;
      const int c3 = p.batch > 0 ? g / p.out_hdiv : 0;
      const bool active = m0 < p.M && n0 < p.N;  // warp-uniform

      // the tile's bias slice, once, into shared memory (double buffered across tiles)
      if (has_bias) {
        const int col = nt0 + et;
        bias_s[as * BN + et] = col < p.N ? e.bias[col] : 0.f;
      }
      named_bar_sync(1, NUM_EPI_THREADS + (FOLD ? 32 : 0));

      // LN fold: this lane's row statistics (lane = row), staged by the statistics warp
      float ln_nmean = 0.f, ln_rstd = 1.f;
      if constexpr (fold) {
        const float2 t = rowst_s[as * BM + q * 32 + lane];
        ln_nmean = t.x;
        ln_rstd = t.y;
      }

      // dGELU: the z boxes of the first two chunks start loading while the MMAs of the tile run
      auto load_z = [&](int c) {  // lane 0 only; box (c & 1) must be drained
        ptx::mbar_arrive_expect_tx(&my_aux[c & 1], BOX_BYTES);
        ptx::tma_load_4d(box0 + (c & 1) * BOX_BYTES, &tma_c1, &my_aux[c & 1], n0 + c * NCOL, m0, c2, c3);
      };
      if (MODE == EPI_DGELU && active && lane == 0) {
        ptx::bulk_wait_read<0>();  // the previous tile's stores have drained both boxes
        load_z(0);
        if (n0 + NCOL < p.N) load_z(1);
      }
      // RESID_LN: the fp32 residual boxes of the first two chunks (ring of three boxes: chunk c
      // lives in box c % 3, is updated in place and stored from there; box 3 takes the bf16 copy)
      auto load_x = [&](int c) {  // lane 0 only; box c % 3 must be drained
        ptx::mbar_arrive_expect_tx(&my_aux[c % 3], BOX_BYTES);
        ptx::tma_load_4d(box0 + (c % 3) * BOX_BYTES, &tma_c2, &my_aux[c % 3], n0 + c * NCOL, m0, c2, c3);
      };
      if (MODE == EPI_RESID_LN && active && lane == 0) {
        ptx::bulk_wait_read<0>();
        load_x(0);
        if (n0 + NCOL < p.N) load_x(1);
      }
      float st_s1 = 0.f, st_s2 = 0.f;  // RESID_LN: this lane's row sums over the warp's 128 columns

      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * BN + h * (BN / 2));
      uint32_t r[2][NCOL];
      auto tmem_load = [&](int c) {
        if constexpr (OBF) ptx::tmem_ld_32x32(taddr + c * NCOL, r[c & 1]);
        else ptx::tmem_ld_32x16(taddr + c * NCOL, r[c & 1]);
      };
      // box `b` may be rewritten once the store issued two boxes ago has read it
      auto box_free = [&](bool first) {
        if (lane == 0) {
          if (first) ptx::bulk_wait_read<0>();
          else ptx::bulk_wait_read<1>();
        }
        __syncwarp();
      };
      auto box_store = [&](const CUtensorMap* map, uint32_t box, int col) {
        ptx::fence_proxy_async();  // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          if constexpr (MODE == EPI_ATOMIC) ptx::tma_reduce_add_4d(map, box, col, m0, c2, c3);
          else ptx::tma_store_4d(map, box, col, m0, c2, c3);
          ptx::bulk_commit();
        }
      };
      if constexpr (MODE == EPI_SOFTMAX) {
        // Row softmax of alpha*acc over columns [0, n_valid), N <= 128.  The two warps that share
        // a TMEM lane quadrant split a row's columns (64 each), run an online max/sum pass over
        // TMEM, exchange (max, sum) through shared memory, then normalise and store their half:
        // two TMEM passes, nothing touches HBM but the bf16 probabilities.
        const int cbase = h * 64;
        const bool rows_ok = m0 < p.M;
        const int nloc = min(64, max(0, p.N - cbase));
        const int nch = rows_ok ? (nloc + NCOL - 1) / NCOL : 0;   // 0, 1 or 2 chunks
        const float sc = e.alpha * 1.4426950408889634f;
        const uint32_t tsm = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * BN + cbase);
        auto ex2 = [](float x) {
          float y;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          return y;
        };
        auto sm_load = [&](int c) { ptx::tmem_ld_32x32(tsm + c * NCOL, r[c & 1]); };
        float mx = -INFINITY, sum = 0.f;
        if (nch > 0) sm_load(0);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < nch) {
            ptx::tmem_ld_wait();
            if (c + 1 < nch) sm_load(c + 1);
            const int col0 = cbase + c * NCOL;
            const bool full = col0 + NCOL <= e.n_valid;
            float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < NCOL; ++i)
              if (full || col0 + i < e.n_valid) cm[i & 3] = fmaxf(cm[i & 3], __uint_as_float(r[c & 1][i]));
            const float mn = fmaxf(fmaxf(mx, fmaxf(cm[0], cm[1])), fmaxf(cm[2], cm[3]));
            const float off = -mn * sc;
            float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < NCOL; ++i)
              if (full || col0 + i < e.n_valid)
                part[i & 3] += ex2(fmaf(__uint_as_float(r[c & 1][i]), sc, off));
            sum = sum * ex2((mx - mn) * sc) + ((part[0] + part[1]) + (part[2] + part[3]));
            mx = mn;
          }
        }
        {  // exchange with the warp that holds the other 64 columns of the same rows
          float2* xch = reinterpret_cast<float2*>(bias_s);  // [2][128] (no bias in this mode)
          xch[h * 128 + q * 32 + lane] = make_float2(mx, sum);
          named_bar_sync(2 + q, 64);
          const float2 o = xch[(h ^ 1) * 128 + q * 32 + lane];
          named_bar_sync(2 + q, 64);  // both have read: the slots may be rewritten next tile
          const float mn = fmaxf(mx, o.x);
          sum = sum * ex2((mx - mn) * sc) + o.y * ex2((o.x - mn) * sc);
          mx = mn;
        }
        if (nch > 0) {
          const float inv = 1.0f / sum, off = -mx * sc;
          sm_load(0);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c < nch) {
              ptx::tmem_ld_wait();
              if (c + 1 < nch) sm_load(c + 1);
              else release_tmem(as);
              const int col0 = cbase + c * NCOL;
              const bool full = col0 + NCOL <= e.n_valid;
              float v[NCOL];
#pragma unroll
              for (int i = 0; i < NCOL; ++i)
                v[i] = (full || col0 + i < e.n_valid)
                           ? ex2(fmaf(__uint_as_float(r[c & 1][i]), sc, off)) * inv : 0.f;
              const uint32_t box = box0 + static_cast<uint32_t>(c & 1) * BOX_BYTES;
              box_free(c == 0);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::sts_v4u(box + piece(k), pack_bf16x2(v[8 * k], v[8 * k + 1]),
                             pack_bf16x2(v[8 * k + 2], v[8 * k + 3]), pack_bf16x2(v[8 * k + 4], v[8 * k + 5]),
                             pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
              box_store(&tma_c0, box, col0);
            }
          }
        } else {
          release_tmem(as);
        }
      } else {
      if (active) tmem_load(0);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int col0 = n0 + c * NCOL;
        if (active && col0 < p.N) {
          ptx::tmem_ld_wait();
          if (c + 1 < NCHUNK && col0 + NCOL < p.N) {
            tmem_load(c + 1);
          } else {
            // every accumulator value of this warp is in registers: hand the TMEM stage back
            release_tmem(as);
          }
          float v[NCOL];
          {
            const uint32_t baddr = bias_addr + static_cast<uint32_t>(as * BN + h * (BN / 2) + c * NCOL) * 4u;
            if constexpr (fold) {
              // out = rstd * (acc - mean * cw[n]) + bias'[n]  (alpha == 1, checked on the host)
#pragma unroll
              for (int j = 0; j < NCOL / 4; ++j) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_bias) b4 = ptx::lds_v4(baddr + 16 * j);
                // columns >= N are clipped by the store: any in-range address will do for them
                const float4 w4 = ptx::lds_v4(cw_addr + (baddr - bias_addr) + 16 * j);
                const uint64_t nm2 = bc2(ln_nmean), rs2 = bc2(ln_rstd);
                up2(fma2(fma2(nm2, pk2(w4.x, w4.y),
                              pk2(__uint_as_float(r[c & 1][4 * j + 0]), __uint_as_float(r[c & 1][4 * j + 1]))),
                         rs2, pk2(b4.x, b4.y)), v[4 * j + 0], v[4 * j + 1]);
                up2(fma2(fma2(nm2, pk2(w4.z, w4.w),
                              pk2(__uint_as_float(r[c & 1][4 * j + 2]), __uint_as_float(r[c & 1][4 * j + 3]))),
                         rs2, pk2(b4.z, b4.w)), v[4 * j + 2], v[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < NCOL / 4; ++j) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_bias) b4 = ptx::lds_v4(baddr + 16 * j);
                const uint64_t al2 = bc2(e.alpha);
                up2(fma2(pk2(__uint_as_float(r[c & 1][4 * j + 0]), __uint_as_float(r[c & 1][4 * j + 1])), al2,
                         pk2(b4.x, b4.y)), v[4 * j + 0], v[4 * j + 1]);
                up2(fma2(pk2(__uint_as_float(r[c & 1][4 * j + 2]), __uint_as_float(r[c & 1][4 * j + 3])), al2,
                         pk2(b4.z, b4.w)), v[4 * j + 2], v[4 * j + 3]);
              }
            }
          }
          const uint32_t box = box0 + static_cast<uint32_t>(c & 1) * BOX_BYTES;
          // dropout between c_fc and the activation (src/model.py:195-201, training with p > 0
          // only): the mask is a pure function of (site seed, row * N + column) -- csrc/dropout.cuh.
          // Forward: z <- dropout(z) before it is saved and activated; backward: d act(dropout(z))
          // / dz = act'(z_saved) * mask / (1 - p), the mask regenerated here.
          if constexpr (MODE == EPI_QUICKGELU || MODE == EPI_DGELU) {
            if (e.drop.on()) {
              const unsigned int ibase = static_cast<unsigned int>(m0 + lane) * static_cast<unsigned int>(p.N) +
                                         static_cast<unsigned int>(col0);
#pragma unroll
              for (int i = 0; i < NCOL; ++i) v[i] *= e.drop.mult(ibase + i);
            }
          }
          if constexpr (MODE == EPI_RESID_LN) {
            const uint32_t fb = box0 + static_cast<uint32_t>(c % 3) * BOX_BYTES;
            ptx::mbar_wait(&my_aux[c % 3], aux_phase[c % 3]);
            aux_phase[c % 3] ^= 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t a = fb + piece(k);
              const float4 x4 = ptx::lds_v4(a);
              v[4 * k + 0] += x4.x;
              v[4 * k + 1] += x4.y;
              v[4 * k + 2] += x4.z;
              v[4 * k + 3] += x4.w;
              ptx::sts_v4(a, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
            }
#pragma unroll
            for (int i = 0; i < NCOL; ++i) {
              st_s1 += v[i];
              st_s2 = fmaf(v[i], v[i], st_s2);
            }
            box_store(&tma_c0, fb, col0);
            if (e.out2 != nullptr) {
              // bf16 copy: two 16-column chunks share one [32 x 64 B] box (pieces 0-1 / 2-3)
              const uint32_t hb = box0 + 3u * BOX_BYTES;
              if ((c & 1) == 0) {  // the store of the previous pair has read the box
                if (lane == 0) ptx::bulk_wait_read<1>();  // (only the fp32 store above may be pending)
                __syncwarp();
              }
              ptx::sts_v4u(hb + piece(2 * (c & 1)), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                           pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              ptx::sts_v4u(hb + piece(2 * (c & 1) + 1), pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                           pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
              if ((c & 1) == 1 || col0 + NCOL >= p.N) box_store(&tma_c1, hb, col0 - (c & 1) * NCOL);
            }
            // refill: chunk c + 2 goes to the box chunk c - 1 was stored from; at most its bf16
            // store and this chunk's two stores were committed after that one
            if (c + 2 < NCHUNK && col0 + 2 * NCOL < p.N && lane == 0) {
              if (e.out2 != nullptr) ptx::bulk_wait_read<2>();
              else ptx::bulk_wait_read<1>();
              load_x(c + 2);
            }
          } else if constexpr (!OBF) {
            box_free(c == 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::sts_v4(box + piece(k), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
            box_store(&tma_c0, box, col0);
          } else if constexpr (MODE == EPI_DGELU) {
            ptx::mbar_wait(&my_aux[c & 1], aux_phase[c & 1]);
            aux_phase[c & 1] ^= 1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t a = box + piece(k);
              const uint4 z = ptx::lds_v4u(a);
              const uint32_t zz[4] = {z.x, z.y, z.z, z.w};
              uint32_t o[4];
              if (e.act == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  erf_gelu_grad_pair(v[8 * k + 2 * i], v[8 * k + 2 * i + 1], bf16_lo(zz[i]), bf16_hi(zz[i]));
                  o[i] = pack_bf16x2(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  quick_gelu_grad_pair(v[8 * k + 2 * i], v[8 * k + 2 * i + 1], bf16_lo(zz[i]), bf16_hi(zz[i]));
                  o[i] = pack_bf16x2(v[8 * k + 2 * i], v[8 * k + 2 * i + 1]);
                }
              }
              ptx::sts_v4u(a, o[0], o[1], o[2], o[3]);
            }
            box_store(&tma_c0, box, col0);
            if (c + 2 < NCHUNK && col0 + 2 * NCOL < p.N && lane == 0) {
              ptx::bulk_wait_read<0>();  // the store above has read the box: refill it
              load_z(c + 2);
            }
          } else if constexpr (MODE == EPI_QUICKGELU) {
            if (e.out != nullptr) {  // training: z -> box 0, u -> box 1, each its own bulk group
              box_free(c == 0);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::sts_v4u(box0 + piece(k), pack_bf16x2(v[8 * k], v[8 * k + 1]),
                             pack_bf16x2(v[8 * k + 2], v[8 * k + 3]),
                             pack_bf16x2(v[8 * k + 4], v[8 * k + 5]),
                             pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
              box_store(&tma_c0, box0, col0);
            }
            if (e.act == 1) {
#pragma unroll
              for (int i = 0; i < NCOL; i += 2) erf_gelu_pair(v[i], v[i + 1]);
            } else {
#pragma unroll
              for (int i = 0; i < NCOL; i += 2) quick_gelu_pair(v[i], v[i + 1]);
            }
            const uint32_t ubox = e.out != nullptr ? box0 + BOX_BYTES : box;
            box_free(e.out == nullptr && c == 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::sts_v4u(ubox + piece(k), pack_bf16x2(v[8 * k], v[8 * k + 1]),
                           pack_bf16x2(v[8 * k + 2], v[8 * k + 3]), pack_bf16x2(v[8 * k + 4], v[8 * k + 5]),
                           pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
            box_store(&tma_c1, ubox, col0);
          } else {
            box_free(c == 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::sts_v4u(box + piece(k), pack_bf16x2(v[8 * k], v[8 * k + 1]),
                           pack_bf16x2(v[8 * k + 2], v[8 * k + 3]), pack_bf16x2(v[8 * k + 4], v[8 * k + 5]),
                           pack_bf16x2(v[8 * k + 6], v[8 * k + 7]));
            box_store(&tma_c0, box, col0);
          }
        }
      }
      if (!active) release_tmem(as);  // nothing to read: still one arrival per warp per tile
      if (MODE == EPI_RESID_LN && active && e.stats_out != nullptr && m0 + lane < p.M)
        reinterpret_cast<float2*>(e.stats_out)[static_cast<long long>(m0 + lane) * e.stats_nt + n0 / (BN / 2)] =
            make_float2(st_s1, st_s2);
      }  // MODE != EPI_SOFTMAX
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) ptx::bulk_wait<0>();  // all stores of this warp have completed at exit
  }

  ptx::tc_fence_before();
  if constexpr (NCTA == 2) ptx::cluster_sync_all();  // neither CTA may leave while its peer works
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (NCTA == 2) ptx::tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace gemm

}  // namespace mmu
