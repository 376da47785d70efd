// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring (4 stages)
//   -> tcgen05.mma (one elected thread, 128x256x16 UMMA, fp32 accumulators in TMEM,
//      two accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1)
//   -> tcgen05.ld -> fused epilogue (bias / QuickGELU / residual / dGELU / split-K atomics).
//
// Computes  C[M,N] = epilogue( sum_k A(m,k) * B(n,k) ).
// Either operand may be K-major (stored [rows][K], K contiguous) or MN-major (stored
// [K][rows], rows contiguous): the latter is what the weight-gradient GEMMs need
// (dW = dY^T X reduces over the row axis of both operands) and what dgrad needs for W.
//
// This is the tensor-core path for every dense contraction of the fusion model
// (reference call sites: src/model.py:262,264 projections; :193 MHA in/out proj;
// :195-201 MLP c_fc/c_proj) and their backward passes.
#pragma once
#include "gemm_api.h"
#include "ptx.cuh"

namespace mmu {

namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int STG_LD = 32;                         // staging row (floats), XOR-swizzled 16 B chunks
constexpr int STG_WARP_BYTES = 32 * STG_LD * 4;    // 4096 B per epilogue warp
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + NUM_EPI_WARPS) * 32;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int TMEM_COLS = 2 * BN;                       // two accumulator stages (512 = all of TMEM)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * STG_WARP_BYTES + 256 /*barriers*/ +
                           1024 /*alignment slack*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

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

__device__ __forceinline__ void store4(void* base, int is_bf16, long long off, float4 v) {
  if (is_bf16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(base) + off) = pk;
  } else {
    *reinterpret_cast<float4*>(static_cast<float*>(base) + off) = v;
  }
}

__device__ __forceinline__ float4 load4_bf16(const void* base, long long off) {
  const uint2 pk = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(base) + off);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int MODE>
struct AuxT { using type = int; };
template <>
struct AuxT<EPI_RESIDUAL> { using type = float4; };
template <>
struct AuxT<EPI_DGELU> { using type = uint2; };

// asm volatile: the prefetch must ISSUE where it is written (ptxas otherwise sinks plain loads
// towards their first use under register pressure, exposing the full DRAM latency).
template <int MODE>
__device__ __forceinline__ typename AuxT<MODE>::type load_aux(const GemmEpilogue& e, long long orow,
                                                              int gcol) {
  if constexpr (MODE == EPI_RESIDUAL) {
    const float* p = static_cast<const float*>(e.aux) + orow * e.ld_aux + gcol;
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
  } else if constexpr (MODE == EPI_DGELU) {
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(e.aux) + orow * e.ld_aux + gcol;
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
  } else {
    return 0;
  }
}

__device__ __forceinline__ float4 unpack_bf16x4(uint2 pk) {
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&pk.x);
  const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <int MODE>
__device__ __forceinline__ void epi_apply(const GemmEpilogue& e, long long orow, int gcol, float4 v,
                                          float4 bias, typename AuxT<MODE>::type aux) {
  v.x = fmaf(v.x, e.alpha, bias.x); v.y = fmaf(v.y, e.alpha, bias.y);
  v.z = fmaf(v.z, e.alpha, bias.z); v.w = fmaf(v.w, e.alpha, bias.w);
  if constexpr (MODE == EPI_STORE) {
    store4(e.out, e.out_bf16, orow * e.ld_out + gcol, v);
  } else if constexpr (MODE == EPI_QUICKGELU) {
    if (e.out != nullptr) store4(e.out, e.out_bf16, orow * e.ld_out + gcol, v);
    float4 u;
    u.x = quick_gelu(v.x); u.y = quick_gelu(v.y); u.z = quick_gelu(v.z); u.w = quick_gelu(v.w);
    store4(e.out2, e.out_bf16, orow * e.ld_out2 + gcol, u);
  } else if constexpr (MODE == EPI_RESIDUAL) {
    v.x += aux.x; v.y += aux.y; v.z += aux.z; v.w += aux.w;
    store4(e.out, e.out_bf16, orow * e.ld_out + gcol, v);
  } else if constexpr (MODE == EPI_DGELU) {
    const float4 z = unpack_bf16x4(aux);
    v.x *= quick_gelu_grad(z.x); v.y *= quick_gelu_grad(z.y);
    v.z *= quick_gelu_grad(z.z); v.w *= quick_gelu_grad(z.w);
    store4(e.out, e.out_bf16, orow * e.ld_out + gcol, v);
  } else if constexpr (MODE == EPI_ATOMIC) {
    ptx::red_add_v4(static_cast<float*>(e.out) + orow * e.ld_out + gcol, v);
  }
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b, const GemmProblem p,
                         const GemmEpilogue e) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  float* smem_stg = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES +
                                               NUM_EPI_WARPS * STG_WARP_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  const int nbatch = p.batch > 0 ? p.batch : 1;
  const int num_tiles = nbatch * m_tiles * n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], NUM_EPI_WARPS);
    }
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int mnb = tile / p.splits;
        const int g = mnb / (m_tiles * n_tiles);
        const int mn = mnb % (m_tiles * n_tiles);
        const int m0 = (mn / n_tiles) * BM;
        const int n0 = (mn % n_tiles) * BN;
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        const int a_mid = p.batch > 0 ? g / p.a_hdiv : 0;
        const int a_off = p.batch > 0 ? (g % p.a_hdiv) * p.a_hstride + p.a_col0 : 0;
        const int b_mid = p.batch > 0 ? g / p.b_hdiv : 0;
        const int b_off = p.batch > 0 ? (g % p.b_hdiv) * p.b_hstride + p.b_col0 : 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          const int k0 = kb * BK;
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
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
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
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              ptx::tma_load_2d(sb + j * 8192, &tma_b, &full_bar[stage], n0 + 64 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(BM, BN, p.a_mn_major, p.b_mn_major);
      // K-major SW128: 8-row groups 1024 B apart, k advances 32 B inside the swizzle atom.
      // MN-major SW128: 64-element MN groups 8192 B apart (one TMA box), 8-k groups 1024 B apart.
      const uint32_t a_lbo = p.a_mn_major ? 8192u : 16u, a_kstep = p.a_mn_major ? 2048u : 32u;
      const uint32_t b_lbo = p.b_mn_major ? 8192u : 16u, b_kstep = p.b_mn_major ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int kb0 = split * kb_per;
        const int kb1 = min(kb_total, kb0 + kb_per);
        ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = ptx::make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024u);
            const uint64_t bdesc = ptx::make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024u);
            ptx::umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    // 8 warps: TMEM lane quadrant q = warp % 4 (hardware rule), column half h = (warp-2)/4.
    // Per 32-column chunk: tcgen05.ld -> padded smem transpose -> 128-bit coalesced row stores.
    // The TMEM load and the aux (residual / z) loads of chunk c+1 are issued before chunk c is
    // written out, so their latency overlaps the store phase.
    using Aux = typename AuxT<MODE>::type;
    // aux prefetch depth: 3 chunks for the 8-byte bf16 z (DGELU), 2 for the 16-byte residual
    constexpr int NBUF = (MODE == EPI_DGELU) ? 3 : 2;
    const int q = warp & 3;
    const int h = (warp - 2) >> 2;
    const uint32_t stg = ptx::smem_u32(smem_stg + (warp - 2) * (32 * STG_LD));
    const int rr = lane >> 3;        // 0..3
    const int cc = (lane & 7) * 4;   // 0..28
    constexpr int NCHUNK = BN / 2 / 32;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mnb = tile / p.splits;
      const int g = mnb / (m_tiles * n_tiles);
      const int mn = mnb % (m_tiles * n_tiles);
      const int m0 = (mn / n_tiles) * BM + q * 32;
      const int n0 = (mn % n_tiles) * BN + h * (BN / 2);
      GemmEpilogue eb = e;
      if (p.batch > 0) {
        const long long boff = (long long)(g / p.out_hdiv) * p.out_mid_stride +
                               (long long)(g % p.out_hdiv) * p.out_hstride;
        eb.out = e.out_bf16 ? static_cast<void*>(static_cast<__nv_bfloat16*>(e.out) + boff)
                            : static_cast<void*>(static_cast<float*>(e.out) + boff);
      }
      int orow[8];
      bool rok[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int grow = m0 + rr + 4 * i;
        rok[i] = grow < p.M;
        orow[i] = grow;
        if (e.seg_len > 0)
          orow[i] = (grow / e.seg_len) * e.seg_stride + e.seg_off + grow % e.seg_len;
      }
      Aux aux[NBUF][8];
      auto prefetch_aux = [&](int c) {
        if (c < NCHUNK && n0 + c * 32 + cc < p.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (rok[i]) aux[c % NBUF][i] = load_aux<MODE>(eb, orow[i], n0 + c * 32 + cc);
        }
      };
      // issued BEFORE waiting for the accumulator: the loads fly while the MMAs finish
#pragma unroll
      for (int c = 0; c < NBUF - 1; ++c) prefetch_aux(c);

      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * BN + h * (BN / 2));
      uint32_t r[32];
      if (n0 < p.N) ptx::tmem_ld_32x32(taddr, r);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 < p.N) {
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            ptx::sts_v4(stg + (lane * STG_LD + 4 * (j ^ (lane & 7))) * 4, v);
          }
          __syncwarp();
          if (c + 1 < NCHUNK && col0 + 32 < p.N) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, r);
          prefetch_aux(c + NBUF - 1);
          const int gcol = col0 + cc;
          if (gcol < p.N) {
            float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias != nullptr) bias = *reinterpret_cast<const float4*>(e.bias + gcol);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int srow = rr + 4 * i;
              const float4 v = ptx::lds_v4(stg + (srow * STG_LD + 4 * ((lane & 7) ^ (srow & 7))) * 4);
              if (rok[i]) epi_apply<MODE>(eb, orow[i], gcol, v, bias, aux[c % NBUF][i]);
            }
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace gemm

}  // namespace mmu
