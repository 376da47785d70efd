// Fused softmax / cross-entropy(+gradient) / accuracy / predictive-entropy / expected-entropy /
// mutual-information / ECE-histogram epilogue over logits (N, E, C).
//
// Reference semantics reproduced (fp32): CrossEntropyLoss on (B*E, C) rows in train mode and on
// the head-mean LOGITS in eval mode (src/model.py:293-304); `acc` = first-index argmax of the
// same (train.py:119-130).  The uncertainty scores and histograms have no reference code
// (SURVEY.md section 0); definitions are those of oracle/uncertainty.py.
//
// HBM-bound design (2 028 algorithmic bytes per sample at E=5, C=101):
//  * every WARP runs its own double-buffered pipeline: lane 0 streams chunks of SPC consecutive
//    samples (SPC*E*C*4 bytes, contiguous, 16-byte aligned because SPC % 4 == 0) into the warp's
//    two private shared-memory slots with 1-D bulk TMA copies (cp.async.bulk + mbarrier), so every
//    DRAM access is a full-width burst regardless of the 2020-byte sample pitch, and no
//    block-wide barrier exists anywhere in the main loop (12 warps x 2 slots x 8 KB per SM keep
//    ~100 KB in flight per SM; the eval-mode kernels for E <= 5 heads hold a whole pass in
//    registers and run 16 warps x 1 slot instead -- see HB / SS at the kernel);
//  * G lanes cooperate on one sample (G*CPL >= C class slots; G=8, CPL=13 wastes 3 % at C=101),
//    reductions are G-wide warp shuffles;
//  * in train mode the gradient overwrites the slot in place and leaves through a bulk TMA store;
//  * histogram bins are accumulated in shared memory (integers; the per-bin confidence sum in
//    2^-32 fixed point, so block results do not depend on the order of the atomics) and flushed
//    once per CTA.
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace mmu {
namespace epi {

constexpr int MAX_WARPS = 12;     // two shared-memory slots per warp
constexpr int MAX_WARPS_SS = 16;  // single-slot eval kernels (SS): register-file bound (4 warps x 128 registers per scheduler)
constexpr int SMEM_BUDGET = 200 * 1024;
constexpr int CONF_BINS = 15;
constexpr int SCORE_BINS = 32;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float NEG_BIG = -1.0e30f;  // stands in for "no class here" (finite: 0 * NEG_BIG == -0)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          ptx::smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(ptx::smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// a / b, correctly rounded, for y == RN(1/b) and b a small integer (Markstein: one residual step on
// q0 = RN(a*y) gives RN(a/b)); replaces the IEEE division (FCHK + slow-path call per element).
__device__ __forceinline__ float div_small_int(float a, float b, float y) {
  const float q0 = a * y;
  const float r = fmaf(-q0, b, a);
  return fmaf(r, y, q0);
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// explicit shared-space accesses on a 32-bit address: keeps the per-head address a single running
// register (the compiler otherwise rebuilds the generic pointer from threadIdx every iteration)
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ int group_min_int(int v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// (max value, first index attaining it)
template <int G>
__device__ __forceinline__ void group_argmax(float& m, int& a) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, a, o);
    if (om > m || (om == m && oa < a)) {
      m = om;
      a = oa;
    }
  }
}

__device__ __forceinline__ int bin_of(float v, float scale_inv, int nbins) {
  const float t = floorf(v * scale_inv * static_cast<float>(nbins));
  return min(nbins - 1, max(0, static_cast<int>(t)));
}

struct BlockAcc {
  unsigned int conf_count[CONF_BINS], conf_correct[CONF_BINS];
  unsigned int hpred_count[SCORE_BINS], mi_count[SCORE_BINS];
  // sum of round(conf * 2^24), split into the low 12 bits and the rest: two native 32-bit shared
  // atomics instead of a 64-bit CAS loop; exact (order independent) up to 2^20 samples per block
  unsigned int conf_fx_lo[CONF_BINS], conf_fx_hi[CONF_BINS];
  unsigned int n_samples, n_rows, n_correct_rows, n_correct_prob;
  double loss_sum, sum_h_pred, sum_h_exp, sum_mi;
};

struct Args {
  const float* logits;
  const long long* labels;
  float* dlogits;
  int* pred_out;
  float* scores_out;
  MetricAccum* acc;
  int ls, les, N, E, C;
  int spc;          // samples per chunk (multiple of 4 and of 32/G)
  int slot_floats;  // floats per shared-memory slot (multiple of 4)
  int nwarps;
  int bulk_in, bulk_out;  // logits / dlogits base 16-byte aligned: chunks may move by bulk TMA
  int hb;                 // heads reduced together in eval mode (1 = one by one)
  int ss;                 // single-slot pipeline allowed (eval mode, hb == E)
  float grad_scale;
};

// MODE 0: train (CE per head row, optional gradient); MODE 1: eval (CE on the head-mean logits).
// EXACT: CPL == ceil(C / G), so only the last class slot of a lane can be out of range.
// HB: heads reduced together (eval mode, E % HB == 0).  With HB = 1 a warp walks the heads one by
// one and every head is a serial chain max-shuffles -> ex2 -> sum-shuffles -> rcp; ncu showed the
// eval kernel issuing on 64 % of the cycles with 3 warps per scheduler, stalled on exactly those
// chains (`wait` 1.4, `short_scoreboard` 0.8 per issue).  With HB > 1 the logits of HB heads sit in
// registers at once and each reduction step shuffles HB independent values back to back, so one
// warp keeps HB chains in flight.  Operation order per value is unchanged: results are bit-identical
// to HB = 1.
//
// SS (single slot; requires HB == E): all logits of a pass are in registers right after the loads,
// so the warp hands its slot back to the bulk copy engine BEFORE the arithmetic and needs one slot
// instead of two; the head-mean-logit CE is finished at once as well, so its 13 registers die
// before the per-head softmax starts.  16 warps per SM (4 per scheduler, 128 registers, no spills)
// instead of 12: measured 4.77 -> 5.45 TB/s at E=5, C=101 (20 warps = 96 registers spill 30 values
// per pass and are no faster than 12; 24 warps = 80 registers: 3.4 TB/s).
template <int G, int CPL, int MODE, bool GRAD, bool EXACT, int HB, bool SS>
__global__ void __launch_bounds__((SS ? MAX_WARPS_SS : MAX_WARPS) * 32)
ce_uncertainty_kernel(const Args a) {
  static_assert(HB == 1 || (MODE == 1 && !GRAD), "batched heads: eval mode only");
  static_assert(!SS || HB > 1, "single-slot mode needs all heads of a sample in registers");
  constexpr int NSLOT = SS ? 1 : 2;
  extern __shared__ __align__(128) float smem_f[];
  __shared__ uint64_t bars[SS ? MAX_WARPS_SS : MAX_WARPS * 2];
  __shared__ BlockAcc bacc;
  constexpr int PP = 32 / G;             // samples per warp pass
  constexpr int LPT = (16 + G - 1) / G;  // labels held per lane (E <= 16)
  constexpr unsigned FULL = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULL, tid >> 5, 0);  // tells the compiler it is warp-uniform
  const int g = lane / G, sub = lane % G;
  const int E = a.E, C = a.C, EC = E * C;

  for (int i = tid; i < static_cast<int>(sizeof(BlockAcc) / 4); i += blockDim.x)
    reinterpret_cast<unsigned int*>(&bacc)[i] = 0u;
  float* const slot0 = smem_f + static_cast<size_t>(warp) * NSLOT * a.slot_floats;
  uint64_t* bar = bars + warp * NSLOT;
  if (lane == 0) {
    ptx::mbar_init(&bar[0], 1);
    if (!SS) ptx::mbar_init(&bar[NSLOT - 1], 1);
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
  }
  __syncthreads();

  const int spc = a.spc;
  const int num_chunks = (a.N + spc - 1) / spc;
  const int gw = blockIdx.x * a.nwarps + warp, GW = gridDim.x * a.nwarps;
  auto chunk_rows = [&](int ch) { return min(spc, a.N - ch * spc); };
  auto issue = [&](int ch, int b) {  // lane 0 only; a ragged tail chunk is copied by the lanes
    const uint32_t bytes = static_cast<uint32_t>(chunk_rows(ch)) * EC * 4u;
    if (a.bulk_in && (bytes & 15u) == 0) {
      ptx::mbar_arrive_expect_tx(&bar[b], bytes);
      bulk_load(slot0 + b * a.slot_floats, a.logits + static_cast<size_t>(ch) * spc * EC, bytes,
                &bar[b]);
    }
  };

  float t_loss = 0.f, t_hp = 0.f, t_he = 0.f, t_mi = 0.f;
  unsigned int t_rows = 0, t_corr_rows = 0, t_corr_prob = 0, t_n = 0;
  const float fE = static_cast<float>(E);
  const float invE = 1.0f / fE;  // IEEE division: RN(1/E), as div_small_int requires
  const float lg2E = lg2(fE);
  const float inv_logC = 1.0f / __logf(static_cast<float>(C));
  const float inv_logE = 1.0f / __logf(static_cast<float>(max(E, 2)));

  auto load_labels = [&](int n, bool valid, int (&ylab)[LPT]) {
#pragma unroll
    for (int t = 0; t < LPT; ++t) {
      const int h = sub + G * t;
      ylab[t] = 0;
      if (MODE == 0) {
        if (valid && h < E)
          ylab[t] = static_cast<int>(a.labels[static_cast<size_t>(n) * a.ls + static_cast<size_t>(h) * a.les]);
      } else {
        if (valid && t == 0) ylab[0] = static_cast<int>(a.labels[static_cast<size_t>(n) * a.ls]);
      }
    }
  };

  int it = 0;
  int ylab_next[LPT];  // labels of the next chunk's first pass: fetched one iteration ahead
  if (gw < num_chunks) {
    if (lane == 0) issue(gw, 0);
    const int rows0 = chunk_rows(gw);
    if (MODE == 0 || SS) load_labels(gw * spc + (g < rows0 ? g : rows0 - 1), g < rows0, ylab_next);
  }
  for (int ch = gw; ch < num_chunks; ch += GW, ++it) {
    const int cur = SS ? 0 : it & 1;
    const int nxt = ch + GW;
    __syncwarp();  // every lane is done reading slot cur^1 (previous iteration)
    if (!SS && lane == 0 && nxt < num_chunks) {
      if (GRAD) bulk_store_wait_read();  // slot cur^1 may still be draining to dlogits
      issue(nxt, cur ^ 1);
    }
    const int rows = chunk_rows(ch);
    const uint32_t bytes = static_cast<uint32_t>(rows) * EC * 4u;
    float* tb = slot0 + cur * a.slot_floats;
    int ylab[LPT];
    if (MODE == 0 || SS) {
#pragma unroll
      for (int t = 0; t < LPT; ++t) ylab[t] = ylab_next[t];
      if (nxt < num_chunks) {  // in flight during the whole of this chunk's arithmetic
        const int rows1 = chunk_rows(nxt);
        load_labels(nxt * spc + (g < rows1 ? g : rows1 - 1), g < rows1, ylab_next);
      }
    }
    if (a.bulk_in && (bytes & 15u) == 0) {
      ptx::mbar_wait(&bar[cur], SS ? (it & 1) : ((it >> 1) & 1));
    } else {
      const float* src = a.logits + static_cast<size_t>(ch) * spc * EC;
      for (int i = lane; i < rows * EC; i += 32) tb[i] = src[i];
      __syncwarp();
    }

    for (int s0 = 0; s0 < rows; s0 += PP) {
      const int s = s0 + g;
      const bool valid = s < rows;
      const int sc = valid ? s : rows - 1;
      const int n = ch * spc + sc;
      if ((MODE == 1 && !SS) || s0 > 0) load_labels(n, valid, ylab);
      int y0_early = 0;
      float zy0_early = 0.f;
      if (SS) y0_early = __shfl_sync(FULL, ylab[0], lane & ~(G - 1));
      float pbar[CPL], zbar[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) { pbar[j] = 0.f; zbar[j] = 0.f; }
      float hexp = 0.f, loss = 0.f;
      unsigned int corr_rows = 0;
      const uint32_t zs_addr = ptx::smem_u32(tb) + static_cast<uint32_t>(sc * EC) * 4u;
      auto load_head = [&](uint32_t za, float (&z)[CPL]) {  // za: this lane's first class slot
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          if (EXACT && j < CPL - 1) z[j] = lds_f32(za + 4 * G * j);
          else z[j] = sub + G * j < C ? lds_f32(za + 4 * G * j) : NEG_BIG;
        }
      };
      // one head: zv holds its logits; the next head's logits are fetched into zn meanwhile
      auto do_head = [&](int e, uint32_t ze_addr, float (&zv)[CPL], float (&zn)[CPL]) {
        if (e + 1 < E) load_head(ze_addr + 4 * (C + sub), zn);
        int y = 0;
        float zy = 0.f;
        if (MODE == 0) {
          int yy = 0;
#pragma unroll
          for (int t = 0; t < LPT; ++t) yy = (e / G == t) ? ylab[t] : yy;
          y = __shfl_sync(FULL, yy, (lane & ~(G - 1)) + (e % G));
          zy = lds_f32(ze_addr + 4 * y);
        }
        float m = zv[0];
#pragma unroll
        for (int j = 1; j < CPL; ++j) m = fmaxf(m, zv[j]);
        m = group_max<G>(m);
        int am = 0;
        if (MODE == 0) {  // first index attaining the maximum
          am = 0x7fffffff;
#pragma unroll
          for (int j = CPL - 1; j >= 0; --j) am = (zv[j] == m) ? sub + G * j : am;
          am = group_min_int<G>(am);
        }
        const float mL = m * LOG2E;
        float sum0 = 0.f, sum1 = 0.f, sz0 = 0.f, sz1 = 0.f, ev[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          ev[j] = ex2(fmaf(zv[j], LOG2E, -mL));
          if (j & 1) { sum1 += ev[j]; sz1 = fmaf(ev[j], zv[j], sz1); }
          else { sum0 += ev[j]; sz0 = fmaf(ev[j], zv[j], sz0); }
          if (MODE == 1) zbar[j] += zv[j];
        }
        const float sum = group_sum<G>(sum0 + sum1);
        const float sz = group_sum<G>(sz0 + sz1);
        const float inv = rcp(sum);
        const float logZ = fmaf(lg2(sum), LN2, m);
        hexp += fmaf(-sz, inv, logZ);
        if (MODE == 0) {
          loss += logZ - zy;
          corr_rows += (am == y) ? 1u : 0u;
        }
        if (GRAD) __syncwarp();  // ze[y] has been read by the whole group before the overwrite
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          pbar[j] = fmaf(ev[j], inv, pbar[j]);
          if (GRAD) {
            const int c = sub + G * j;
            if (valid && c < C)
              sts_f32(ze_addr + 4 * c, (ev[j] * inv - (c == y ? 1.f : 0.f)) * a.grad_scale);
          }
        }
      };
      // eval mode: head-mean logits from their head-ordered sums zb[] (a correctly rounded
      // division = torch.mean), first-index argmax, log-sum-exp, CE against the label
      int a2_mean = 0;
      auto mean_logit_ce = [&](float (&zb)[CPL], float zy_sum) {
        float m2 = NEG_BIG;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = sub + G * j;
          zb[j] = ((EXACT && j < CPL - 1) || c < C) ? div_small_int(zb[j], fE, invE) : NEG_BIG;
          m2 = fmaxf(m2, zb[j]);
        }
        m2 = group_max<G>(m2);
        int a2 = 0x7fffffff;
#pragma unroll
        for (int j = CPL - 1; j >= 0; --j) a2 = (zb[j] == m2) ? sub + G * j : a2;
        a2_mean = group_min_int<G>(a2);
        const float m2L = m2 * LOG2E;
        float s20 = 0.f, s21 = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const float t = ex2(fmaf(zb[j], LOG2E, -m2L));
          if (j & 1) s21 += t; else s20 += t;
        }
        const float s2 = group_sum<G>(s20 + s21);
        loss = fmaf(lg2(s2), LN2, m2) - div_small_int(zy_sum, fE, invE);
      };
      // HB heads at once (eval mode): every step runs over the HB heads in its inner loop, so the
      // shuffles / MUFU results of different heads overlap
      auto do_heads = [&](uint32_t ze_addr) {
        float z[HB][CPL], m[HB], sum[HB], sz[HB];
#pragma unroll
        for (int h = 0; h < HB; ++h) load_head(ze_addr + 4 * (h * C + sub), z[h]);
#pragma unroll
        for (int h = 0; h < HB; ++h) {
          m[h] = z[h][0];
#pragma unroll
          for (int j = 1; j < CPL; ++j) m[h] = fmaxf(m[h], z[h][j]);
        }
        if constexpr (SS) {
          // the label's logit of every head (summed in head order, like zbar), then the slot goes
          // back to the copy engine: every value this pass needs from it is in registers (the
          // local maxima above consumed the logits; the bulk copy lands a DRAM round trip later)
#pragma unroll
          for (int h = 0; h < HB; ++h) zy0_early += lds_f32(ze_addr + 4 * (h * C + y0_early));
          if (s0 + PP >= rows) {
            __syncwarp();
            if (lane == 0 && nxt < num_chunks) issue(nxt, 0);
          }
          // all heads are here, so the head-mean logits are complete now: their 13 registers are
          // dead before the per-head softmax needs its own
#pragma unroll
          for (int h = 0; h < HB; ++h) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) zbar[j] += z[h][j];
          }
          mean_logit_ce(zbar, zy0_early);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
          for (int h = 0; h < HB; ++h) m[h] = fmaxf(m[h], __shfl_xor_sync(FULL, m[h], o));
        }
#pragma unroll
        for (int h = 0; h < HB; ++h) {
          const float mL = m[h] * LOG2E;
          float sum0 = 0.f, sum1 = 0.f, sz0 = 0.f, sz1 = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float ev = ex2(fmaf(z[h][j], LOG2E, -mL));
            if (j & 1) { sum1 += ev; sz1 = fmaf(ev, z[h][j], sz1); }
            else { sum0 += ev; sz0 = fmaf(ev, z[h][j], sz0); }
            if (!SS) zbar[j] += z[h][j];
            z[h][j] = ev;  // the logit is dead: its register now holds exp(z - max)
          }
          sum[h] = sum0 + sum1;
          sz[h] = sz0 + sz1;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
#pragma unroll
          for (int h = 0; h < HB; ++h) {
            sum[h] += __shfl_xor_sync(FULL, sum[h], o);
            sz[h] += __shfl_xor_sync(FULL, sz[h], o);
          }
        }
#pragma unroll
        for (int h = 0; h < HB; ++h) {
          const float inv = rcp(sum[h]);
          const float logZ = fmaf(lg2(sum[h]), LN2, m[h]);
          hexp += fmaf(-sz[h], inv, logZ);
          sum[h] = inv;
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
#pragma unroll
          for (int h = 0; h < HB; ++h) pbar[j] = fmaf(z[h][j], sum[h], pbar[j]);
        }
      };
      if constexpr (HB > 1) {
        uint32_t ze_addr = zs_addr;
        for (int e = 0; e < E; e += HB, ze_addr += 4 * HB * C) do_heads(ze_addr);
      } else {
        float za[CPL], zb[CPL];  // ping-pong register buffers: no copies between heads
        load_head(zs_addr + 4 * sub, za);
        uint32_t ze_addr = zs_addr;
        for (int e = 0; e < E; e += 2, ze_addr += 8 * C) {
          do_head(e, ze_addr, za, zb);
          if (e + 1 < E) do_head(e + 1, ze_addr + 4 * C, zb, za);
        }
      }
      // the label: MODE 0 holds it from the prefetch; MODE 1 only needs it from here on, so its
      // load (issued at the top of the pass) has had the whole head loop to land
      const int y0 = SS ? y0_early : __shfl_sync(FULL, ylab[0], lane & ~(G - 1));
      // ---- ensemble scores from pbar = sum over heads of p_k (p_bar = pbar / E):
      //      sum p_bar log2 p_bar = (1/E) sum pbar log2 pbar - log2 E   (sum p_bar = 1)
      float cmax = 0.f, h0 = 0.f, h1 = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        cmax = fmaxf(cmax, pbar[j]);  // out-of-range class slots hold exactly 0
        const float t = pbar[j] * lg2(fmaxf(pbar[j], 1e-37f));
        if (j & 1) h1 += t; else h0 += t;
      }
      cmax = group_max<G>(cmax);
      int pred = 0x7fffffff;
#pragma unroll
      for (int j = CPL - 1; j >= 0; --j) pred = (pbar[j] == cmax) ? sub + G * j : pred;
      pred = group_min_int<G>(pred);
      const float conf = cmax * invE;
      const float hp = -LN2 * fmaf(group_sum<G>(h0 + h1), invE, -lg2E);
      const float he = hexp * invE;
      const float mi = hp - he;
      int pred_acc = pred;  // prediction that feeds `acc`
      if (MODE == 1) {
        if (!SS) {
          float zy0 = 0.f;  // the label's head-mean logit, summed in the same order as zbar[]
          for (int e = 0; e < E; ++e) zy0 += lds_f32(zs_addr + 4 * (e * C + y0));
          mean_logit_ce(zbar, zy0);
        }
        const int a2 = a2_mean;
        pred_acc = a2;
        corr_rows = (a2 == y0) ? 1u : 0u;
      }
      if (valid && sub == 0) {
        t_loss += loss;
        t_rows += MODE == 0 ? static_cast<unsigned int>(E) : 1u;
        t_corr_rows += corr_rows;
        t_corr_prob += (pred == y0) ? 1u : 0u;
        t_n += 1u;
        t_hp += hp; t_he += he; t_mi += mi;
        const int cb = bin_of(conf, 1.0f, CONF_BINS);
        const unsigned int fx = __float2uint_rn(conf * 16777216.0f);
        atomicAdd(&bacc.conf_count[cb], 1u);
        if (pred == y0) atomicAdd(&bacc.conf_correct[cb], 1u);
        atomicAdd(&bacc.conf_fx_lo[cb], fx & 0xfffu);
        atomicAdd(&bacc.conf_fx_hi[cb], fx >> 12);
        atomicAdd(&bacc.hpred_count[bin_of(hp, inv_logC, SCORE_BINS)], 1u);
        atomicAdd(&bacc.mi_count[bin_of(mi, inv_logE, SCORE_BINS)], 1u);
        if (a.pred_out != nullptr)
          *reinterpret_cast<int2*>(a.pred_out + 2 * static_cast<size_t>(n)) = make_int2(pred_acc, pred);
        if (a.scores_out != nullptr)
          *reinterpret_cast<float4*>(a.scores_out + 4 * static_cast<size_t>(n)) =
              make_float4(conf, hp, he, mi);
      }
    }

    if (GRAD) {
      float* dst = a.dlogits + static_cast<size_t>(ch) * spc * EC;
      if (a.bulk_out && (bytes & 15u) == 0) {
        ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the bulk store
        __syncwarp();
        if (lane == 0) bulk_store(dst, tb, bytes);
      } else {
        __syncwarp();
        for (int i = lane; i < rows * EC; i += 32) dst[i] = tb[i];
      }
    }
  }
  if (GRAD && lane == 0) bulk_store_wait_all();

  // ---- warp reduction of the scalar accumulators (only sub == 0 lanes hold non-zero values)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    t_loss += __shfl_xor_sync(FULL, t_loss, o);
    t_hp += __shfl_xor_sync(FULL, t_hp, o);
    t_he += __shfl_xor_sync(FULL, t_he, o);
    t_mi += __shfl_xor_sync(FULL, t_mi, o);
    t_rows += __shfl_xor_sync(FULL, t_rows, o);
    t_corr_rows += __shfl_xor_sync(FULL, t_corr_rows, o);
    t_corr_prob += __shfl_xor_sync(FULL, t_corr_prob, o);
    t_n += __shfl_xor_sync(FULL, t_n, o);
  }
  if (lane == 0 && t_n > 0) {
    atomicAdd(&bacc.loss_sum, static_cast<double>(t_loss));
    atomicAdd(&bacc.sum_h_pred, static_cast<double>(t_hp));
    atomicAdd(&bacc.sum_h_exp, static_cast<double>(t_he));
    atomicAdd(&bacc.sum_mi, static_cast<double>(t_mi));
    atomicAdd(&bacc.n_rows, t_rows);
    atomicAdd(&bacc.n_correct_rows, t_corr_rows);
    atomicAdd(&bacc.n_correct_prob, t_corr_prob);
    atomicAdd(&bacc.n_samples, t_n);
  }
  __syncthreads();
  MetricAccum* acc = a.acc;
  if (acc != nullptr && bacc.n_samples > 0) {
    for (int i = tid; i < CONF_BINS; i += blockDim.x) {
      atomicAdd(&acc->conf_count[i], static_cast<unsigned long long>(bacc.conf_count[i]));
      atomicAdd(&acc->conf_correct[i], static_cast<unsigned long long>(bacc.conf_correct[i]));
      const unsigned long long fx = (static_cast<unsigned long long>(bacc.conf_fx_hi[i]) << 12) +
                                    bacc.conf_fx_lo[i];
      atomicAdd(&acc->conf_sum[i], static_cast<double>(fx) * (1.0 / 16777216.0));
    }
    for (int i = tid; i < SCORE_BINS; i += blockDim.x) {
      atomicAdd(&acc->hpred_count[i], static_cast<unsigned long long>(bacc.hpred_count[i]));
      atomicAdd(&acc->mi_count[i], static_cast<unsigned long long>(bacc.mi_count[i]));
    }
    if (tid == blockDim.x - 1) {
      atomicAdd(&acc->n_samples, static_cast<unsigned long long>(bacc.n_samples));
      atomicAdd(&acc->n_rows, static_cast<unsigned long long>(bacc.n_rows));
      atomicAdd(&acc->n_correct_rows, static_cast<unsigned long long>(bacc.n_correct_rows));
      atomicAdd(&acc->n_correct_prob, static_cast<unsigned long long>(bacc.n_correct_prob));
      atomicAdd(&acc->loss_sum, bacc.loss_sum);
      atomicAdd(&acc->sum_h_pred, bacc.sum_h_pred);
      atomicAdd(&acc->sum_h_exp, bacc.sum_h_exp);
      atomicAdd(&acc->sum_mi, bacc.sum_mi);
    }
  }
}

template <int G, int CPL, int MODE, bool GRAD, bool EXACT, int HB = 1, bool SS = false>
int launch_one(const Args& a0, cudaStream_t stream) {
  Args a = a0;
  constexpr int PP = 32 / G;
  const int EC = a.E * a.C;
  // chunk: a multiple of the pass width and of 4 samples (16-byte alignment), >= ~8 KB
  int unit = PP < 4 ? 4 : PP;
  int k = 8192 / (unit * EC * 4);
  if (k < 1) k = 1;
  a.spc = unit * k;
  a.slot_floats = (a.spc * EC + 3) & ~3;
  const size_t slot_bytes = static_cast<size_t>(a.slot_floats) * 4;
  constexpr int NSLOT = SS ? 1 : 2;
  int wmax = static_cast<int>(SMEM_BUDGET / (NSLOT * slot_bytes));
  if (wmax < 1) return MMU_ERR_SHAPE;
  if (wmax > (SS ? MAX_WARPS_SS : MAX_WARPS)) wmax = SS ? MAX_WARPS_SS : MAX_WARPS;
  const int chunks = (a.N + a.spc - 1) / a.spc;
  const int sms = sm_count();
  int w = (chunks + sms - 1) / sms;  // few chunks: spread them over the SMs, one warp each
  if (w > wmax) w = wmax;
  if (w < 1) w = 1;
  a.nwarps = w;
  int grid = (chunks + w - 1) / w;
  if (grid > sms) grid = sms;
  const size_t smem = static_cast<size_t>(w) * NSLOT * slot_bytes;
  auto kernel = ce_uncertainty_kernel<G, CPL, MODE, GRAD, EXACT, HB, SS>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           SMEM_BUDGET) != cudaSuccess)
    return MMU_ERR_CUDA;
  kernel<<<grid, w * 32, smem, stream>>>(a);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    fprintf(stderr, "mmu: ce_uncertainty launch failed: %s\n", cudaGetErrorString(err));
    return MMU_ERR_CUDA;
  }
  count_launch();
  return 0;
}

template <int G, int CPL, bool EXACT>
int launch(const Args& a, int mode, cudaStream_t stream) {
  if (mode == 1) {
    if constexpr (G * CPL <= 104 && CPL <= 13) {  // HB * CPL registers of logits per lane
      if (a.hb == a.E && a.ss) {
        if (a.hb == 5) return launch_one<G, CPL, 1, false, EXACT, 5, true>(a, stream);
        if (a.hb == 4) return launch_one<G, CPL, 1, false, EXACT, 4, true>(a, stream);
        if (a.hb == 3) return launch_one<G, CPL, 1, false, EXACT, 3, true>(a, stream);
        if (a.hb == 2) return launch_one<G, CPL, 1, false, EXACT, 2, true>(a, stream);
      }
      if (a.hb == 5) return launch_one<G, CPL, 1, false, EXACT, 5>(a, stream);
      if (a.hb == 4) return launch_one<G, CPL, 1, false, EXACT, 4>(a, stream);
      if (a.hb == 3) return launch_one<G, CPL, 1, false, EXACT, 3>(a, stream);
      if (a.hb == 2) return launch_one<G, CPL, 1, false, EXACT, 2>(a, stream);
    }
    return launch_one<G, CPL, 1, false, EXACT>(a, stream);
  }
  if (a.dlogits != nullptr) return launch_one<G, CPL, 0, true, EXACT>(a, stream);
  return launch_one<G, CPL, 0, false, EXACT>(a, stream);
}

}  // namespace epi

int ce_uncertainty(const float* logits, const long long* labels, int label_stride,
                   int label_estride, int N, int E, int C, int mode, float grad_scale,
                   float* dlogits, int* pred_out, float* scores_out, MetricAccum* acc,
                   cudaStream_t stream) {
  if (N <= 0) return 0;
  if (E < 1 || E > 16 || C < 1 || (mode != 0 && mode != 1)) return MMU_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(logits) & 3) != 0) return MMU_ERR_ALIGN;
  if (pred_out != nullptr && (reinterpret_cast<uintptr_t>(pred_out) & 7) != 0) return MMU_ERR_ALIGN;
  if (scores_out != nullptr && (reinterpret_cast<uintptr_t>(scores_out) & 15) != 0) return MMU_ERR_ALIGN;
  // the per-block fixed-point confidence sums are exact up to 2^20 samples per block: split
  // launches that could exceed it (>= 64 blocks whenever N is this large)
  constexpr int PIECE = 1 << 26;
  if (N > PIECE) {
    for (long long off = 0; off < N; off += PIECE) {
      const int n = static_cast<int>(N - off < PIECE ? N - off : PIECE);
      const int rc = ce_uncertainty(
          logits + off * E * C, labels + off * label_stride, label_stride, label_estride, n, E, C,
          mode, grad_scale, dlogits ? dlogits + off * E * C : nullptr,
          pred_out ? pred_out + 2 * off : nullptr, scores_out ? scores_out + 4 * off : nullptr, acc,
          stream);
      if (rc != 0) return rc;
    }
    return 0;
  }
  epi::Args a{};
  a.logits = logits; a.labels = labels; a.dlogits = mode == 0 ? dlogits : nullptr;
  a.pred_out = pred_out; a.scores_out = scores_out; a.acc = acc;
  a.ls = label_stride; a.les = label_estride; a.N = N; a.E = E; a.C = C;
  a.grad_scale = grad_scale;
  // eval mode: E <= 5 heads are reduced together, larger even E in pairs (MMU_CE_HB overrides:
  // the A/B switch of tools/bench_epilogue.py)
  a.hb = E <= 5 ? E : (E % 2 == 0 ? 2 : 1);
  if (const char* env = getenv("MMU_CE_HB")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 5 && E % v == 0) a.hb = v;
  }
  a.ss = 1;
  if (const char* env = getenv("MMU_CE_SS")) a.ss = atoi(env) != 0;
  // slices of a larger logits tensor need not be 16-byte aligned: such calls move their chunks
  // with ordinary loads / stores instead of bulk TMA copies
  a.bulk_in = (reinterpret_cast<uintptr_t>(logits) & 15) == 0;
  a.bulk_out = dlogits == nullptr || (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0;
  // (lanes per sample, class slots per lane): exact fits for the reference's class counts
  // (2: hateful memes, 10: FashionMNIST, 101: Food-101), predicated generic shapes otherwise.
  if (C <= 2) return epi::launch<1, 2, false>(a, mode, stream);
  if (C <= 4) return epi::launch<1, 4, false>(a, mode, stream);
  if (C == 9 || C == 10) return epi::launch<2, 5, true>(a, mode, stream);
  if (C <= 16) return epi::launch<2, 8, false>(a, mode, stream);
  if (C <= 32) return epi::launch<4, 8, false>(a, mode, stream);
  if (C <= 64) return epi::launch<8, 8, false>(a, mode, stream);
  if (C >= 97 && C <= 104) return epi::launch<8, 13, true>(a, mode, stream);
  if (C <= 128) return epi::launch<8, 16, false>(a, mode, stream);
  if (C <= 256) return epi::launch<16, 16, false>(a, mode, stream);
  if (C <= 512) return epi::launch<32, 16, false>(a, mode, stream);
  if (C <= 1024) return epi::launch<32, 32, false>(a, mode, stream);
  return MMU_ERR_SHAPE;
}

}  // namespace mmu
