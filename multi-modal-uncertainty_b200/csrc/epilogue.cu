// Fused softmax / cross-entropy(+gradient) / accuracy / predictive-entropy / expected-entropy /
// mutual-information / ECE-histogram epilogue over logits (N, E, C).
//
// Reference semantics reproduced (fp32): CrossEntropyLoss on (B*E, C) rows in train mode and on
// the head-mean LOGITS in eval mode (src/model.py:293-304); `acc` = first-index argmax of the
// same (train.py:119-130).  The uncertainty scores and histograms have no reference code
// (SURVEY.md section 0); definitions are those of oracle/uncertainty.py.
//
// HBM-bound design: a CTA streams tiles of TS consecutive samples (TS*E*C*4 bytes, contiguous,
// 16-byte aligned because TS % 4 == 0) into shared memory with 1-D bulk TMA copies
// (cp.async.bulk, double buffered, mbarrier completion), so every DRAM access is a full-width
// burst regardless of the 2020-byte row pitch of the E*C = 505 case.  G = 8 lanes cooperate on
// one sample (4 samples per warp), which cuts the shuffle count per sample ~7x versus a warp per
// sample; in train mode the gradient overwrites the tile in place and leaves through a bulk
// TMA store.  Histogram bins are accumulated in shared memory and flushed once per CTA.
#include <cstdio>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace mmu {
namespace epi {

constexpr int THREADS = 256;
constexpr int CONF_BINS = 15;
constexpr int SCORE_BINS = 32;

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

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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
  float conf_sum[CONF_BINS];
  unsigned int n_samples, n_rows, n_correct_rows, n_correct_prob;
  float loss_sum, sum_h_pred, sum_h_exp, sum_mi;
};

template <int G, int CPL>
__global__ void __launch_bounds__(THREADS)
ce_uncertainty_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                      int label_stride, int label_estride, int N, int E, int C, int mode,
                      float grad_scale, float* __restrict__ dlogits, int* __restrict__ pred_out,
                      float* __restrict__ scores_out, MetricAccum* __restrict__ acc, int TS) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bars[2];
  __shared__ BlockAcc bacc;
  const int EC = E * C;
  const size_t tile_floats = (static_cast<size_t>(TS) * EC + 3) & ~static_cast<size_t>(3);
  float* buf[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + tile_floats};

  const int tid = threadIdx.x;
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_mbar_init();
    ptx::fence_proxy_async();
  }
  for (int i = tid; i < static_cast<int>(sizeof(BlockAcc) / 4); i += THREADS)
    reinterpret_cast<unsigned int*>(&bacc)[i] = 0u;
  __syncthreads();

  const int num_tiles = (N + TS - 1) / TS;
  auto tile_rows = [&](int t) { return min(TS, N - t * TS); };
  auto tile_bytes = [&](int t) { return static_cast<uint32_t>(tile_rows(t)) * EC * 4u; };
  auto issue = [&](int t, int b) {  // thread 0 only
    const uint32_t bytes = tile_bytes(t);
    if ((bytes & 15u) == 0) {
      ptx::mbar_arrive_expect_tx(&bars[b], bytes);
      bulk_load(buf[b], logits + static_cast<size_t>(t) * TS * EC, bytes, &bars[b]);
    } else {
      ptx::mbar_arrive(&bars[b]);  // ragged tail: loaded cooperatively by all threads below
    }
  };

  constexpr int NG = THREADS / G;
  const int gidx = tid / G, sub = tid % G;
  float t_loss = 0.f, t_hp = 0.f, t_he = 0.f, t_mi = 0.f;
  unsigned int t_rows = 0, t_corr_rows = 0, t_corr_prob = 0, t_n = 0;

  int it = 0;
  if (tid == 0 && blockIdx.x < num_tiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int cur = it & 1;
    const int next = tile + gridDim.x;
    if (tid == 0 && next < num_tiles) {
      if (dlogits != nullptr) bulk_store_wait_read();  // buffer cur^1 may still be draining
      issue(next, cur ^ 1);
    }
    ptx::mbar_wait(&bars[cur], (it >> 1) & 1);
    const int rows = tile_rows(tile);
    float* tb = buf[cur];
    if ((tile_bytes(tile) & 15u) != 0) {
      const float* src = logits + static_cast<size_t>(tile) * TS * EC;
      for (int i = tid; i < rows * EC; i += THREADS) tb[i] = src[i];
      __syncthreads();
    }

    for (int s0 = 0; s0 < rows; s0 += NG) {
      const int s = s0 + gidx;
      const bool valid = s < rows;
      const int sc = valid ? s : rows - 1;
      const int n = tile * TS + sc;
      float* zs = tb + static_cast<size_t>(sc) * EC;
      float pbar[CPL], zbar[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) { pbar[j] = 0.f; zbar[j] = 0.f; }
      float hexp = 0.f, loss = 0.f;
      unsigned int corr_rows = 0;
      for (int e = 0; e < E; ++e) {
        float* ze = zs + e * C;
        float zv[CPL];
        float m = -INFINITY;
        int am = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = sub + G * j;
          zv[j] = c < C ? ze[c] : -INFINITY;
          if (zv[j] > m) { m = zv[j]; am = c; }
        }
        group_argmax<G>(m, am);
        float sum = 0.f, sz = 0.f, ev[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = sub + G * j;
          ev[j] = c < C ? __expf(zv[j] - m) : 0.f;
          sum += ev[j];
          sz += c < C ? ev[j] * zv[j] : 0.f;
        }
        sum = group_sum<G>(sum);
        sz = group_sum<G>(sz);
        const float inv = 1.0f / sum;
        const float logZ = m + __logf(sum);
        hexp += logZ - sz * inv;
        int y = 0;
        if (mode == 0) {
          y = static_cast<int>(labels[static_cast<size_t>(n) * label_stride +
                                      static_cast<size_t>(e) * label_estride]);
          loss += logZ - ze[y];
          corr_rows += (am == y) ? 1u : 0u;
        }
        __syncwarp();  // every lane of the group has read ze[y] before the in-place overwrite
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = sub + G * j;
          if (c < C) {
            const float p = ev[j] * inv;
            pbar[j] += p;
            zbar[j] += zv[j];
            if (mode == 0 && dlogits != nullptr && valid)
              ze[c] = (p - (c == y ? 1.f : 0.f)) * grad_scale;
          }
        }
      }
      // ---- ensemble scores
      const float invE = 1.0f / static_cast<float>(E);
      float conf = -1.f;
      int pred = 0x7fffffff;
      float hp = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = sub + G * j;
        if (c < C) {
          const float p = pbar[j] * invE;
          if (p > conf) { conf = p; pred = c; }
          hp -= p > 0.f ? p * __logf(p) : 0.f;
        }
      }
      group_argmax<G>(conf, pred);
      hp = group_sum<G>(hp);
      const float he = hexp * invE;
      const float mi = hp - he;
      int pred_acc = pred;  // prediction that feeds `acc`
      const int y0 = static_cast<int>(labels[static_cast<size_t>(n) * label_stride]);
      if (mode == 1) {
        float m2 = -INFINITY;
        int a2 = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = sub + G * j;
          if (c < C) {
            zbar[j] *= invE;
            if (zbar[j] > m2) { m2 = zbar[j]; a2 = c; }
          }
        }
        group_argmax<G>(m2, a2);
        float s2 = 0.f, pick = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = sub + G * j;
          if (c < C) {
            s2 += __expf(zbar[j] - m2);
            pick += c == y0 ? zbar[j] : 0.f;
          }
        }
        s2 = group_sum<G>(s2);
        pick = group_sum<G>(pick);
        loss = m2 + __logf(s2) - pick;
        pred_acc = a2;
        corr_rows = (a2 == y0) ? 1u : 0u;
      }
      if (valid && sub == 0) {
        t_loss += loss;
        t_rows += mode == 0 ? static_cast<unsigned int>(E) : 1u;
        t_corr_rows += corr_rows;
        t_corr_prob += (pred == y0) ? 1u : 0u;
        t_n += 1u;
        t_hp += hp; t_he += he; t_mi += mi;
        const int cb = bin_of(conf, 1.0f, CONF_BINS);
        atomicAdd(&bacc.conf_count[cb], 1u);
        if (pred == y0) atomicAdd(&bacc.conf_correct[cb], 1u);
        atomicAdd(&bacc.conf_sum[cb], conf);
        atomicAdd(&bacc.hpred_count[bin_of(hp, 1.0f / __logf(static_cast<float>(C)), SCORE_BINS)], 1u);
        atomicAdd(&bacc.mi_count[bin_of(mi, 1.0f / __logf(static_cast<float>(max(E, 2))), SCORE_BINS)], 1u);
        if (pred_out != nullptr) {
          pred_out[2 * static_cast<size_t>(n)] = pred_acc;
          pred_out[2 * static_cast<size_t>(n) + 1] = pred;
        }
        if (scores_out != nullptr)
          *reinterpret_cast<float4*>(scores_out + 4 * static_cast<size_t>(n)) =
              make_float4(conf, hp, he, mi);
      }
    }

    if (mode == 0 && dlogits != nullptr) {
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the bulk store
      __syncthreads();
      const uint32_t bytes = tile_bytes(tile);
      float* dst = dlogits + static_cast<size_t>(tile) * TS * EC;
      if ((bytes & 15u) == 0) {
        if (tid == 0) bulk_store(dst, tb, bytes);
      } else {
        for (int i = tid; i < rows * EC; i += THREADS) dst[i] = tb[i];
        __syncthreads();
      }
    } else {
      __syncthreads();
    }
  }
  if (tid == 0 && dlogits != nullptr) bulk_store_wait_all();

  // ---- block reduction of the scalar accumulators, then one flush per CTA
  atomicAdd(&bacc.loss_sum, t_loss);
  atomicAdd(&bacc.sum_h_pred, t_hp);
  atomicAdd(&bacc.sum_h_exp, t_he);
  atomicAdd(&bacc.sum_mi, t_mi);
  atomicAdd(&bacc.n_rows, t_rows);
  atomicAdd(&bacc.n_correct_rows, t_corr_rows);
  atomicAdd(&bacc.n_correct_prob, t_corr_prob);
  atomicAdd(&bacc.n_samples, t_n);
  __syncthreads();
  if (acc != nullptr) {
    if (tid < CONF_BINS) {
      atomicAdd(&acc->conf_count[tid], static_cast<unsigned long long>(bacc.conf_count[tid]));
      atomicAdd(&acc->conf_correct[tid], static_cast<unsigned long long>(bacc.conf_correct[tid]));
      atomicAdd(&acc->conf_sum[tid], static_cast<double>(bacc.conf_sum[tid]));
    }
    if (tid >= 32 && tid < 32 + SCORE_BINS) {
      atomicAdd(&acc->hpred_count[tid - 32], static_cast<unsigned long long>(bacc.hpred_count[tid - 32]));
      atomicAdd(&acc->mi_count[tid - 32], static_cast<unsigned long long>(bacc.mi_count[tid - 32]));
    }
    if (tid == 64) {
      atomicAdd(&acc->n_samples, static_cast<unsigned long long>(bacc.n_samples));
      atomicAdd(&acc->n_rows, static_cast<unsigned long long>(bacc.n_rows));
      atomicAdd(&acc->n_correct_rows, static_cast<unsigned long long>(bacc.n_correct_rows));
      atomicAdd(&acc->n_correct_prob, static_cast<unsigned long long>(bacc.n_correct_prob));
      atomicAdd(&acc->loss_sum, static_cast<double>(bacc.loss_sum));
      atomicAdd(&acc->sum_h_pred, static_cast<double>(bacc.sum_h_pred));
      atomicAdd(&acc->sum_h_exp, static_cast<double>(bacc.sum_h_exp));
      atomicAdd(&acc->sum_mi, static_cast<double>(bacc.sum_mi));
    }
  }
}

template <int G, int CPL>
int launch(const float* logits, const long long* labels, int ls, int les, int N, int E, int C,
           int mode, float gs, float* dlogits, int* pred_out, float* scores_out, MetricAccum* acc,
           cudaStream_t stream) {
  const int EC = E * C;
  int TS = (32768 / (EC * 4)) / 4 * 4;
  if (TS < 4) TS = 4;
  if (TS > 64) TS = 64;
  const size_t tile_floats = (static_cast<size_t>(TS) * EC + 3) & ~static_cast<size_t>(3);
  const size_t smem = 2 * tile_floats * sizeof(float);
  if (smem > 200 * 1024) return MMU_ERR_SHAPE;
  auto kernel = ce_uncertainty_kernel<G, CPL>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           static_cast<int>(smem)) != cudaSuccess)
    return MMU_ERR_CUDA;
  const int tiles = (N + TS - 1) / TS;
  int per_sm = static_cast<int>((220 * 1024) / (smem + 2048));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = sm_count() * per_sm;
  if (grid > tiles) grid = tiles;
  kernel<<<grid, THREADS, smem, stream>>>(logits, labels, ls, les, N, E, C, mode, gs, dlogits,
                                          pred_out, scores_out, acc, TS);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    fprintf(stderr, "mmu: ce_uncertainty launch failed: %s\n", cudaGetErrorString(err));
    return MMU_ERR_CUDA;
  }
  count_launch();
  return 0;
}

}  // namespace epi

int ce_uncertainty(const float* logits, const long long* labels, int label_stride,
                   int label_estride, int N, int E, int C, int mode, float grad_scale,
                   float* dlogits, int* pred_out, float* scores_out, MetricAccum* acc,
                   cudaStream_t stream) {
  if (N <= 0) return 0;
  if (E < 1 || C < 1 || (mode != 0 && mode != 1)) return MMU_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(logits) & 15) != 0) return MMU_ERR_ALIGN;
  if (dlogits != nullptr && (reinterpret_cast<uintptr_t>(dlogits) & 15) != 0) return MMU_ERR_ALIGN;
  if (C <= 32)
    return epi::launch<8, 4>(logits, labels, label_stride, label_estride, N, E, C, mode, grad_scale,
                             dlogits, pred_out, scores_out, acc, stream);
  if (C <= 128)
    return epi::launch<8, 16>(logits, labels, label_stride, label_estride, N, E, C, mode,
                              grad_scale, dlogits, pred_out, scores_out, acc, stream);
  if (C <= 512)
    return epi::launch<32, 16>(logits, labels, label_stride, label_estride, N, E, C, mode,
                               grad_scale, dlogits, pred_out, scores_out, acc, stream);
  return MMU_ERR_SHAPE;
}

}  // namespace mmu
