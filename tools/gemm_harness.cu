// Standalone correctness + timing harness for the tcgen05 GEMM (no torch, no python).
// usage: gemm_harness <case-id>      (one case per process so a fault cannot mask others)
#include <cmath>
#include <cstdint>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#include "../multi-modal-uncertainty_b200/csrc/common.h"
#include "../multi-modal-uncertainty_b200/csrc/gemm_api.h"

using namespace mmu;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t err__ = (x);                                                       \
    if (err__ != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(err__), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

struct Case {
  const char* name;
  int M, N, K, a_mn, b_mn, mode, out_bf16, splits, bias, seg, timed;
  int fold;    // LayerNorm folded into the epilogue (EPI_STORE / EPI_QUICKGELU)
  int no_out;  // QUICKGELU: eval mode (u only); RESID_LN: no bf16 copy
  int act;     // QUICKGELU / DGELU: 1 = erf-GELU (the MMBT path)
};

static const Case kCases[] = {
    {"kk_small_f32", 256, 256, 128, 0, 0, EPI_STORE, 0, 1, 0, 0, 0},
    {"kk_1tile_k64", 128, 256, 64, 0, 0, EPI_STORE, 0, 1, 0, 0, 0},
    {"kk_tails_f32", 1000, 520, 200, 0, 0, EPI_STORE, 0, 1, 1, 0, 0},
    {"mnA_kB_small", 256, 256, 128, 1, 0, EPI_STORE, 0, 1, 0, 0, 0},
    {"kA_mnB_small", 256, 256, 128, 0, 1, EPI_STORE, 0, 1, 0, 0, 0},
    {"mn_mn_small", 256, 512, 256, 1, 1, EPI_STORE, 0, 1, 0, 0, 0},
    {"proj_bf16_bias", 30336, 768, 768, 0, 0, EPI_STORE, 1, 1, 1, 0, 1},
    {"inproj_bf16", 30336, 2304, 768, 0, 0, EPI_STORE, 1, 1, 1, 0, 1},
    {"cfc_gelu", 30336, 3072, 768, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 1},
    {"cproj_bf16", 30336, 768, 3072, 0, 0, EPI_STORE, 1, 1, 1, 0, 1},
    {"dgrad_dgelu", 30336, 3072, 768, 0, 1, EPI_DGELU, 1, 1, 0, 0, 1},
    {"wgrad_split", 3072, 768, 30336, 1, 1, EPI_ATOMIC, 0, 8, 0, 0, 1},
    {"wgrad_sq", 768, 768, 30336, 1, 1, EPI_ATOMIC, 0, 16, 0, 0, 1},
    {"seg_remap", 25216, 768, 768, 0, 0, EPI_STORE, 0, 1, 1, 1, 0},
    {"tails_bf16", 1000, 520, 200, 0, 0, EPI_STORE, 1, 1, 1, 0, 0},
    {"gelu_tails", 300, 520, 192, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 0},
    {"dgelu_tails", 300, 520, 192, 0, 1, EPI_DGELU, 1, 1, 0, 0, 0},
    {"atomic_tails", 304, 520, 1000, 1, 1, EPI_ATOMIC, 0, 3, 0, 0, 0},
    {"n104_bf16", 72, 104, 96, 0, 0, EPI_STORE, 1, 1, 1, 0, 0},
    {"dgrad_fc", 30336, 768, 3072, 0, 1, EPI_STORE, 1, 1, 0, 0, 1},
    {"dgrad_in", 30336, 768, 2304, 0, 1, EPI_STORE, 1, 1, 0, 0, 1},
    {"wgrad_in", 2304, 768, 30336, 1, 1, EPI_ATOMIC, 0, 3, 0, 0, 1},
    // 22.. : round-2 eval path (residual stream through the epilogue, LayerNorm folded into the consumer)
    {"resid_tails", 300, 520, 192, 0, 0, EPI_RESID_LN, 0, 1, 1, 0, 0},
    {"resid_pair_tails", 1000, 520, 200, 0, 0, EPI_RESID_LN, 0, 1, 1, 0, 0},
    {"resid_n104", 72, 104, 96, 0, 0, EPI_RESID_LN, 0, 1, 1, 0, 0},
    {"fold_store_tails", 1000, 520, 200, 0, 0, EPI_STORE, 1, 1, 1, 0, 0, 1},
    {"fold_gelu_tails", 300, 520, 192, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 0, 1},
    {"resid_outproj", 30336, 768, 768, 0, 0, EPI_RESID_LN, 0, 1, 1, 0, 1},
    {"resid_cproj", 30336, 768, 3072, 0, 0, EPI_RESID_LN, 0, 1, 1, 0, 1},
    {"fold_inproj", 30336, 2304, 768, 0, 0, EPI_STORE, 1, 1, 1, 0, 1, 1},
    {"fold_cfc_eval", 30336, 3072, 768, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 1, 1, 1},
    {"sweep_outproj", 30336, 768, 768, 0, 0, EPI_STORE, 1, 1, 1, 0, 1},
    {"sweep_cproj", 30336, 768, 3072, 0, 0, EPI_STORE, 1, 1, 1, 0, 1},
    {"sweep_inproj", 30336, 2304, 768, 0, 0, EPI_STORE, 1, 1, 1, 0, 1},
    {"sweep_cfc_eval", 30336, 3072, 768, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 1, 0, 1},
    {"resid_nocopy", 1000, 520, 200, 0, 0, EPI_RESID_LN, 0, 1, 1, 0, 0, 0, 1},
    // 36.. : erf-GELU epilogues (BERT's activation), single-CTA and CTA-pair kernels
    {"erf_gelu_tails", 300, 520, 192, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 0, 0, 0, 1},
    {"erf_gelu_pair_tails", 1000, 520, 200, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 0, 0, 0, 1},
    {"erf_dgelu_tails", 300, 520, 192, 0, 1, EPI_DGELU, 1, 1, 0, 0, 0, 0, 0, 1},
    {"erf_dgelu_pair_tails", 1000, 520, 200, 0, 1, EPI_DGELU, 1, 1, 0, 0, 0, 0, 0, 1},
    {"erf_cfc", 16384, 3072, 768, 0, 0, EPI_QUICKGELU, 1, 1, 1, 0, 1, 0, 0, 1},
    {"erf_dgrad_dgelu", 16384, 3072, 768, 0, 1, EPI_DGELU, 1, 1, 0, 0, 1, 0, 0, 1},
};

__global__ void ref_gemm(const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int M, int N,
                         int K, int a_mn, int b_mn) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __bfloat162float(a_mn ? A[(long long)k * M + m] : A[(long long)m * K + k]);
    const float b = __bfloat162float(b_mn ? B[(long long)k * N + n] : B[(long long)n * K + k]);
    acc += a * b;
  }
  C[(long long)m * N + n] = acc;
}

static float frand(uint64_t& s) {
  s = s * 6364136223846793005ULL + 1442695040888963407ULL;
  return ((s >> 40) & 0xFFFF) / 32768.0f - 1.0f;
}
static float qgelu(float z) { return z / (1.f + expf(-1.702f * z)); }
static float qgelu_grad(float z) {
  float s = 1.f / (1.f + expf(-1.702f * z));
  return s * (1.f + 1.702f * z * (1.f - s));
}
static float egelu(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678f)); }
static float egelu_grad(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678f)) + z * 0.3989422804f * expf(-0.5f * z * z);
}
static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main(int argc, char** argv) {
  const int ncases = sizeof(kCases) / sizeof(kCases[0]);
  if (argc < 2) {
    printf("%d\n", ncases);
    return 0;
  }
  const int id = atoi(argv[1]);
  if (id < 0 || id >= ncases) return 1;
  Case c = kCases[id];
  // MMU_M=<rows> + MMU_TIMING_ONLY=1: time a case at another row count (e.g. the packed sweep's
  // 169 856 rows) without the host-side verification
  const bool timing_only = getenv("MMU_TIMING_ONLY") != nullptr;
  if (getenv("MMU_M") != nullptr && c.timed) c.M = atoi(getenv("MMU_M"));
  const int M = c.M, N = c.N, K = c.K;
  const int L = 237, lseg = 197;  // seg remap: rows (b, l<197) -> b*237 + l
  const long long out_rows = c.seg ? (long long)(M / lseg) * L : M;
  printf("case %d %s M=%d N=%d K=%d a_mn=%d b_mn=%d mode=%d bf16=%d splits=%d\n", id, c.name, M, N,
         K, c.a_mn, c.b_mn, c.mode, c.out_bf16, c.splits);

  uint64_t seed = 1234 + id;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
  const float scale = 1.0f / sqrtf((float)K) * 4.f;
  // timing-only runs fill 1 Mi random elements and tile them (the host generator is slow)
  auto fill = [&](auto& vec, auto gen) {
    const size_t n = vec.size(), blk = timing_only ? std::min<size_t>(n, 1u << 20) : n;
    for (size_t i = 0; i < blk; ++i) vec[i] = gen();
    for (size_t i = blk; i < n; ++i) vec[i] = vec[i - blk];
  };
  fill(hA, [&] { return __float2bfloat16_rn(frand(seed)); });
  fill(hB, [&] { return __float2bfloat16_rn(frand(seed) * scale); });
  std::vector<float> hbias(N), haux_f((size_t)out_rows * N);
  std::vector<__nv_bfloat16> haux_b((size_t)M * N);
  for (auto& v : hbias) v = frand(seed);
  fill(haux_f, [&] { return frand(seed); });
  fill(haux_b, [&] { return __float2bfloat16_rn(frand(seed) * 2.f); });

  __nv_bfloat16 *dA, *dB, *daux_b;
  float *dref, *dbias, *daux_f;
  void *dout, *dout2;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dref, (size_t)M * N * 4));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&daux_f, haux_f.size() * 4));
  CK(cudaMalloc(&daux_b, haux_b.size() * 2));
  CK(cudaMalloc(&dout, (size_t)out_rows * N * 4));
  CK(cudaMalloc(&dout2, (size_t)out_rows * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(daux_f, haux_f.data(), haux_f.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(daux_b, haux_b.data(), haux_b.size() * 2, cudaMemcpyHostToDevice));
  // LN fold: A holds raw rows; statistics of each row split over 3 partials, cw[n] = sum_k B[n][k]
  const int NT_IN = 3, NT_OUT = (N + 127) / 128;
  std::vector<float> hstats((size_t)M * NT_IN * 2, 0.f), hcw(N, 0.f), hmean(M), hrstd(M);
  if (c.fold) {
    for (int m = 0; m < (timing_only ? std::min(M, 4096) : M); ++m) {
      double s1 = 0, s2 = 0;
      for (int k = 0; k < K; ++k) {
        const float x = __bfloat162float(hA[(size_t)m * K + k]);
        const int part = k * NT_IN / K;
        hstats[((size_t)m * NT_IN + part) * 2] += x;
        hstats[((size_t)m * NT_IN + part) * 2 + 1] += x * x;
        s1 += x;
        s2 += (double)x * x;
      }
      const double mean = s1 / K, var = s2 / K - mean * mean;
      hmean[m] = (float)mean;
      hrstd[m] = (float)(1.0 / sqrt(var + 1e-5));
    }
    for (int n = 0; n < N; ++n) {
      double t = 0;
      for (int k = 0; k < K; ++k) t += __bfloat162float(hB[(size_t)n * K + k]);
      hcw[n] = (float)t;
    }
  }
  float *dstats_in, *dcw, *dstats_out;
  CK(cudaMalloc(&dstats_in, hstats.size() * 4));
  CK(cudaMalloc(&dcw, (size_t)N * 4));
  CK(cudaMalloc(&dstats_out, (size_t)M * NT_OUT * 2 * 4));
  CK(cudaMemcpy(dstats_in, hstats.data(), hstats.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dcw, hcw.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dstats_out, 0, (size_t)M * NT_OUT * 2 * 4));
  CK(cudaMemset(dout, 0, (size_t)out_rows * N * 4));
  CK(cudaMemset(dout2, 0, (size_t)out_rows * N * 4));

  if (!timing_only) ref_gemm<<<dim3((N + 127) / 128, M), 128>>>(dA, dB, dref, M, N, K, c.a_mn, c.b_mn);
  CK(cudaDeviceSynchronize());
  std::vector<float> href((size_t)M * N);
  CK(cudaMemcpy(href.data(), dref, href.size() * 4, cudaMemcpyDeviceToHost));

  GemmProblem p{M, N, K, c.a_mn, c.b_mn, c.splits};
  GemmEpilogue e{};
  e.mode = c.mode;
  e.out_bf16 = c.out_bf16;
  e.out = dout;
  e.out2 = dout2;
  e.bias = c.bias ? dbias : nullptr;
  e.aux = (c.mode == EPI_RESIDUAL || c.mode == EPI_RESID_LN) ? (const void*)daux_f : (const void*)daux_b;
  if (c.no_out && c.mode == EPI_QUICKGELU) e.out = nullptr;
  if (c.no_out && c.mode == EPI_RESID_LN) e.out2 = nullptr;
  if (c.mode == EPI_RESID_LN) {
    e.stats_out = dstats_out;
    e.stats_nt = NT_OUT;
  }
  if (c.fold) {
    e.ln_stats = dstats_in;
    e.ln_cw = dcw;
    e.ln_nt = NT_IN;
    e.ln_inv_d = 1.0f / K;
    e.ln_eps = 1e-5f;
  }
  e.ld_out = e.ld_out2 = e.ld_aux = N;
  e.seg_len = c.seg ? lseg : 0;
  e.seg_stride = L;
  e.seg_off = 0;
  e.alpha = 1.0f;
  e.act = c.act;
  const long long lda = c.a_mn ? M : K, ldb = c.b_mn ? N : K;

  int rc = gemm_bf16_launch(dA, lda, dB, ldb, p, e, 0);
  if (rc != 0) {
    printf("RESULT %s LAUNCH_ERROR rc=%d\n", c.name, rc);
    return 3;
  }
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    printf("RESULT %s KERNEL_ERROR %s\n", c.name, cudaGetErrorString(err));
    return 4;
  }

  // ---- verify
  std::vector<float> got((size_t)out_rows * N), got2;
  std::vector<float> gstats;
  if (c.mode == EPI_RESID_LN) {
    gstats.resize((size_t)M * NT_OUT * 2);
    CK(cudaMemcpy(gstats.data(), dstats_out, gstats.size() * 4, cudaMemcpyDeviceToHost));
    if (!c.no_out) {
      std::vector<__nv_bfloat16> tmp((size_t)out_rows * N);
      got2.resize(tmp.size());
      CK(cudaMemcpy(tmp.data(), dout2, tmp.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < tmp.size(); ++i) got2[i] = __bfloat162float(tmp[i]);
    }
  }
  if (c.out_bf16 && c.no_out && c.mode == EPI_QUICKGELU) {
    std::vector<__nv_bfloat16> tmp((size_t)out_rows * N);
    got2.resize(tmp.size());
    CK(cudaMemcpy(tmp.data(), dout2, tmp.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) got2[i] = __bfloat162float(tmp[i]);
  } else if (c.out_bf16) {
    std::vector<__nv_bfloat16> tmp((size_t)out_rows * N);
    CK(cudaMemcpy(tmp.data(), dout, tmp.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) got[i] = __bfloat162float(tmp[i]);
    if (c.mode == EPI_QUICKGELU) {
      got2.resize(tmp.size());
      CK(cudaMemcpy(tmp.data(), dout2, tmp.size() * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < tmp.size(); ++i) got2[i] = __bfloat162float(tmp[i]);
    }
  } else {
    CK(cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost));
  }
  long long bad = 0;
  double maxerr = 0;
  int shown = 0;
  for (int m = 0; m < (timing_only ? 0 : M); ++m) {
    const long long orow = c.seg ? (long long)(m / lseg) * L + m % lseg : m;
    for (int n = 0; n < N; ++n) {
      float ref = href[(size_t)m * N + n] + (c.bias ? hbias[n] : 0.f);
      if (c.fold) ref = hrstd[m] * (href[(size_t)m * N + n] - hmean[m] * hcw[n]) + (c.bias ? hbias[n] : 0.f);
      float ref2 = 0.f;
      switch (c.mode) {
        case EPI_QUICKGELU: ref2 = c.act ? egelu(ref) : qgelu(ref); break;
        case EPI_RESIDUAL:
        case EPI_RESID_LN: ref += haux_f[(size_t)orow * N + n]; break;
        case EPI_DGELU: {
          const float z = __bfloat162float(haux_b[(size_t)m * N + n]);
          ref *= c.act ? egelu_grad(z) : qgelu_grad(z);
          break;
        }
        default: break;
      }
      float g = got[(size_t)orow * N + n];
      const float tol = c.out_bf16 ? (2e-2f + 1e-2f * fabsf(ref)) : (1e-3f + 1e-3f * fabsf(ref));
      if (c.mode == EPI_QUICKGELU && c.no_out) g = ref;  // eval mode: z is not written
      float d = fabsf(g - ref);
      if (c.mode == EPI_QUICKGELU) d = fmaxf(d, fabsf(got2[(size_t)orow * N + n] - ref2));
      if (c.mode == EPI_RESID_LN && !c.no_out && got2[(size_t)orow * N + n] != bf16r(g)) d = 1e9f;  // exact copy
      if (d > maxerr) maxerr = d;
      if (!(d <= tol)) {
        ++bad;
        if (shown < 12) {
          printf("  mismatch m=%d n=%d got=%f ref=%f\n", m, n, g, ref);
          ++shown;
        }
      }
    }
  }
  if (c.mode == EPI_RESID_LN && !timing_only) {  // per-slab row sums of what was written
    for (int m = 0; m < M; ++m)
      for (int t = 0; t < NT_OUT; ++t) {
        double s1 = 0, s2 = 0;
        for (int n = t * 128; n < N && n < (t + 1) * 128; ++n) {
          const double v = got[(size_t)m * N + n];
          s1 += v;
          s2 += v * v;
        }
        const float g1 = gstats[((size_t)m * NT_OUT + t) * 2], g2 = gstats[((size_t)m * NT_OUT + t) * 2 + 1];
        if (!(fabs(g1 - s1) <= 1e-3 + 1e-4 * fabs(s2)) || !(fabs(g2 - s2) <= 1e-3 + 1e-4 * fabs(s2))) {
          ++bad;
          if (shown < 12) {
            printf("  stats mismatch m=%d slab=%d got=(%f, %f) ref=(%f, %f)\n", m, t, g1, g2, s1, s2);
            ++shown;
          }
        }
      }
  }
  printf("RESULT %s %s bad=%lld/%lld maxerr=%.3e\n", c.name, timing_only ? "UNVERIFIED" : bad == 0 ? "PASS" : "FAIL", bad,
         (long long)M * N, maxerr);

  if (c.timed && (bad == 0 || getenv("MMU_FORCE_TIMING") != nullptr)) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 20;
    for (int i = 0; i < 3; ++i) gemm_bf16_launch(dA, lda, dB, ldb, p, e, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) gemm_bf16_launch(dA, lda, dB, ldb, p, e, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("TIMING %s M=%d %.3f ms  %.1f TFLOP/s\n", c.name, M, ms, 2.0 * M * N * K / ms / 1e9);
  }
  return (bad == 0 || getenv("MMU_FORCE_TIMING") != nullptr) ? 0 : 5;
}
