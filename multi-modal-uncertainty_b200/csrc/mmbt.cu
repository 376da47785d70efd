// MMBT engine implementation.  See mmbt.h.
//
// Data layout in HBM: row r = b*S + j of every (B, S, .) activation, S = number of sequence
// positions that enter the encoder (all n_img + 2 + S_txt of them, or the subset of
// forward_img_only / forward_txt_only / forward_control).  Activations and GEMM operands are fp32
// (parity path) or bf16 (tensor cores, fp32 accumulation); the pre-LayerNorm sums s1 / s2 and the
// gradient stream stay fp32.  Per layer (pytorch_pretrained_bert BertLayer, post-LN):
//   qkv = h W_qkv^T + b          (query | key | value packed: their weights are contiguous in the
//                                 flat buffer, one [3D, D] GEMM)
//   ctx = softmax(q k^T / sqrt(hd) + mask) v            (sequence-axis attention, attention.cu)
//   a   = LN(h + ctx W_o^T + b_o)                       (GEMM, then add + LayerNorm in one row pass)
//   u   = gelu_erf(a W_i^T + b_i) ; h' = LN(a + u W_o2^T + b_o2)
// Backward mirrors it; LN'(branch + residual) is one kernel (postln_bwd), so the sum of the two
// gradient paths of a post-LN block is never written to HBM.
#include "mmbt.h"

#include <cuda_bf16.h>

#include <cstdio>
#include <cstring>

#include "common.h"
#include "gemm_api.h"
#include "kernels.h"

namespace mmu {

namespace {

#define MB_TRY(x)               \
  do {                          \
    const int rc__ = (x);       \
    if (rc__ != 0) return rc__; \
  } while (0)
#define MB_CHECK_LAUNCH()                                                       \
  do {                                                                          \
    const cudaError_t err__ = cudaGetLastError();                               \
    if (err__ != cudaSuccess) {                                                 \
      fprintf(stderr, "mmu: launch failed at %s:%d: %s\n", __FILE__, __LINE__, \
              cudaGetErrorString(err__));                                       \
      return MMU_ERR_CUDA;                                                      \
    }                                                                           \
    count_launch();                                                             \
  } while (0)

constexpr long long ALIGN_ELEMS = 64;
constexpr float BERT_LN_EPS = 1e-12f;
long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------- parameter layout
struct LayerP {
  long long q_w, k_w, v_w, q_b, k_b, v_b, ao_w, ao_b, ln1_w, ln1_b, i_w, i_b, o_w, o_b, ln2_w, ln2_b;
};
struct Layout {
  long long word, pos, type, eln_w, eln_b, img_w, img_b;
  LayerP layer[48];
  long long pool_w, pool_b, clf_w, clf_b;
  long long total;
};

struct TableBuilder {
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
      e.offset = off;
      e.numel = numel;
      e.rows = rows;
      e.cols = cols;
      e.stage = 0;
    }
    ++n;
    cursor = align_up(cursor + numel, ALIGN_ELEMS);
    return off;
  }
};

int check_config_drop(const MmbtConfig& c) {
  for (float p : {c.drop_hidden, c.drop_attn, c.drop_img})
    if (!(p >= 0.f && p < 1.f)) return MMU_ERR_ARG;
  return 0;
}

int check_config(const MmbtConfig& c) {
  if (int rc = check_config_drop(c)) return rc;
  if (c.B < 1 || c.S_txt < 0 || c.n_img < 1 || c.d_img < 8 || c.D < 64 || c.D % 64 != 0 || c.D > 1024)
    return MMU_ERR_SHAPE;
  if (c.n_head < 1 || c.D % c.n_head != 0 || c.n_layers < 1 || c.n_layers > 48) return MMU_ERR_SHAPE;
  if (c.d_ff < 8 || c.d_ff % 8 != 0 || c.d_img % 8 != 0) return MMU_ERR_SHAPE;
  if (c.vocab < 1 || c.n_types < 1 || c.n_types > 4 || c.C < 1) return MMU_ERR_SHAPE;
  if (c.max_pos < c.n_img + 2 || c.max_pos < c.S_txt) return MMU_ERR_SHAPE;
  if (c.cls_id < 0 || c.cls_id >= c.vocab || c.sep_id < 0 || c.sep_id >= c.vocab) return MMU_ERR_ARG;
  if (c.precision != PREC_FP32 && c.precision != PREC_BF16) return MMU_ERR_ARG;
  if (c.precision == PREC_BF16 && (c.D / c.n_head) % 64 != 0) return MMU_ERR_SHAPE;
  if (c.max_seq < 0) return MMU_ERR_SHAPE;
  return 0;
}

// Order = the reference module's named_parameters() order (src/mmbt.py:86-96,238-243) for the
// tensors this engine owns; q/k/v weights (and biases) are adjacent so that they form one
// [3D, D] operand.  D % 64 == 0 keeps them gap-free under the 64-element alignment.
int build_layout(const MmbtConfig& c, Layout* L, ParamEntry* out, int max_entries) {
  TableBuilder tb{out, max_entries, 0, 0};
  char nm[96];
  L->word = tb.add("enc.txt_embeddings.word_embeddings.weight", c.vocab, c.D);
  L->pos = tb.add("enc.txt_embeddings.position_embeddings.weight", c.max_pos, c.D);
  L->type = tb.add("enc.txt_embeddings.token_type_embeddings.weight", c.n_types, c.D);
  L->eln_w = tb.add("enc.txt_embeddings.LayerNorm.weight", c.D, 0);
  L->eln_b = tb.add("enc.txt_embeddings.LayerNorm.bias", c.D, 0);
  L->img_w = tb.add("enc.img_embeddings.img_embeddings.weight", c.D, c.d_img);
  L->img_b = tb.add("enc.img_embeddings.img_embeddings.bias", c.D, 0);
  for (int i = 0; i < c.n_layers; ++i) {
    LayerP& p = L->layer[i];
    auto name = [&](const char* suffix) {
      std::snprintf(nm, sizeof(nm), "enc.encoder.layer.%d.%s", i, suffix);
      return nm;
    };
    // weights first (adjacent), then the biases (adjacent); the Python side registers them in
    // the reference's order (query.weight, query.bias, key.weight, ...)
    p.q_w = tb.add(name("attention.self.query.weight"), c.D, c.D);
    p.k_w = tb.add(name("attention.self.key.weight"), c.D, c.D);
    p.v_w = tb.add(name("attention.self.value.weight"), c.D, c.D);
    p.q_b = tb.add(name("attention.self.query.bias"), c.D, 0);
    p.k_b = tb.add(name("attention.self.key.bias"), c.D, 0);
    p.v_b = tb.add(name("attention.self.value.bias"), c.D, 0);
    p.ao_w = tb.add(name("attention.output.dense.weight"), c.D, c.D);
    p.ao_b = tb.add(name("attention.output.dense.bias"), c.D, 0);
    p.ln1_w = tb.add(name("attention.output.LayerNorm.weight"), c.D, 0);
    p.ln1_b = tb.add(name("attention.output.LayerNorm.bias"), c.D, 0);
    p.i_w = tb.add(name("intermediate.dense.weight"), c.d_ff, c.D);
    p.i_b = tb.add(name("intermediate.dense.bias"), c.d_ff, 0);
    p.o_w = tb.add(name("output.dense.weight"), c.D, c.d_ff);
    p.o_b = tb.add(name("output.dense.bias"), c.D, 0);
    p.ln2_w = tb.add(name("output.LayerNorm.weight"), c.D, 0);
    p.ln2_b = tb.add(name("output.LayerNorm.bias"), c.D, 0);
  }
  L->pool_w = tb.add("enc.pooler.dense.weight", c.D, c.D);
  L->pool_b = tb.add("enc.pooler.dense.bias", c.D, 0);
  L->clf_w = tb.add("clf.weight", c.C, c.D);
  L->clf_b = tb.add("clf.bias", c.C, 0);
  L->total = tb.cursor;
  return tb.n;
}

// ------------------------------------------------------------------------- workspace
struct LayerWs {
  void* qkv;     // act [M, 3D]
  void* probs;   // bf16 [G, S, Sp] / fp32 [G, S, S]
  void* pdrop;   // bf16 training with attention dropout: the dropped probabilities (dV = Pd^T dO)
  void* ctx;     // act [M, D]
  float* s1;     // fp32 [M, D]  h + attention branch (pre-LN)
  float* st1;    // mean | rstd
  void* a;       // act [M, D]
  void* z;       // act [M, F]
  void* u;       // act [M, F]
  float* s2;     // fp32 [M, D]
  float* st2;
  void* h;       // act [M, D]  block output
};
struct Ws {
  void* params_lp;
  void* img_lp;     // act [B*n_img, d_img] (bf16 path: cast of the fp32 image tokens)
  float* imgp;      // fp32 [B*n_img, D] projected image tokens
  float* addmask;   // fp32 [B, S]
  int* row_word;    // int32 [M] word id or -1
  int* row_pos;
  int* row_type;
  int* row_img;     // b*n_img + slot, or -1
  int* row_side;    // 1: image side of the sequence ([CLS] img.. [SEP]), 0: text
  float* s0;        // fp32 [M, D] embedding sum (pre-LN)
  float* st0;
  void* h0;         // act [M, D]
  LayerWs layer[48];
  void* ybuf;       // act [M, D] branch output of the closing projections
  float* scores;    // fp32 [G, S, Sp] scratch
  float* pooled;    // fp32 [B, D]
  // backward
  float* gA;        // fp32 [M, D]
  float* gB;
  void* g_lp;       // act [M, D]
  void* dbig;       // act [M, max(F, 3D)]
  void* dh;         // act [M, D]
  void* dprobs;     // bf16 [G, S, Sp]
  float* dpooled;   // fp32 [B, D] x2
  float* dimgp;     // fp32 [B*n_img, D]
  void* dimgp_lp;   // act copy
  long long bytes;
};

struct Bump {
  char* base;
  long long off;
  template <typename T>
  T* take(long long bytes) {
    const long long o = off;
    off = align_up(off + bytes, 256);
    return base != nullptr ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};

void carve(const MmbtConfig& c, int training, void* base, const Layout& lay, Ws* w) {
  Bump b{static_cast<char*>(base), 0};
  const bool bf = c.precision == PREC_BF16;
  const long long s = bf ? 2 : 4;
  const long long S = c.max_seq > 0 ? c.max_seq : c.n_img + 2 + c.S_txt;  // capacity
  const long long M = static_cast<long long>(c.B) * S, D = c.D, F = c.d_ff;
  const long long G = static_cast<long long>(c.B) * c.n_head, Sp = (S + 7) / 8 * 8;
  const long long sq = bf ? G * S * Sp : G * S * S;
  w->params_lp = bf ? b.take<void>(lay.total * 2) : nullptr;
  w->img_lp = bf ? b.take<void>(static_cast<long long>(c.B) * c.n_img * c.d_img * 2) : nullptr;
  w->imgp = b.take<float>(static_cast<long long>(c.B) * c.n_img * D * 4);
  w->addmask = b.take<float>(M * 4);
  w->row_word = b.take<int>(M * 4);
  w->row_pos = b.take<int>(M * 4);
  w->row_type = b.take<int>(M * 4);
  w->row_img = b.take<int>(M * 4);
  w->row_side = b.take<int>(M * 4);
  w->s0 = b.take<float>(M * D * 4);
  w->st0 = b.take<float>(2 * M * 4);
  w->h0 = b.take<void>(M * D * s);
  const int slots = training ? c.n_layers : 2;
  for (int i = 0; i < slots && i < c.n_layers; ++i) {
    LayerWs& l = w->layer[i];
    const bool shared = !training && i == 1;  // eval: everything but the output is shared
    if (shared) {
      l = w->layer[0];
      l.h = b.take<void>(M * D * s);
      continue;
    }
    l.qkv = b.take<void>(M * 3 * D * s);
    l.probs = b.take<void>(sq * (bf ? 2 : 4));
    l.pdrop = (training && bf && c.drop_attn > 0.f) ? b.take<void>(sq * 2) : nullptr;
    l.ctx = b.take<void>(M * D * s);
    l.s1 = b.take<float>(M * D * 4);
    l.st1 = b.take<float>(2 * M * 4);
    l.a = b.take<void>(M * D * s);
    l.z = b.take<void>(M * F * s);
    l.u = b.take<void>(M * F * s);
    l.s2 = b.take<float>(M * D * 4);
    l.st2 = b.take<float>(2 * M * 4);
    l.h = b.take<void>(M * D * s);
  }
  for (int i = slots; i < c.n_layers; ++i) w->layer[i] = w->layer[i & 1];
  w->ybuf = b.take<void>(M * D * s);
  w->scores = b.take<float>((bf ? G * S * Sp : G * S * S) * 4);
  w->pooled = b.take<float>(static_cast<long long>(c.B) * D * 4);
  if (training) {
    w->gA = b.take<float>(M * D * 4);
    w->gB = b.take<float>(M * D * 4);
    // activation-dtype copy of the gradient stream; with hidden dropout the copy differs from the
    // stream itself (masked), so the fp32 path needs its own buffer too
    w->g_lp = (bf || c.drop_hidden > 0.f) ? b.take<void>(M * D * s) : nullptr;
    w->dbig = b.take<void>(M * (F > 3 * D ? F : 3 * D) * s);
    w->dh = b.take<void>(M * D * s);
    w->dprobs = bf ? b.take<void>(sq * 2) : nullptr;
    w->dpooled = b.take<float>(2LL * c.B * D * 4);
    w->dimgp = b.take<float>(static_cast<long long>(c.B) * c.n_img * D * 4);
    w->dimgp_lp = bf ? b.take<void>(static_cast<long long>(c.B) * c.n_img * D * 2) : nullptr;
  } else {
    w->gA = w->gB = nullptr; w->g_lp = nullptr; w->dbig = nullptr; w->dh = nullptr;
    w->dprobs = nullptr; w->dpooled = nullptr; w->dimgp = nullptr; w->dimgp_lp = nullptr;
  }
  w->bytes = b.off;
}

// ------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Embeddings of one row (b, j) of the encoder input (src/mmbt.py:47-84 ImageBertEmbeddings for
// positions < n_img + 2, BertEmbeddings for the text positions), their sum s0 (fp32, kept for the
// backward), LayerNorm (eps 1e-12) -> h0, the additive attention mask and the row's bookkeeping.
template <typename T>
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const long long* __restrict__ txt, const long long* __restrict__ mask,
                 const long long* __restrict__ segment, const float* __restrict__ imgp,
                 const int* __restrict__ indices, int idx_stride, const float* __restrict__ word,
                 const float* __restrict__ pos, const float* __restrict__ type,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int B, int S,
                 int S_txt, int n_img, int D, int cls_id, int sep_id, int vocab, int n_types,
                 float* __restrict__ s0, float* __restrict__ st0, T* __restrict__ h0,
                 float* __restrict__ addmask, int* __restrict__ row_word, int* __restrict__ row_pos,
                 int* __restrict__ row_type, int* __restrict__ row_img, int* __restrict__ row_side,
                 const dropout::Site drop_img, const dropout::Site drop_txt) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int M = B * S;
  if (r >= M) return;
  const int b = r / S, j = r % S;
  const int p = indices != nullptr ? indices[static_cast<long long>(b) * idx_stride + j] : j;
  const int n2 = n_img + 2;
  int wid = -1, pid, tid = 0, iid = -1;
  float m = 1.f;
  if (p < n2) {
    pid = p;
    if (p == 0) wid = cls_id;
    else if (p == n2 - 1) wid = sep_id;
    else iid = b * n_img + p - 1;
  } else {
    const int t = p - n2;
    long long w = txt[static_cast<long long>(b) * S_txt + t];
    w = w < 0 ? 0 : (w >= vocab ? vocab - 1 : w);
    wid = static_cast<int>(w);
    pid = t;
    long long sg = segment[static_cast<long long>(b) * S_txt + t];
    tid = static_cast<int>(sg < 0 ? 0 : (sg >= n_types ? n_types - 1 : sg));
    m = static_cast<float>(mask[static_cast<long long>(b) * S_txt + t]);
  }
  const float* tok = wid >= 0 ? word + static_cast<long long>(wid) * D
                              : imgp + static_cast<long long>(iid) * D;
  const float* pe = pos + static_cast<long long>(pid) * D;
  const float* te = type + static_cast<long long>(tid) * D;
  float v[32];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < D ? tok[c] + pe[c] + te[c] : 0.f;
    sum += v[i];
  }
  const float mean = wsum(sum) / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < D) ss += (v[i] - mean) * (v[i] - mean);
  }
  const float rstd = rsqrtf(wsum(ss) / D + BERT_LN_EPS);
  const long long ro = static_cast<long long>(r) * D;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      if (s0 != nullptr) s0[ro + c] = v[i];
      float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
      // embedding dropout after the LayerNorm: ImageBertEmbeddings.dropout (args.dropout,
      // src/mmbt.py:56,82) on the image side, BertEmbeddings.dropout on the text side
      const dropout::Site& ds = p < n2 ? drop_img : drop_txt;
      if (ds.on()) o *= ds.mult(static_cast<unsigned int>(r) * D + c);
      stf(h0 + ro + c, o);
    }
  }
  if (lane == 0) {
    if (st0 != nullptr) {
      st0[r] = mean;
      st0[M + r] = rstd;
    }
    addmask[r] = (1.0f - m) * -10000.0f;
    row_word[r] = wid;
    row_pos[r] = pid;
    row_type[r] = tid;
    row_img[r] = iid;
    row_side[r] = p < n2 ? 1 : 0;
  }
}

// dE (fp32 [M, D]) scattered to the embedding tables (atomics) and to the projected image tokens.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ dE, const int* __restrict__ row_word,
                 const int* __restrict__ row_pos, const int* __restrict__ row_img, int M, int D,
                 float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dimgp) {
  const long long n = static_cast<long long>(M) * D;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / D), c = static_cast<int>(i % D);
    const float g = dE[i];
    const int wid = row_word[r];
    // nn.Embedding(padding_idx=0) of BertEmbeddings: the gradient of word row 0 is dropped
    if (wid > 0) atomicAdd(dword + static_cast<long long>(wid) * D + c, g);
    else if (wid < 0) dimgp[static_cast<long long>(row_img[r]) * D + c] = g;
    atomicAdd(dpos + static_cast<long long>(row_pos[r]) * D + c, g);
  }
}
// token-type table: few rows, so each thread reduces a slab of rows per column first
__global__ void __launch_bounds__(256)
type_bwd_kernel(const float* __restrict__ dE, const int* __restrict__ row_type, int M, int D,
                int n_types, int slab, float* __restrict__ dtype) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const int r0 = blockIdx.y * slab, r1 = min(M, r0 + slab);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; ++r) {
    const int t = row_type[r];
    const float g = dE[static_cast<long long>(r) * D + c];
    acc[0] += t == 0 ? g : 0.f;
    acc[1] += t == 1 ? g : 0.f;
    acc[2] += t == 2 ? g : 0.f;
    acc[3] += t == 3 ? g : 0.f;
  }
  for (int t = 0; t < n_types; ++t) atomicAdd(dtype + static_cast<long long>(t) * D + c, acc[t]);
}

// erf-GELU (pytorch_pretrained_bert.modeling.gelu: x * 0.5 * (1 + erf(x / sqrt(2))))
__device__ __forceinline__ float gelu_erf(float z) { return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float z) {
  return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
}
template <typename T>
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const T* __restrict__ z, T* __restrict__ u, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    stf(u + i, gelu_erf(ldf(z + i)));
}
template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const T* __restrict__ z, T* __restrict__ du, long long n) {  // in place: du -> dz
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    stf(du + i, ldf(du + i) * gelu_erf_grad(ldf(z + i)));
}

// Small dense layers on B rows (pooler 768x768 + tanh on the first token, classifier): fp32
// weights, one warp per output element.  x row b starts at x + b * x_stride.
template <typename T>
__global__ void __launch_bounds__(256)
rows_linear_fwd_kernel(const T* __restrict__ x, long long x_stride, const float* __restrict__ W,
                       const float* __restrict__ bias, float* __restrict__ out, int B, int N, int K,
                       int act_tanh) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= B * N) return;
  const int b = o / N, n = o % N;
  const T* xr = x + b * x_stride;
  const float* wr = W + static_cast<long long>(n) * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(ldf(xr + k), wr[k], acc);
  acc = wsum(acc);
  if (lane == 0) {
    acc += bias[n];
    out[o] = act_tanh ? tanhf(acc) : acc;
  }
}
// dW[n, k] += sum_b dy[b, n] x[b, k] ; db[n] += sum_b dy[b, n]   (dy already includes tanh')
template <typename T>
__global__ void __launch_bounds__(256)
rows_linear_bwd_w_kernel(const float* __restrict__ dy, const T* __restrict__ x, long long x_stride,
                         float* __restrict__ dW, float* __restrict__ db, int B, int N, int K) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(N) * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = dy[b * N + n];
    acc = fmaf(g, ldf(x + b * x_stride + k), acc);
    accb += g;
  }
  dW[i] += acc;
  if (k == 0) db[n] += accb;
}
// dx[b, k] = sum_n dy[b, n] W[n, k], written to dx + b * dx_stride
__global__ void __launch_bounds__(256)
rows_linear_bwd_x_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                         float* __restrict__ dx, long long dx_stride, int B, int N, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i % K;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(dy[b * N + n], W[static_cast<long long>(n) * K + k], acc);
  dx[b * dx_stride + k] = acc;
}
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dy[i] * (1.0f - y[i] * y[i]);
}

int grid1d(long long n, int per_block) {
  long long g = (n + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

struct Gemm {
  int prec;
  cudaStream_t stream;
  int operator()(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn,
                 int M, int N, int K, GemmEpilogue e, int splits = 1) const {
    GemmProblem p{M, N, K, a_mn, b_mn, splits};
    e.out_bf16 = prec == PREC_BF16 ? e.out_bf16 : 0;
    if (prec == PREC_BF16) return gemm_bf16_launch(A, lda, B, ldb, p, e, stream);
    return gemm_f32_launch(static_cast<const float*>(A), lda, static_cast<const float*>(B), ldb, p,
                           e, stream);
  }
};
GemmEpilogue epi(int mode, void* out, int out_bf16, long long ld_out, const float* bias) {
  GemmEpilogue e{};
  e.mode = mode; e.out_bf16 = out_bf16; e.out = out; e.ld_out = ld_out; e.bias = bias;
  e.alpha = 1.0f;
  return e;
}
// split-K factor of a weight-gradient GEMM (see engine.cu wgrad_splits)
int wsplits(int Mg, int Ng, int K) {
  const bool pair = Mg >= 512;
  const int tiles = ((Mg + (pair ? 255 : 127)) / (pair ? 256 : 128)) * ((Ng + 255) / 256);
  const int units = pair ? sm_count() / 2 : sm_count();
  int smax = ((K + 63) / 64) / 4;
  if (smax > 16) smax = 16;
  if (smax < 1) smax = 1;
  int best_s = 1;
  double best = 0.0;
  for (int sp = 1; sp <= smax; ++sp) {
    const int t = tiles * sp;
    const double eff = static_cast<double>(t) / (static_cast<double>((t + units - 1) / units) * units);
    if (eff > best + 0.02) { best = eff; best_s = sp; }
  }
  return best_s;
}

int resolve_S(const MmbtConfig& c, const MmbtInputs& in) {
  const int full = c.n_img + 2 + c.S_txt;
  const int cap = c.max_seq > 0 ? c.max_seq : full;
  if (in.indices == nullptr) return full <= cap ? full : MMU_ERR_WORKSPACE;
  if (in.n_sel < 1 || in.n_sel > full) return MMU_ERR_SHAPE;
  return in.n_sel <= cap ? in.n_sel : MMU_ERR_WORKSPACE;
}

}  // namespace

// =========================================================================== public
int mmbt_param_table(const MmbtConfig& c, ParamEntry* out, int max_entries) {
  if (int rc = check_config(c)) return rc;
  Layout lay;
  return build_layout(c, &lay, out, max_entries);
}
long long mmbt_param_count(const MmbtConfig& c) {
  if (int rc = check_config(c)) return rc;
  Layout lay;
  build_layout(c, &lay, nullptr, 0);
  return lay.total;
}
long long mmbt_workspace_bytes(const MmbtConfig& c, int training) {
  if (int rc = check_config(c)) return rc;
  Layout lay;
  build_layout(c, &lay, nullptr, 0);
  Ws w;
  carve(c, training, nullptr, lay, &w);
  return w.bytes;
}

int mmbt_forward(const MmbtConfig& c, const float* params, const MmbtInputs& in, void* ws,
                 long long ws_bytes, int training, float* logits, cudaStream_t stream) {
  MB_TRY(check_config(c));
  if (params == nullptr || ws == nullptr || logits == nullptr || in.img == nullptr) return MMU_ERR_ARG;
  if (c.S_txt > 0 && (in.txt == nullptr || in.mask == nullptr || in.segment == nullptr)) return MMU_ERR_ARG;
  Layout lay;
  build_layout(c, &lay, nullptr, 0);
  Ws w;
  carve(c, training, ws, lay, &w);
  if (w.bytes > ws_bytes) return MMU_ERR_WORKSPACE;
  const int S = resolve_S(c, in);
  if (S < 0) return S;
  const bool bf = c.precision == PREC_BF16;
  const int dt = bf ? DT_BF16 : DT_F32;
  const int D = c.D, F = c.d_ff, M = c.B * S, Mi = c.B * c.n_img;
  const Gemm gemm{c.precision, stream};
  const void* shadow = in.params_bf16 != nullptr ? in.params_bf16 : w.params_lp;
  auto W = [&](long long off) -> const void* {
    return bf ? static_cast<const void*>(static_cast<const uint16_t*>(shadow) + off)
              : static_cast<const void*>(params + off);
  };
  if (bf && in.params_bf16 == nullptr)
    MB_TRY(cast_f32_to_bf16(params, w.params_lp, static_cast<size_t>(lay.total), stream));

  // ---- image tokens -> hidden size (src/mmbt.py:68: self.img_embeddings(input_imgs))
  const void* img_op = in.img;
  if (bf) {
    MB_TRY(cast_f32_to_bf16(in.img, w.img_lp, static_cast<size_t>(Mi) * c.d_img, stream));
    img_op = w.img_lp;
  }
  MB_TRY(gemm(img_op, c.d_img, 0, W(lay.img_w), c.d_img, 0, Mi, D, c.d_img,
              epi(EPI_STORE, w.imgp, 0, D, params + lay.img_b)));
  // ---- dropout sites of this forward (training only; csrc/dropout.cuh).  0: embeddings, per layer
  //      l: 4l+1 attention probabilities, 4l+2 attention-output dense, 4l+3 FFN-output dense.
  const dropout::Site off{0u, 0u, 0u, 1.0f};
  auto site = [&](float p, int id) { return (training && p > 0.f) ? dropout::make_site(p, in.drop_seed, id) : off; };
  // ---- embeddings + LayerNorm + mask + bookkeeping, only for the selected positions
  {
    const int grid = (M + 7) / 8;
    float* s0 = training ? w.s0 : nullptr;
    float* st0 = training ? w.st0 : nullptr;
    if (bf)
      embed_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
          in.txt, in.mask, in.segment, w.imgp, in.indices, in.indices_per_sample ? S : 0, params + lay.word,
          params + lay.pos, params + lay.type, params + lay.eln_w, params + lay.eln_b, c.B, S, c.S_txt, c.n_img, D,
          c.cls_id, c.sep_id, c.vocab, c.n_types, s0, st0, static_cast<__nv_bfloat16*>(w.h0),
          w.addmask, w.row_word, w.row_pos, w.row_type, w.row_img, w.row_side, site(c.drop_img, 0),
          site(c.drop_hidden, 0));
    else
      embed_fwd_kernel<float><<<grid, 256, 0, stream>>>(
          in.txt, in.mask, in.segment, w.imgp, in.indices, in.indices_per_sample ? S : 0, params + lay.word,
          params + lay.pos, params + lay.type, params + lay.eln_w, params + lay.eln_b, c.B, S, c.S_txt, c.n_img, D,
          c.cls_id, c.sep_id, c.vocab, c.n_types, s0, st0, static_cast<float*>(w.h0), w.addmask,
          w.row_word, w.row_pos, w.row_type, w.row_img, w.row_side, site(c.drop_img, 0),
          site(c.drop_hidden, 0));
    MB_CHECK_LAUNCH();
  }
  // ---- encoder
  const void* h = w.h0;
  for (int i = 0; i < c.n_layers; ++i) {
    const LayerP& p = lay.layer[i];
    const LayerWs& l = w.layer[i];
    MB_TRY(gemm(h, D, 0, W(p.q_w), D, 0, M, 3 * D, D, epi(EPI_STORE, l.qkv, bf, 3 * D, params + p.q_b)));
    // attention-probability dropout: the dropped copy that feeds P V goes to a scratch (bf16: the
    // backward's dS buffer; fp32: the score scratch, unused by the fp32 forward)
    MB_TRY(seq_attention_fwd(l.qkv, w.addmask, l.ctx, l.probs, w.scores, dt, c.B, S, D, c.n_head, stream,
                             training, 1, site(c.drop_attn, 4 * i + 1),
                             bf ? (l.pdrop != nullptr ? l.pdrop : w.dprobs) : static_cast<void*>(w.scores)));
    MB_TRY(gemm(l.ctx, D, 0, W(p.ao_w), D, 0, M, D, D, epi(EPI_STORE, w.ybuf, bf, D, params + p.ao_b)));
    MB_TRY(postln_fwd(h, w.ybuf, training ? l.s1 : nullptr, params + p.ln1_w, params + p.ln1_b, l.a, dt,
                      training ? l.st1 : nullptr, training ? l.st1 + M : nullptr, M, D, BERT_LN_EPS,
                      stream, site(c.drop_hidden, 4 * i + 2)));
    if (bf) {  // u = gelu_erf(a Wi^T + b) in the GEMM epilogue (z kept only for the backward)
      GemmEpilogue e = epi(EPI_QUICKGELU, training ? l.z : nullptr, 1, F, params + p.i_b);
      e.out2 = l.u; e.ld_out2 = F; e.act = 1;
      MB_TRY(gemm(l.a, D, 0, W(p.i_w), D, 0, M, F, D, e));
    } else {
      MB_TRY(gemm(l.a, D, 0, W(p.i_w), D, 0, M, F, D, epi(EPI_STORE, l.z, 0, F, params + p.i_b)));
      const long long n = static_cast<long long>(M) * F;
      gelu_fwd_kernel<float><<<grid1d(n, 256), 256, 0, stream>>>(static_cast<const float*>(l.z),
                                                                static_cast<float*>(l.u), n);
      MB_CHECK_LAUNCH();
    }
    MB_TRY(gemm(l.u, F, 0, W(p.o_w), F, 0, M, D, F, epi(EPI_STORE, w.ybuf, bf, D, params + p.o_b)));
    MB_TRY(postln_fwd(l.a, w.ybuf, training ? l.s2 : nullptr, params + p.ln2_w, params + p.ln2_b, l.h, dt,
                      training ? l.st2 : nullptr, training ? l.st2 + M : nullptr, M, D, BERT_LN_EPS,
                      stream, site(c.drop_hidden, 4 * i + 3)));
    h = l.h;
  }
  // ---- pooler (tanh(dense(first token))) + classifier (src/mmbt.py:129, :246-247)
  {
    const int gp = (c.B * D + 7) / 8;
    if (bf)
      rows_linear_fwd_kernel<__nv_bfloat16><<<gp, 256, 0, stream>>>(
          static_cast<const __nv_bfloat16*>(h), static_cast<long long>(S) * D, params + lay.pool_w,
          params + lay.pool_b, w.pooled, c.B, D, D, 1);
    else
      rows_linear_fwd_kernel<float><<<gp, 256, 0, stream>>>(
          static_cast<const float*>(h), static_cast<long long>(S) * D, params + lay.pool_w,
          params + lay.pool_b, w.pooled, c.B, D, D, 1);
    MB_CHECK_LAUNCH();
    rows_linear_fwd_kernel<float><<<(c.B * c.C + 7) / 8, 256, 0, stream>>>(
        w.pooled, D, params + lay.clf_w, params + lay.clf_b, logits, c.B, c.C, D, 0);
    MB_CHECK_LAUNCH();
  }
  return 0;
}

int mmbt_backward(const MmbtConfig& c, const float* params, const MmbtInputs& in, void* ws,
                  long long ws_bytes, const float* dlogits, float* grads, cudaStream_t stream) {
  MB_TRY(check_config(c));
  if (params == nullptr || ws == nullptr || dlogits == nullptr || grads == nullptr || in.img == nullptr)
    return MMU_ERR_ARG;
  Layout lay;
  build_layout(c, &lay, nullptr, 0);
  Ws w;
  carve(c, 1, ws, lay, &w);
  if (w.bytes > ws_bytes) return MMU_ERR_WORKSPACE;
  const int S = resolve_S(c, in);
  if (S < 0) return S;
  const bool bf = c.precision == PREC_BF16;
  const int dt = bf ? DT_BF16 : DT_F32;
  const int D = c.D, F = c.d_ff, M = c.B * S, Mi = c.B * c.n_img;
  const Gemm gemm{c.precision, stream};
  const void* shadow = in.params_bf16 != nullptr ? in.params_bf16 : w.params_lp;
  auto W = [&](long long off) -> const void* {
    return bf ? static_cast<const void*>(static_cast<const uint16_t*>(shadow) + off)
              : static_cast<const void*>(params + off);
  };
  // activation-dtype copy of the fp32 gradient stream as the BRANCH sees it (fp32 path without
  // hidden dropout: the stream itself; with hidden dropout the copy carries the dropout mask)
  const dropout::Site off{0u, 0u, 0u, 1.0f};
  auto site = [&](float p, int id) { return p > 0.f ? dropout::make_site(p, in.drop_seed, id) : off; };
  const bool hdrop = c.drop_hidden > 0.f;
  auto LP = [&](float* g) -> void* { return (bf || hdrop) ? w.g_lp : static_cast<void*>(g); };
  auto out_mask = [&](int id) { PostLnDropout d; d.out = site(c.drop_hidden, id); return d; };

  // ---- classifier + pooler
  const void* h_last = w.layer[c.n_layers - 1].h;
  float* dpooled = w.dpooled;
  float* dpre = w.dpooled + static_cast<long long>(c.B) * D;
  rows_linear_bwd_w_kernel<float><<<(c.C * D + 255) / 256, 256, 0, stream>>>(
      dlogits, w.pooled, D, grads + lay.clf_w, grads + lay.clf_b, c.B, c.C, D);
  MB_CHECK_LAUNCH();
  rows_linear_bwd_x_kernel<<<(c.B * D + 255) / 256, 256, 0, stream>>>(dlogits, params + lay.clf_w,
                                                                     dpooled, D, c.B, c.C, D);
  MB_CHECK_LAUNCH();
  tanh_bwd_kernel<<<(c.B * D + 255) / 256, 256, 0, stream>>>(dpooled, w.pooled, dpre, c.B * D);
  MB_CHECK_LAUNCH();
  if (bf)
    rows_linear_bwd_w_kernel<__nv_bfloat16><<<(D * D + 255) / 256, 256, 0, stream>>>(
        dpre, static_cast<const __nv_bfloat16*>(h_last), static_cast<long long>(S) * D,
        grads + lay.pool_w, grads + lay.pool_b, c.B, D, D);
  else
    rows_linear_bwd_w_kernel<float><<<(D * D + 255) / 256, 256, 0, stream>>>(
        dpre, static_cast<const float*>(h_last), static_cast<long long>(S) * D, grads + lay.pool_w,
        grads + lay.pool_b, c.B, D, D);
  MB_CHECK_LAUNCH();
  // gradient of the last hidden state: zero but for the first token of every sample
  if (cudaMemsetAsync(w.gB, 0, static_cast<size_t>(M) * D * 4, stream) != cudaSuccess) return MMU_ERR_CUDA;
  rows_linear_bwd_x_kernel<<<(c.B * D + 255) / 256, 256, 0, stream>>>(
      dpre, params + lay.pool_w, w.gB, static_cast<long long>(S) * D, c.B, D, D);
  MB_CHECK_LAUNCH();

  // ---- encoder, last layer first.  (branch, res) = gradient of the layer's output split into
  //      the part arriving through the next layer's qkv GEMM and the part arriving through its
  //      residual connection.
  const void* branch = nullptr;
  for (int i = c.n_layers - 1; i >= 0; --i) {
    const LayerP& p = lay.layer[i];
    const LayerWs& l = w.layer[i];
    const void* h_in = i > 0 ? w.layer[i - 1].h : w.h0;
    // h' = LN2(s2):  ds2 -> gA ; d(output.dense.bias) += colsum(ds2)
    MB_TRY(postln_bwd(branch, w.gB, dt, l.s2, l.st2, l.st2 + M, params + p.ln2_w, w.gA,
                      (bf || hdrop) ? w.g_lp : nullptr, grads + p.ln2_w, grads + p.ln2_b, grads + p.o_b, M, D,
                      stream, out_mask(4 * i + 3)));
    // s2 = a + u Wo2^T + b:  du = ds2 Wo2 -> dz = du * gelu'(z)
    if (bf) {  // dGELU fused into the dgrad GEMM epilogue (z arrives by TMA)
      GemmEpilogue e = epi(EPI_DGELU, w.dbig, 1, F, nullptr);
      e.aux = l.z; e.ld_aux = F; e.act = 1;
      MB_TRY(gemm(LP(w.gA), D, 0, W(p.o_w), F, 1, M, F, D, e));
    } else {
      MB_TRY(gemm(LP(w.gA), D, 0, W(p.o_w), F, 1, M, F, D, epi(EPI_STORE, w.dbig, 0, F, nullptr)));
      const long long n = static_cast<long long>(M) * F;
      gelu_bwd_kernel<float><<<grid1d(n, 256), 256, 0, stream>>>(static_cast<const float*>(l.z),
                                                                static_cast<float*>(w.dbig), n);
      MB_CHECK_LAUNCH();
    }
    // dWo2[D, F] += ds2^T u ; dWi[F, D] += dz^T a ; dbi += colsum(dz) ; da_branch = dz Wi
    MB_TRY(gemm(LP(w.gA), D, 1, l.u, F, 1, D, F, M, epi(EPI_ATOMIC, grads + p.o_w, 0, F, nullptr),
                wsplits(D, F, M)));
    MB_TRY(gemm(w.dbig, F, 1, l.a, D, 1, F, D, M, epi(EPI_ATOMIC, grads + p.i_w, 0, D, nullptr),
                wsplits(F, D, M)));
    MB_TRY(colsum_accumulate(w.dbig, dt, grads + p.i_b, M, F, stream));
    MB_TRY(gemm(w.dbig, F, 0, W(p.i_w), D, 1, M, D, F, epi(EPI_STORE, w.dh, bf, D, nullptr)));
    // a = LN1(s1):  ds1 = LN1'(da_branch + ds2) -> gB ; d(attention.output.dense.bias) += colsum
    MB_TRY(postln_bwd(w.dh, w.gA, dt, l.s1, l.st1, l.st1 + M, params + p.ln1_w, w.gB,
                      (bf || hdrop) ? w.g_lp : nullptr, grads + p.ln1_w, grads + p.ln1_b, grads + p.ao_b, M, D,
                      stream, out_mask(4 * i + 2)));
    // s1 = h + ctx Wo^T + b:  dWo += ds1^T ctx ; dctx = ds1 Wo
    MB_TRY(gemm(LP(w.gB), D, 1, l.ctx, D, 1, D, D, M, epi(EPI_ATOMIC, grads + p.ao_w, 0, D, nullptr),
                wsplits(D, D, M)));
    MB_TRY(gemm(LP(w.gB), D, 0, W(p.ao_w), D, 1, M, D, D, epi(EPI_STORE, w.dh, bf, D, nullptr)));
    MB_TRY(seq_attention_bwd(l.qkv, w.dh, l.probs, w.scores, w.dprobs, w.dbig, dt, c.B, S, D, c.n_head,
                             stream, site(c.drop_attn, 4 * i + 1), bf ? l.pdrop : nullptr));
    // dWqkv[3D, D] += dqkv^T h ; dbqkv += colsum(dqkv) ; dh_branch = dqkv Wqkv
    MB_TRY(gemm(w.dbig, 3 * D, 1, h_in, D, 1, 3 * D, D, M, epi(EPI_ATOMIC, grads + p.q_w, 0, D, nullptr),
                wsplits(3 * D, D, M)));
    MB_TRY(colsum_accumulate(w.dbig, dt, grads + p.q_b, M, 3 * D, stream));
    MB_TRY(gemm(w.dbig, 3 * D, 0, W(p.q_w), D, 1, M, D, 3 * D, epi(EPI_STORE, w.dh, bf, D, nullptr)));
    branch = w.dh;
  }
  // ---- embeddings: dE = LN_emb'(branch + gB) -> gA, scattered to the tables / image tokens
  PostLnDropout edrop;  // the forward dropped the OUTPUT of the embedding LayerNorm
  edrop.in_a = site(c.drop_img, 0);
  edrop.in_b = site(c.drop_hidden, 0);
  edrop.row_side = w.row_side;
  MB_TRY(postln_bwd(branch, w.gB, dt, w.s0, w.st0, w.st0 + M, params + lay.eln_w, w.gA, nullptr,
                    grads + lay.eln_w, grads + lay.eln_b, nullptr, M, D, stream, edrop));
  if (cudaMemsetAsync(w.dimgp, 0, static_cast<size_t>(Mi) * D * 4, stream) != cudaSuccess) return MMU_ERR_CUDA;
  embed_bwd_kernel<<<grid1d(static_cast<long long>(M) * D, 256), 256, 0, stream>>>(
      w.gA, w.row_word, w.row_pos, w.row_img, M, D, grads + lay.word, grads + lay.pos, w.dimgp);
  MB_CHECK_LAUNCH();
  {
    const int slab = 128;
    dim3 grid((D + 255) / 256, (M + slab - 1) / slab);
    type_bwd_kernel<<<grid, 256, 0, stream>>>(w.gA, w.row_type, M, D, c.n_types, slab, grads + lay.type);
    MB_CHECK_LAUNCH();
  }
  // ---- image-token projection: dW[D, d_img] += dimgp^T X ; db += colsum ; dX = dimgp W
  const void* dip = w.dimgp;
  const void* img_op = in.img;
  if (bf) {
    MB_TRY(cast_f32_to_bf16(w.dimgp, w.dimgp_lp, static_cast<size_t>(Mi) * D, stream));
    dip = w.dimgp_lp;
    img_op = w.img_lp;
  }
  MB_TRY(gemm(dip, D, 1, img_op, c.d_img, 1, D, c.d_img, Mi,
              epi(EPI_ATOMIC, grads + lay.img_w, 0, c.d_img, nullptr), 1));
  MB_TRY(colsum_accumulate(w.dimgp, DT_F32, grads + lay.img_b, Mi, D, stream));
  if (in.dimg != nullptr)
    MB_TRY(gemm(dip, D, 0, W(lay.img_w), c.d_img, 1, Mi, c.d_img, D, epi(EPI_STORE, in.dimg, 0, c.d_img, nullptr)));
  return 0;
}

// =========================================================================== BertAdam
namespace {
__global__ void __launch_bounds__(256)
seg_sumsq_kernel(const float* __restrict__ g, const long long* __restrict__ segs, float* __restrict__ norms,
                 int chunk) {
  const int sgi = blockIdx.y;
  const long long off = segs[2 * sgi], numel = segs[2 * sgi + 1];
  const long long c0 = static_cast<long long>(blockIdx.x) * chunk;
  if (c0 >= numel) return;
  const long long c1 = min(numel, c0 + chunk);
  float acc = 0.f;
  const long long n4 = (c1 - c0) / 4;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 x = *reinterpret_cast<const float4*>(g + off + c0 + 4 * i);
    acc = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, acc))));
  }
  for (long long i = c0 + 4 * n4 + threadIdx.x; i < c1; i += blockDim.x) {
    const float x = g[off + i];
    acc = fmaf(x, x, acc);
  }
  acc = wsum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(norms + sgi, s);
  }
}
__global__ void __launch_bounds__(256)
bertadam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                __nv_bfloat16* __restrict__ p_lp, const long long* __restrict__ segs,
                const float* __restrict__ seg_hyper, const float* __restrict__ norms, int chunk,
                float b1, float b2, float eps, float max_norm, float grad_scale) {
  const int sgi = blockIdx.y;
  const long long off = segs[2 * sgi], numel = segs[2 * sgi + 1];
  const long long c0 = static_cast<long long>(blockIdx.x) * chunk;
  if (c0 >= numel) return;
  const long long c1 = min(numel, c0 + chunk);
  // torch.nn.utils.clip_grad_norm_(p, max_norm): coef = max_norm / (||g|| + 1e-6), applied if < 1
  // grad_scale (1 / world size after a sum-all-reduce) is applied before the clip, as if the
  // averaged gradient had been stored: ||s g|| = s ||g||
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float cc = max_norm / (grad_scale * sqrtf(norms[sgi]) + 1e-6f);
    coef = cc < 1.0f ? cc * grad_scale : grad_scale;
  }
  const float wd = seg_hyper[2 * sgi], lr = seg_hyper[2 * sgi + 1];
  auto update = [&](float gr, float& mm, float& vv, float& pv) {
    gr *= coef;
    mm = b1 * mm + (1.0f - b1) * gr;
    vv = b2 * vv + (1.0f - b2) * gr * gr;
    float upd = mm / (sqrtf(vv) + eps);
    if (wd > 0.f) upd += wd * pv;
    pv -= lr * upd;
    return gr;
  };
  // tensors start on 64-element boundaries and chunks are multiples of 4: 16-byte vectors for the
  // body (30 B/param of traffic: p, g, m, v read, p, m, v (+ bf16 shadow) written), scalars for the tail
  const long long n4 = (c1 - c0) / 4;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const long long k = off + c0 + 4 * i;
    float4 g4 = *reinterpret_cast<const float4*>(g + k), m4 = *reinterpret_cast<const float4*>(m + k);
    float4 v4 = *reinterpret_cast<const float4*>(v + k), p4 = *reinterpret_cast<const float4*>(p + k);
    g4.x = update(g4.x, m4.x, v4.x, p4.x);
    g4.y = update(g4.y, m4.y, v4.y, p4.y);
    g4.z = update(g4.z, m4.z, v4.z, p4.z);
    g4.w = update(g4.w, m4.w, v4.w, p4.w);
    if (coef != 1.0f) *reinterpret_cast<float4*>(g + k) = g4;  // the reference clips p.grad in place
    *reinterpret_cast<float4*>(m + k) = m4;
    *reinterpret_cast<float4*>(v + k) = v4;
    *reinterpret_cast<float4*>(p + k) = p4;
    if (p_lp != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(p_lp + k) = pk;
    }
  }
  for (long long i = c0 + 4 * n4 + threadIdx.x; i < c1; i += blockDim.x) {
    const long long k = off + i;
    float mm = m[k], vv = v[k], pv = p[k];
    const float gr = update(g[k], mm, vv, pv);
    if (coef != 1.0f) g[k] = gr;
    m[k] = mm;
    v[k] = vv;
    p[k] = pv;
    if (p_lp != nullptr) p_lp[k] = __float2bfloat16_rn(pv);
  }
}
}  // namespace

int bertadam_flat(float* p, float* g, float* m, float* v, void* p_bf16, const long long* segs,
                  const float* seg_hyper, float* norms, int n_seg, long long max_seg_numel, float b1,
                  float b2, float eps, float max_grad_norm, float grad_scale, cudaStream_t stream) {
  if (p == nullptr || g == nullptr || m == nullptr || v == nullptr || segs == nullptr ||
      seg_hyper == nullptr || norms == nullptr)
    return MMU_ERR_ARG;
  if (n_seg < 1 || n_seg > 65535 || max_seg_numel < 1) return MMU_ERR_SHAPE;
  // one grid row per tensor; columns cover the largest tensor in chunks
  const int chunk = 1 << 15;
  const long long max_chunks = (max_seg_numel + chunk - 1) / chunk;
  if (cudaMemsetAsync(norms, 0, static_cast<size_t>(n_seg) * 4, stream) != cudaSuccess) return MMU_ERR_CUDA;
  dim3 grid(static_cast<unsigned>(max_chunks), static_cast<unsigned>(n_seg));
  if (max_grad_norm > 0.f) {
    seg_sumsq_kernel<<<grid, 256, 0, stream>>>(g, segs, norms, chunk);
    MB_CHECK_LAUNCH();
  }
  bertadam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), segs,
                                            seg_hyper, norms, chunk, b1, b2, eps, max_grad_norm, grad_scale);
  MB_CHECK_LAUNCH();
  return 0;
}

}  // namespace mmu
