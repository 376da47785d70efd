// FLAVA-fusion engine implementation.  See engine.h.
//
// Data layout in HBM (all row-major; POSITION-major rows: row r = l*B + b holds token position l
// of sample b, so the B rows one batch-axis attention problem reads are adjacent):
//   residual stream x          fp32 [B*L, D]      (kept in fp32 in both precisions)
//   GEMM operands / outputs    fp32 or bf16       (h = LN(x), qkv, attention out, z, u, grads)
//   parameters / gradients     one flat fp32 buffer each (+ bf16 shadow of the parameters)
// Forward, per layer (src/model.py:209-212):
//   h1 = ln_1(x0) ; qkv = h1 Win^T + bin ; o = batch-axis attention(qkv)
//   x1 = x0 + o Wout^T + bout                       (GEMM, residual epilogue)
//   h2 = ln_2(x1) ; z = h2 Wfc^T + bfc ; u = z sigmoid(1.702 z)   (GEMM, QuickGELU epilogue)
//   x2 = x1 + u Wproj^T + bproj                     (GEMM, residual epilogue)
// Backward mirrors it with dgrad GEMMs (MN-major weight operand), split-K wgrad GEMMs
// (MN-major activations, fp32 atomics into the flat gradient buffer), the dGELU epilogue and
// LayerNorm-backward kernels that also emit the bias gradient of the preceding projection.
#include "engine.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "gemm_api.h"
#include "kernels.h"

namespace mmu {

namespace {

constexpr long long ALIGN_ELEMS = 64;

long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

struct LayerParams {
  long long ln1_w, ln1_b, in_w, in_b, out_w, out_b, ln2_w, ln2_b, fc_w, fc_b, proj_w, proj_b;
};
struct Layout {
  long long img_w, img_b, txt_w, txt_b, cls, lnpre_w, lnpre_b;
  LayerParams layer[64];
  long long lnpost_w, lnpost_b, head_w[16], head_b[16];
  long long total;
};

struct TableBuilder {
  ParamEntry* out;
  int max, n;
  long long cursor;
  long long add(const char* name, int rows, int cols, int stage) {
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
      e.stage = stage;
    }
    ++n;
    cursor = align_up(cursor + numel, ALIGN_ELEMS);
    return off;
  }
};

int build_layout(const FlavaConfig& c, Layout* L, ParamEntry* out, int max_entries) {
  if (c.n_layers < 0 || c.n_layers > 64 || c.E < 1 || c.E > 16) return MMU_ERR_SHAPE;
  TableBuilder tb{out, max_entries, 0, 0};
  const int stem = c.n_layers + 1;
  char nm[96];
  L->img_w = tb.add("image_to_mm_projection.weight", c.D, c.d_img, stem);
  L->img_b = tb.add("image_to_mm_projection.bias", c.D, 0, stem);
  L->txt_w = L->txt_b = -1;
  if (c.d_txt > 0) {
    L->txt_w = tb.add("text_to_mm_projection.weight", c.D, c.d_txt, stem);
    L->txt_b = tb.add("text_to_mm_projection.bias", c.D, 0, stem);
  }
  L->cls = -1;
  if (c.cls_token) L->cls = tb.add("class_embeddings", c.D, c.E, stem);
  L->lnpre_w = tb.add("ln_pre.weight", c.D, 0, stem);
  L->lnpre_b = tb.add("ln_pre.bias", c.D, 0, stem);
  for (int i = 0; i < c.n_layers; ++i) {
    LayerParams& p = L->layer[i];
    const int st = c.n_layers - i;  // backward visits the last layer first
    auto name = [&](const char* suffix) {
      std::snprintf(nm, sizeof(nm), "mm_encoder.resblocks.%d.%s", i, suffix);
      return nm;
    };
    p.ln1_w = tb.add(name("ln_1.weight"), c.D, 0, st);
    p.ln1_b = tb.add(name("ln_1.bias"), c.D, 0, st);
    p.in_w = tb.add(name("attn.in_proj_weight"), 3 * c.D, c.D, st);
    p.in_b = tb.add(name("attn.in_proj_bias"), 3 * c.D, 0, st);
    p.out_w = tb.add(name("attn.out_proj.weight"), c.D, c.D, st);
    p.out_b = tb.add(name("attn.out_proj.bias"), c.D, 0, st);
    p.ln2_w = tb.add(name("ln_2.weight"), c.D, 0, st);
    p.ln2_b = tb.add(name("ln_2.bias"), c.D, 0, st);
    p.fc_w = tb.add(name("mlp.c_fc.weight"), 4 * c.D, c.D, st);
    p.fc_b = tb.add(name("mlp.c_fc.bias"), 4 * c.D, 0, st);
    p.proj_w = tb.add(name("mlp.c_proj.weight"), c.D, 4 * c.D, st);
    p.proj_b = tb.add(name("mlp.c_proj.bias"), c.D, 0, st);
  }
  L->lnpost_w = tb.add("ln_post.weight", c.D, 0, 0);
  L->lnpost_b = tb.add("ln_post.bias", c.D, 0, 0);
  for (int e = 0; e < c.E; ++e) {
    std::snprintf(nm, sizeof(nm), "output_layers.%d.weight", e);
    L->head_w[e] = tb.add(nm, c.C, c.D, 0);
    std::snprintf(nm, sizeof(nm), "output_layers.%d.bias", e);
    L->head_b[e] = tb.add(nm, c.C, 0, 0);
  }
  L->total = tb.cursor;
  return tb.n;
}

// ------------------------------------------------------------------ workspace carving
struct LayerWs {
  float* x0;       // layer input (residual stream)
  float* stats1;   // mean | rstd of ln_1
  void* h1;
  void* qkv;
  float* lse;
  void* probs;     // bf16 [L*H, B, Bp] attention probabilities (tensor-core attention path)
  void* o;
  float* x1;
  float* stats2;
  void* h2;
  void* z;
  void* u;
};
struct Ws {
  void* params_lp;
  void* img_t;
  void* txt_t;
  float* mm_x;
  float* stats_pre;
  LayerWs layer[64];
  float* x_out[2];   // ping-pong outputs when activations are not saved
  float* x_final;    // training: dedicated buffer
  float* stats_post;
  float* vec;
  // backward
  float* dvec;
  float* dx;
  void* dx_lp;
  void* dbig;   // dz / dqkv
  void* dh;     // dh2 / do / dh1
  float* delta;
  void* ybuf;      // act dtype [M, D]: branch output of out_proj / c_proj before the residual add
  float* scores;   // fp32 [L*H, B, Bp] scratch (S / dP)
  void* dprobs;    // bf16 [L*H, B, Bp] scratch (dS)
  void* dimg;
  void* dtxt;
  // eval, bf16: LayerNorm folded into the consumer GEMMs (fold_path): per-layer folded weights
  // (in_proj with ln_1, c_fc with ln_2) and two row-statistics buffers [M][nt][2]
  struct FoldedLayer { void* in_wf; float* in_cw; float* in_bf; void* fc_wf; float* fc_cw; float* fc_bf; };
  FoldedLayer fold[64];
  float* fold_stats[2];
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

// Eval forwards of the bf16 engine never materialise LayerNorm outputs: the projections that close
// a residual branch add the fp32 residual stream in their epilogue (EPI_RESID_LN: x' fp32, a raw
// bf16 copy, per-slab row sums) and the GEMMs that consume ln_1 / ln_2 apply the normalisation in
// THEIR epilogue from those sums (GemmEpilogue::ln_stats) -- the add+LayerNorm row kernels and their
// 12 B/element of HBM traffic disappear.  Training keeps the row kernels (their outputs and
// statistics are what the backward reads).  MMU_EVAL_NOFOLD=1: A/B switch back to the row kernels.
int fold_slabs(const FlavaConfig& c) { return (c.D + 127) / 128; }
bool fold_path(const FlavaConfig& c, int training) {
  static const bool disabled = getenv("MMU_EVAL_NOFOLD") != nullptr;
  return !training && !disabled && c.precision == PREC_BF16 && c.n_layers > 0 && fold_slabs(c) <= 8;
}

// The tensor-core path needs 16-byte rows (feature widths % 8): a stem whose input width is not
// (MIMOTransfomer: 14*14 = 196 pixels per view) runs its small projection GEMMs in fp32 instead.
bool stem_is_f32(const FlavaConfig& c) {
  return c.precision != PREC_BF16 || c.d_img % 8 != 0 || c.d_txt % 8 != 0;
}

void carve(const FlavaConfig& c, int training, void* base, const Layout& lay, Ws* w) {
  Bump b{static_cast<char*>(base), 0};
  const long long s = c.precision == PREC_BF16 ? 2 : 4;
  const long long ss = stem_is_f32(c) ? 4 : 2;
  const long long L = (c.cls_token ? c.E : 0) + c.l_img + c.l_txt;
  const long long M = static_cast<long long>(c.B) * L;
  const long long D = c.D;
  const long long Bp = (c.B + 7) / 8 * 8;
  const long long sq_elems = L * c.n_head * c.B * Bp;
  w->params_lp = c.precision == PREC_BF16 ? b.take<void>(lay.total * 2) : nullptr;
  w->img_t = b.take<void>(static_cast<long long>(c.B) * c.l_img * c.d_img * ss);
  w->txt_t = b.take<void>(static_cast<long long>(c.B) * c.l_txt * c.d_txt * ss);
  w->mm_x = b.take<float>(M * D * 4);
  w->stats_pre = b.take<float>(2 * M * 4);
  const int slots = training ? c.n_layers : (c.n_layers > 0 ? 1 : 0);
  for (int i = 0; i < slots; ++i) {
    LayerWs& l = w->layer[i];
    l.x0 = training ? b.take<float>(M * D * 4) : nullptr;
    l.stats1 = b.take<float>(2 * M * 4);
    l.h1 = b.take<void>(M * D * s);
    l.qkv = b.take<void>(M * 3 * D * s);
    l.lse = b.take<float>(L * c.n_head * c.B * 4);
    l.probs = c.precision == PREC_BF16 ? b.take<void>(sq_elems * 2) : nullptr;
    l.o = b.take<void>(M * D * s);
    l.x1 = b.take<float>(M * D * 4);
    l.stats2 = b.take<float>(2 * M * 4);
    l.h2 = b.take<void>(M * D * s);
    l.z = b.take<void>(M * 4 * D * s);
    l.u = b.take<void>(M * 4 * D * s);
  }
  for (int i = slots; i < c.n_layers; ++i) w->layer[i] = w->layer[0];
  w->x_out[0] = b.take<float>(M * D * 4);
  w->x_out[1] = training ? nullptr : b.take<float>(M * D * 4);
  w->x_final = w->x_out[0];
  w->stats_post = b.take<float>(2 * M * 4);
  w->vec = b.take<float>(static_cast<long long>(c.max_variants > 1 ? c.max_variants : 1) * c.B *
                         c.E * D * 4);
  w->ybuf = b.take<void>(M * D * s);
  w->scores = c.precision == PREC_BF16 ? b.take<float>(sq_elems * 4) : nullptr;
  if (training) {
    w->dvec = b.take<float>(static_cast<long long>(c.B) * c.E * D * 4);
    w->dx = b.take<float>(M * D * 4);
    w->dx_lp = c.precision == PREC_BF16 ? b.take<void>(M * D * s) : static_cast<void*>(w->dx);
    w->dbig = b.take<void>(M * 4 * D * s);
    w->dh = b.take<void>(M * D * s);
    w->delta = b.take<float>(L * c.n_head * c.B * 4);
    w->dprobs = c.precision == PREC_BF16 ? b.take<void>(sq_elems * 2) : nullptr;
    w->dimg = b.take<void>(static_cast<long long>(c.B) * c.l_img * D * ss);
    w->dtxt = b.take<void>(static_cast<long long>(c.B) * c.l_txt * D * ss);
  } else {
    w->dvec = nullptr; w->dx = nullptr; w->dx_lp = nullptr; w->dbig = nullptr; w->dh = nullptr;
    w->delta = nullptr; w->dprobs = nullptr; w->dimg = nullptr; w->dtxt = nullptr;
  }
  if (fold_path(c, training)) {
    for (int i = 0; i < c.n_layers; ++i) {
      Ws::FoldedLayer& f = w->fold[i];
      f.in_wf = b.take<void>(3 * D * D * 2);
      f.in_cw = b.take<float>(3 * D * 4);
      f.in_bf = b.take<float>(3 * D * 4);
      f.fc_wf = b.take<void>(4 * D * D * 2);
      f.fc_cw = b.take<float>(4 * D * 4);
      f.fc_bf = b.take<float>(4 * D * 4);
    }
    for (int k = 0; k < 2; ++k) w->fold_stats[k] = b.take<float>(M * fold_slabs(c) * 2 * 4);
  }
  w->bytes = b.off;
}

int check_config(const FlavaConfig& c) {
  if (c.B < 1 || c.D < 4 || c.n_head < 1 || c.E < 1 || c.E > 16 || c.C < 1) return MMU_ERR_SHAPE;
  if (c.D % c.n_head != 0 || (c.D / c.n_head) % 4 != 0) return MMU_ERR_SHAPE;
  if (c.D % 4 != 0 || c.d_img % 4 != 0 || c.d_txt % 4 != 0 || c.D > 1024) return MMU_ERR_SHAPE;
  if (c.d_img < 4 || c.d_txt < 0 || c.group_pool < 0) return MMU_ERR_SHAPE;
  if (c.precision == PREC_BF16 && c.D % 8 != 0) return MMU_ERR_SHAPE;
  if (c.group_pool > 0 && (c.avg_pool || c.cls_token)) return MMU_ERR_ARG;
  if (c.precision != PREC_FP32 && c.precision != PREC_BF16) return MMU_ERR_ARG;
  if (c.avg_pool && !c.cls_token && c.E != 2) return MMU_ERR_SHAPE;  // src/model.py:282-284
  return 0;
}

// One dense contraction, dispatched on precision.  A: [M,K] (K-major) or [K,M] (MN-major).
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
  e.mode = mode;
  e.out_bf16 = out_bf16;
  e.out = out;
  e.ld_out = ld_out;
  e.bias = bias;
  e.alpha = 1.0f;
  e.seg_len = 0;
  return e;
}

// Split-K factor of a weight-gradient GEMM (Mg x Ng output, K = tokens): the few output tiles
// cannot fill the machine, so K is split until the tile count lands just under a whole number of
// waves (CTA pairs: 256x256 tiles over 74 clusters; small Mg: 128x256 tiles over 148 CTAs).
int wgrad_splits(int Mg, int Ng, int K) {
  const bool pair = Mg >= 512;
  const int tiles = ((Mg + (pair ? 255 : 127)) / (pair ? 256 : 128)) * ((Ng + 255) / 256);
  const int units = pair ? gemm_sms() / 2 : gemm_sms();
  const int kb = (K + 63) / 64;
  int smax = kb / 4;
  if (smax > 16) smax = 16;
  if (smax < 1) smax = 1;
  int best_s = 1;
  double best = 0.0;
  for (int sp = 1; sp <= smax; ++sp) {
    const int t = tiles * sp;
    const double eff = static_cast<double>(t) / (static_cast<double>((t + units - 1) / units) * units);
    if (eff > best + 0.02) {  // prefer the smaller factor on near-ties: fewer atomic passes
      best = eff;
      best_s = sp;
    }
  }
  return best_s;
}

#define MMU_TRY(x)            \
  do {                        \
    const int rc__ = (x);     \
    if (rc__ != 0) return rc__; \
  } while (0)

struct Shape {
  int n_cls, n_img, n_txt, L, M;
};
int resolve_shape(const FlavaConfig& c, const FlavaInputs& in, Shape* s) {
  if (in.n_variants > 1) {  // packed variants: eval only, capacity-checked, no CLS rows
    if (c.cls_token || in.var_segments == nullptr || in.keep != nullptr) return MMU_ERR_ARG;
    if (in.n_variants > (c.max_variants > 1 ? c.max_variants : 1)) return MMU_ERR_WORKSPACE;
  }
  s->n_cls = c.cls_token ? c.E : 0;
  s->n_img = in.img != nullptr ? in.n_img : 0;
  s->n_txt = in.txt != nullptr ? in.n_txt : 0;
  if (s->n_img < 0 || s->n_img > c.l_img || s->n_txt < 0 || s->n_txt > c.l_txt) return MMU_ERR_SHAPE;
  if (c.d_txt == 0 && s->n_txt > 0) return MMU_ERR_ARG;  // the model has no text projection
  s->L = s->n_cls + s->n_img + s->n_txt;
  s->M = c.B * s->L;
  if (s->n_img + s->n_txt == 0) return MMU_ERR_SHAPE;
  if ((!c.avg_pool || c.cls_token) && in.n_variants <= 1) {
    // head i reads token position i (src/model.py:286-287) or its group of positions (:148-153)
    if (c.E * (c.group_pool > 0 ? c.group_pool : 1) > s->L) return MMU_ERR_SHAPE;
  }
  return 0;
}

HeadSegments head_segments(const FlavaConfig& c, const Shape& s) {
  HeadSegments hs{};
  hs.E = c.E;
  if (c.group_pool > 0) {
    for (int e = 0; e < c.E; ++e) {
      hs.seg_begin[e] = e * c.group_pool;
      hs.seg_end[e] = (e + 1) * c.group_pool;
    }
  } else if (c.avg_pool && !c.cls_token) {
    hs.seg_begin[0] = 0; hs.seg_end[0] = s.n_img;
    hs.seg_begin[1] = s.n_img; hs.seg_end[1] = s.n_img + s.n_txt;
  } else {
    for (int e = 0; e < c.E; ++e) { hs.seg_begin[e] = e; hs.seg_end[e] = e + 1; }
  }
  return hs;
}

}  // namespace

// =========================================================================== public
int flava_param_table(const FlavaConfig& c, ParamEntry* out, int max_entries) {
  Layout lay;
  return build_layout(c, &lay, out, max_entries);
}

long long flava_param_count(const FlavaConfig& c) {
  Layout lay;
  if (build_layout(c, &lay, nullptr, 0) < 0) return MMU_ERR_SHAPE;
  return lay.total;
}

int flava_num_stages(const FlavaConfig& c) { return c.n_layers + 2; }

long long flava_workspace_bytes(const FlavaConfig& c, int training) {
  if (int rc = check_config(c)) return rc;
  Layout lay;
  if (build_layout(c, &lay, nullptr, 0) < 0) return MMU_ERR_SHAPE;
  Ws w;
  carve(c, training, nullptr, lay, &w);
  return w.bytes;
}

int flava_forward(const FlavaConfig& c, const float* params, const FlavaInputs& in, void* ws,
                  long long ws_bytes, int training, float* logits, cudaStream_t stream) {
  MMU_TRY(check_config(c));
  if (params == nullptr || ws == nullptr || logits == nullptr) return MMU_ERR_ARG;
  Layout lay;
  if (build_layout(c, &lay, nullptr, 0) < 0) return MMU_ERR_SHAPE;
  Ws w;
  carve(c, training, ws, lay, &w);
  if (w.bytes > ws_bytes) return MMU_ERR_WORKSPACE;
  Shape s;
  MMU_TRY(resolve_shape(c, in, &s));
  if (in.n_variants > 1 && training) return MMU_ERR_ARG;
  if (!(in.drop_p >= 0.f && in.drop_p < 1.f)) return MMU_ERR_ARG;
  const int src_l_img = in.src_l_img > 0 ? in.src_l_img : c.l_img;
  const int src_l_txt = in.src_l_txt > 0 ? in.src_l_txt : c.l_txt;
  const int bf = c.precision == PREC_BF16;
  const int dt = bf ? DT_BF16 : DT_F32;
  const int D = c.D, M = s.M;
  const Gemm gemm{c.precision, stream};
  // operand view of the parameters (bf16 shadow refreshed from the fp32 master every forward)
  const void* shadow = in.params_bf16 != nullptr ? in.params_bf16 : w.params_lp;
  auto W = [&](long long off) -> const void* {
    return bf ? static_cast<const void*>(static_cast<const uint16_t*>(shadow) + off)
              : static_cast<const void*>(params + off);
  };
  if (bf && in.params_bf16 == nullptr)
    MMU_TRY(cast_f32_to_bf16(params, w.params_lp, static_cast<size_t>(lay.total), stream));

  // ---- stem: gather/mask/cast inputs, per-modality projections written straight into the
  //      concatenated (B, L, D) buffer (fuses torch.cat, src/model.py:262-273), CLS rows
  const bool stem32 = stem_is_f32(c);  // fp32 stem GEMMs read the fp32 master weights
  const Gemm gemm_stem{stem32 ? static_cast<int>(PREC_FP32) : c.precision, stream};
  const int dt_stem = stem32 ? DT_F32 : DT_BF16;
  auto Wstem = [&](long long off) -> const void* {
    return stem32 ? static_cast<const void*>(params + off) : W(off);
  };
  // Rows are POSITION-major from here on: row (l, b) = l*B + b.  The concatenation [CLS | image |
  // text] along the token axis (torch.cat, src/model.py:273) is then a concatenation of contiguous
  // row blocks, and the B rows of one token position -- the operands of its batch-axis attention
  // problems -- are adjacent in memory.
  if (s.n_img > 0) {
    MMU_TRY(cast_gather(in.img, w.img_t, dt_stem, c.B, src_l_img, c.d_img, in.idx_img, s.n_img,
                        in.keep, 0, stream, 1, in.src_bf16 ? DT_BF16 : DT_F32));
    GemmEpilogue e = epi(EPI_STORE, w.mm_x + static_cast<long long>(s.n_cls) * c.B * D, 0, D,
                         params + lay.img_b);
    MMU_TRY(gemm_stem(w.img_t, c.d_img, 0, Wstem(lay.img_w), c.d_img, 0, c.B * s.n_img, D, c.d_img, e));
  }
  if (s.n_txt > 0) {
    MMU_TRY(cast_gather(in.txt, w.txt_t, dt_stem, c.B, src_l_txt, c.d_txt, in.idx_txt, s.n_txt,
                        in.keep, 1, stream, 1, in.src_bf16 ? DT_BF16 : DT_F32));
    GemmEpilogue e = epi(EPI_STORE, w.mm_x + static_cast<long long>(s.n_cls + s.n_img) * c.B * D, 0, D,
                         params + lay.txt_b);
    MMU_TRY(gemm_stem(w.txt_t, c.d_txt, 0, Wstem(lay.txt_w), c.d_txt, 0, c.B * s.n_txt, D, c.d_txt, e));
  }
  if (c.cls_token) MMU_TRY(cls_fill(params + lay.cls, w.mm_x, c.B, s.L, D, c.E, stream));

  float* x = training && c.n_layers > 0 ? w.layer[0].x0 : w.x_out[0];
  if (fold_path(c, training)) {
    // ---- eval, bf16: no LayerNorm output ever reaches HBM (see fold_path)
    const int nt = fold_slabs(c);
    const float inv_d = 1.0f / static_cast<float>(D);
    {  // folded weights of every layer in one launch (layers are laid out at fixed strides)
      const LayerParams& p = lay.layer[0];
      const Ws::FoldedLayer& f = w.fold[0];
      const long long pstride = c.n_layers > 1 ? lay.layer[1].in_w - lay.layer[0].in_w : 0;
      const long long wstride = c.n_layers > 1 ? static_cast<char*>(w.fold[1].in_wf) - static_cast<char*>(w.fold[0].in_wf) : 0;
      MMU_TRY(ln_fold_weights(params + p.in_w, params + p.ln1_w, params + p.ln1_b, params + p.in_b, f.in_wf,
                              f.in_cw, f.in_bf, 3 * D, params + p.fc_w, params + p.ln2_w, params + p.ln2_b,
                              params + p.fc_b, f.fc_wf, f.fc_cw, f.fc_bf, 4 * D, D, stream, c.n_layers, pstride,
                              wstride));
    }
    auto folded = [&](GemmEpilogue e, const float* stats, const float* cw) {
      e.ln_stats = stats; e.ln_cw = cw; e.ln_nt = nt; e.ln_inv_d = inv_d; e.ln_eps = 1e-5f;
      return e;
    };
    auto resid = [&](const float* x_in, float* x_out, void* raw, float* stats, const float* bias) {
      GemmEpilogue e = epi(EPI_RESID_LN, x_out, 0, D, bias);
      e.aux = x_in; e.ld_aux = D;
      e.out2 = raw; e.ld_out2 = D;
      e.stats_out = stats; e.stats_nt = nt;
      return e;
    };
    const LayerWs& l = w.layer[0];  // eval: one slot shared by every layer
    // x = ln_pre(mm_x) in fp32 + its raw bf16 copy + row sums (the first ln_1 is folded into in_proj)
    MMU_TRY(layernorm_raw_stats_fwd(w.mm_x, params + lay.lnpre_w, params + lay.lnpre_b, x, l.h1,
                                    w.fold_stats[0], nt, M, D, stream));
    for (int i = 0; i < c.n_layers; ++i) {
      const LayerParams& p = lay.layer[i];
      const Ws::FoldedLayer& f = w.fold[i];
      float* x_next = w.x_out[(i + 1) & 1];
      const bool last = i + 1 == c.n_layers;
      MMU_TRY(gemm(l.h1, D, 0, f.in_wf, D, 0, M, 3 * D, D,
                   folded(epi(EPI_STORE, l.qkv, 1, 3 * D, f.in_bf), w.fold_stats[0], f.in_cw)));
      MMU_TRY(attention_fwd(l.qkv, l.o, l.lse, l.probs, w.scores, dt, c.B, s.L, D, c.n_head, stream, 1, 0));
      MMU_TRY(gemm(l.o, D, 0, W(p.out_w), D, 0, M, D, D,
                   resid(x, l.x1, l.h2, w.fold_stats[1], params + p.out_b)));
      {
        GemmEpilogue e = epi(EPI_QUICKGELU, nullptr, 1, 4 * D, f.fc_bf);
        e.out2 = l.u; e.ld_out2 = 4 * D;
        MMU_TRY(gemm(l.h2, D, 0, f.fc_wf, D, 0, M, 4 * D, D, folded(e, w.fold_stats[1], f.fc_cw)));
      }
      MMU_TRY(gemm(l.u, 4 * D, 0, W(p.proj_w), 4 * D, 0, M, D, 4 * D,
                   resid(l.x1, x_next, last ? nullptr : l.h1, last ? nullptr : w.fold_stats[0],
                         params + p.proj_b)));
      x = x_next;
    }
  } else {
  if (c.n_layers > 0) {  // ln_pre chained with the first block's ln_1 in one pass over the rows
    const LayerParams& p0 = lay.layer[0];
    const LayerWs& l0 = w.layer[0];
    MMU_TRY(layernorm2_fwd(w.mm_x, params + lay.lnpre_w, params + lay.lnpre_b, x, w.stats_pre,
                           w.stats_pre + M, params + p0.ln1_w, params + p0.ln1_b, l0.h1, dt,
                           l0.stats1, l0.stats1 + M, M, D, stream));
  } else {
    MMU_TRY(layernorm_fwd(w.mm_x, params + lay.lnpre_w, params + lay.lnpre_b, x, DT_F32, w.stats_pre,
                          w.stats_pre + M, M, D, stream));
  }

  // ---- transformer blocks.  The projections that close a residual branch (out_proj, c_proj)
  //      store the branch output y in the activation dtype; the add x + y is fused into the
  //      LayerNorm kernel that consumes the sum (or into a plain add after the last block).
  for (int i = 0; i < c.n_layers; ++i) {
    const LayerParams& p = lay.layer[i];
    const LayerWs& l = w.layer[i];
    float* x_next;
    if (training) x_next = (i + 1 < c.n_layers) ? w.layer[i + 1].x0 : w.x_final;
    else x_next = w.x_out[(i + 1) & 1];
    // h1 / stats1 of this block were produced by the stem (block 0) or by the previous block's
    // closing add+LN
    MMU_TRY(gemm(l.h1, D, 0, W(p.in_w), D, 0, M, 3 * D, D,
                 epi(EPI_STORE, l.qkv, bf, 3 * D, params + p.in_b)));
    MMU_TRY(attention_fwd(l.qkv, l.o, l.lse, l.probs, w.scores, dt, c.B, s.L, D, c.n_head, stream, 1));
    MMU_TRY(gemm(l.o, D, 0, W(p.out_w), D, 0, M, D, D, epi(EPI_STORE, w.ybuf, bf, D, params + p.out_b)));
    MMU_TRY(add_layernorm_fwd(x, w.ybuf, l.x1, params + p.ln2_w, params + p.ln2_b, l.h2, dt, l.stats2,
                              l.stats2 + M, M, D, stream));
    {
      GemmEpilogue e = epi(EPI_QUICKGELU, training ? l.z : nullptr, bf, 4 * D, params + p.fc_b);
      e.out2 = l.u; e.ld_out2 = 4 * D;
      // nn.Dropout between c_fc and QuickGELU (src/model.py:195-201): training with p > 0 only
      if (training && in.drop_p > 0.f) e.drop = dropout::make_site(in.drop_p, in.drop_seed, i);
      MMU_TRY(gemm(l.h2, D, 0, W(p.fc_w), D, 0, M, 4 * D, D, e));
    }
    MMU_TRY(gemm(l.u, 4 * D, 0, W(p.proj_w), 4 * D, 0, M, D, 4 * D,
                 epi(EPI_STORE, w.ybuf, bf, D, params + p.proj_b)));
    if (i + 1 < c.n_layers) {
      const LayerParams& pn = lay.layer[i + 1];
      const LayerWs& ln = w.layer[i + 1];
      MMU_TRY(add_layernorm_fwd(l.x1, w.ybuf, x_next, params + pn.ln1_w, params + pn.ln1_b, ln.h1, dt,
                                ln.stats1, ln.stats1 + M, M, D, stream));
    } else {
      MMU_TRY(add_layernorm_fwd(l.x1, w.ybuf, x_next, nullptr, nullptr, nullptr, dt, nullptr, nullptr,
                                M, D, stream));
    }
    x = x_next;
  }

  }  // !fold_path

  // ---- ln_post + row gather / mean pooling + heads (src/model.py:277-289)
  HeadParams hp{};
  for (int e = 0; e < c.E; ++e) {
    hp.w[e] = params + lay.head_w[e];
    hp.b[e] = params + lay.head_b[e];
  }
  if (in.n_variants > 1) {
    MMU_TRY(pool_ln_fwd_variants(x, params + lay.lnpost_w, params + lay.lnpost_b, in.var_segments,
                                 in.n_variants, c.E, w.vec, c.B, s.L, D, stream));
    MMU_TRY(heads_fwd(w.vec, hp, logits, in.n_variants * c.B, c.E, c.C, D, stream));
    return 0;
  }
  const HeadSegments hs = head_segments(c, s);
  MMU_TRY(pool_ln_fwd(x, params + lay.lnpost_w, params + lay.lnpost_b, hs, w.vec, w.stats_post,
                      w.stats_post + M, c.B, s.L, D, stream));
  MMU_TRY(heads_fwd(w.vec, hp, logits, c.B, c.E, c.C, D, stream));
  return 0;
}

int flava_backward(const FlavaConfig& c, const float* params, const FlavaInputs& in, void* ws,
                   long long ws_bytes, const float* dlogits, float* grads, int stage_begin,
                   int stage_end, cudaStream_t stream) {
  MMU_TRY(check_config(c));
  if (params == nullptr || ws == nullptr || dlogits == nullptr || grads == nullptr) return MMU_ERR_ARG;
  Layout lay;
  if (build_layout(c, &lay, nullptr, 0) < 0) return MMU_ERR_SHAPE;
  Ws w;
  carve(c, 1, ws, lay, &w);
  if (w.bytes > ws_bytes) return MMU_ERR_WORKSPACE;
  Shape s;
  MMU_TRY(resolve_shape(c, in, &s));
  if (in.n_variants > 1) return MMU_ERR_ARG;
  const int bf = c.precision == PREC_BF16;
  const int dt = bf ? DT_BF16 : DT_F32;
  const int D = c.D, M = s.M;
  const Gemm gemm{c.precision, stream};
  const void* shadow = in.params_bf16 != nullptr ? in.params_bf16 : w.params_lp;
  auto W = [&](long long off) -> const void* {
    return bf ? static_cast<const void*>(static_cast<const uint16_t*>(shadow) + off)
              : static_cast<const void*>(params + off);
  };
  const int n_stages = c.n_layers + 2;
  if (stage_begin < 0) stage_begin = 0;
  if (stage_end > n_stages) stage_end = n_stages;

  for (int st = stage_begin; st < stage_end; ++st) {
    if (st == 0) {
      // ---- heads + ln_post: dx (fp32) for the head rows, zero elsewhere
      const HeadSegments hs = head_segments(c, s);
      HeadParams hp{};
      for (int e = 0; e < c.E; ++e) {
        hp.w[e] = params + lay.head_w[e];
        hp.b[e] = params + lay.head_b[e];
        hp.dw[e] = grads + lay.head_w[e];
        hp.db[e] = grads + lay.head_b[e];
      }
      MMU_TRY(heads_bwd(dlogits, w.vec, hp, w.dvec, c.B, c.E, c.C, D, stream));
      if (cudaMemsetAsync(w.dx, 0, static_cast<size_t>(M) * D * 4, stream) != cudaSuccess)
        return MMU_ERR_CUDA;
      const float* x_last = c.n_layers > 0 ? w.x_final : w.x_out[0];
      MMU_TRY(pool_ln_bwd(w.dvec, x_last, w.stats_post, w.stats_post + M, params + lay.lnpost_w, hs,
                          w.dx, grads + lay.lnpost_w, grads + lay.lnpost_b, c.B, s.L, D, stream));
      if (c.n_layers > 0) {
        // bias gradient of the last c_proj = column sum of dx; later layers get it from ln_1 bwd
        MMU_TRY(colsum_accumulate(w.dx, DT_F32, grads + lay.layer[c.n_layers - 1].proj_b, M, D,
                                  stream));
        if (bf) MMU_TRY(cast_f32_to_bf16(w.dx, w.dx_lp, static_cast<size_t>(M) * D, stream));
      }
    } else if (st <= c.n_layers) {
      const int i = c.n_layers - st;
      const LayerParams& p = lay.layer[i];
      const LayerWs& l = w.layer[i];
      // x2 = x1 + u Wproj^T + bproj
      {  // dz = (dx Wproj) * gelu'(z)
        GemmEpilogue e = epi(EPI_DGELU, w.dbig, bf, 4 * D, nullptr);
        e.aux = l.z; e.ld_aux = 4 * D;
        if (in.drop_p > 0.f) e.drop = dropout::make_site(in.drop_p, in.drop_seed, i);  // same mask
        MMU_TRY(gemm(w.dx_lp, D, 0, W(p.proj_w), 4 * D, 1, M, 4 * D, D, e));
      }
      // dWproj[D, 4D] += dx^T u
      MMU_TRY(gemm(w.dx_lp, D, 1, l.u, 4 * D, 1, D, 4 * D, M,
                   epi(EPI_ATOMIC, grads + p.proj_w, 0, 4 * D, nullptr),
                   wgrad_splits(D, 4 * D, M)));
      // dWfc[4D, D] += dz^T h2 ; dbfc += colsum(dz)
      MMU_TRY(gemm(w.dbig, 4 * D, 1, l.h2, D, 1, 4 * D, D, M,
                   epi(EPI_ATOMIC, grads + p.fc_w, 0, D, nullptr), wgrad_splits(4 * D, D, M)));
      MMU_TRY(colsum_accumulate(w.dbig, dt, grads + p.fc_b, M, 4 * D, stream));
      // dh2 = dz Wfc
      MMU_TRY(gemm(w.dbig, 4 * D, 0, W(p.fc_w), D, 1, M, D, 4 * D,
                   epi(EPI_STORE, w.dh, bf, D, nullptr)));
      // dx1 = dx + LN2'(dh2); dbout += colsum(dx1)
      MMU_TRY(layernorm_bwd(w.dh, dt, l.x1, l.stats2, l.stats2 + M, params + p.ln2_w, w.dx, 1,
                            bf ? w.dx_lp : nullptr, dt, grads + p.ln2_w, grads + p.ln2_b,
                            grads + p.out_b, M, D, stream));
      // x1 = x0 + o Wout^T + bout:  dWout += dx1^T o ; do = dx1 Wout
      MMU_TRY(gemm(w.dx_lp, D, 1, l.o, D, 1, D, D, M, epi(EPI_ATOMIC, grads + p.out_w, 0, D, nullptr),
                   wgrad_splits(D, D, M)));
      MMU_TRY(gemm(w.dx_lp, D, 0, W(p.out_w), D, 1, M, D, D, epi(EPI_STORE, w.dh, bf, D, nullptr)));
      // attention backward: dqkv
      MMU_TRY(attention_bwd(l.qkv, l.o, w.dh, l.lse, w.delta, l.probs, w.scores, w.dprobs, w.dbig,
                            dt, c.B, s.L, D, c.n_head, stream, 1));
      // dWin[3D, D] += dqkv^T h1 ; dbin += colsum(dqkv) ; dh1 = dqkv Win
      MMU_TRY(gemm(w.dbig, 3 * D, 1, l.h1, D, 1, 3 * D, D, M,
                   epi(EPI_ATOMIC, grads + p.in_w, 0, D, nullptr), wgrad_splits(3 * D, D, M)));
      MMU_TRY(colsum_accumulate(w.dbig, dt, grads + p.in_b, M, 3 * D, stream));
      MMU_TRY(gemm(w.dbig, 3 * D, 0, W(p.in_w), D, 1, M, D, 3 * D,
                   epi(EPI_STORE, w.dh, bf, D, nullptr)));
      // dx0 = dx1 + LN1'(dh1); previous layer's dbproj += colsum(dx0)
      float* dprev_bias = i > 0 ? grads + lay.layer[i - 1].proj_b : nullptr;
      MMU_TRY(layernorm_bwd(w.dh, dt, l.x0, l.stats1, l.stats1 + M, params + p.ln1_w, w.dx, 1,
                            bf ? w.dx_lp : nullptr, dt, grads + p.ln1_w, grads + p.ln1_b,
                            dprev_bias, M, D, stream));
    } else {
      // ---- stem: ln_pre backward, CLS rows, per-modality projection wgrads
      // dmm = LNpre'(dx), written (not accumulated) into an fp32 [M, D] buffer that is free by now
      float* dmm = c.n_layers > 0 ? w.layer[0].x1 : w.x_out[0];
      MMU_TRY(layernorm_bwd(w.dx, DT_F32, w.mm_x, w.stats_pre, w.stats_pre + M, params + lay.lnpre_w,
                            dmm, 0, nullptr, DT_F32, grads + lay.lnpre_w, grads + lay.lnpre_b,
                            nullptr, M, D, stream));
      if (c.cls_token) MMU_TRY(cls_bwd(dmm, grads + lay.cls, c.B, s.L, D, c.E, stream));
      const bool stem32 = stem_is_f32(c);
      const Gemm gemm_stem{stem32 ? static_cast<int>(PREC_FP32) : c.precision, stream};
      const int dt_stem = stem32 ? DT_F32 : DT_BF16;
      MMU_TRY(split_rows(dmm, w.dimg, w.dtxt, dt_stem, c.B, s.L, s.n_cls, s.n_img, s.n_txt, D, stream));
      if (s.n_img > 0) {
        const int Mi = c.B * s.n_img;
        MMU_TRY(gemm_stem(w.dimg, D, 1, w.img_t, c.d_img, 1, D, c.d_img, Mi,
                          epi(EPI_ATOMIC, grads + lay.img_w, 0, c.d_img, nullptr),
                          wgrad_splits(D, c.d_img, Mi)));
        MMU_TRY(colsum_accumulate(w.dimg, dt_stem, grads + lay.img_b, Mi, D, stream));
      }
      if (s.n_txt > 0) {
        const int Mt = c.B * s.n_txt;
        MMU_TRY(gemm_stem(w.dtxt, D, 1, w.txt_t, c.d_txt, 1, D, c.d_txt, Mt,
                          epi(EPI_ATOMIC, grads + lay.txt_w, 0, c.d_txt, nullptr),
                          wgrad_splits(D, c.d_txt, Mt)));
        MMU_TRY(colsum_accumulate(w.dtxt, dt_stem, grads + lay.txt_b, Mt, D, stream));
      }
    }
  }
  return 0;
}

}  // namespace mmu
