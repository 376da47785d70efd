"""Operator-level Python wrappers over the C ABI (one function per exported kernel family).

These are what the parity tests drive; the model classes in ``src/`` use the engine entry points
instead.  Every function takes CUDA tensors, enqueues on torch's current stream and raises
``MMUError`` on failure.  Nothing here computes anything in PyTorch.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32, check, lib, ptr, stream_ptr


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("mmu ops need contiguous CUDA tensors")


def gemm(A, B, *, a_mn_major=False, b_mn_major=False, mode=_lib.EPI_STORE, out=None, out2=None,
         bias=None, aux=None, alpha=1.0, splits=1, seg=None, out_dtype=None, dropout=None,
         ln_fold=None, stats_out=None):
    """C[M,N] = epilogue(sum_k A(m,k) B(n,k)); see ``mmu_gemm`` in include/mmu_b200.h.
    ``ln_fold=(stats, cw, eps)``: LayerNorm folded into a STORE / QUICKGELU epilogue (A = raw rows,
    B = folded weights, stats fp32 (M, nt, 2)); ``stats_out`` fp32 (M, nt, 2): EPI_RESID_LN's row sums."""
    _cuda(A, B, out, out2, bias, aux, stats_out)
    dt = _dt(A)
    if a_mn_major:
        K, M = A.shape
    else:
        M, K = A.shape
    N = B.shape[1] if b_mn_major else B.shape[0]
    if out is None and mode != _lib.EPI_QUICKGELU:
        odt = out_dtype or (A.dtype if mode in (_lib.EPI_STORE, _lib.EPI_DGELU) else torch.float32)
        out = torch.empty(M, N, device=A.device, dtype=odt)
    ref = out if out is not None else out2
    e = _lib.GemmEpilogue()
    e.mode = mode
    e.out_lp = 1 if ref.dtype == torch.bfloat16 else 0
    e.out, e.out2, e.bias, e.aux = ptr(out), ptr(out2), ptr(bias), ptr(aux)
    e.ld_out = out.stride(0) if out is not None else 0
    e.ld_out2 = out2.stride(0) if out2 is not None else 0
    e.ld_aux = aux.stride(0) if aux is not None else 0
    e.seg_len, e.seg_stride, e.seg_off = seg if seg is not None else (0, 0, 0)
    e.alpha = alpha
    if dropout is not None:   # (p, seed, site): QUICKGELU / DGELU modes only
        e.drop_p, e.drop_seed, e.drop_site = float(dropout[0]), int(dropout[1]), int(dropout[2])
    if ln_fold is not None:
        stats, cw, eps = ln_fold
        _cuda(stats, cw)
        e.ln_stats, e.ln_cw, e.ln_nt = ptr(stats), ptr(cw), stats.shape[1]
        e.ln_inv_d, e.ln_eps = 1.0 / K, float(eps)
    if stats_out is not None:
        e.stats_out, e.stats_nt = ptr(stats_out), stats_out.shape[1]
    check(lib.mmu_gemm(dt, ptr(A), A.stride(0), int(a_mn_major), ptr(B), B.stride(0),
                       int(b_mn_major), M, N, K, splits, C.byref(e), stream_ptr()), "mmu_gemm")
    return out if out is not None else out2


def ln_fold_weights(W, gamma, beta, bias=None):
    """(Wf bf16 (N, K), cw fp32 (N,), bf fp32 (N,)) of a Linear that consumes LayerNorm output."""
    _cuda(W, gamma, beta, bias)
    N, K = W.shape
    Wf = torch.empty(N, K, device=W.device, dtype=torch.bfloat16)
    cw = torch.empty(N, device=W.device, dtype=torch.float32)
    bf = torch.empty(N, device=W.device, dtype=torch.float32)
    check(lib.mmu_ln_fold_weights(ptr(W), ptr(gamma), ptr(beta), ptr(bias), ptr(Wf), ptr(cw), ptr(bf), N, K,
                                  stream_ptr()), "mmu_ln_fold_weights")
    return Wf, cw, bf


def layernorm_raw_stats(x, gamma, beta, nt):
    """(y fp32, yraw bf16, stats fp32 (M, nt, 2)): LayerNorm output, its raw bf16 copy and row sums."""
    _cuda(x, gamma, beta)
    M, D = x.shape
    y = torch.empty_like(x)
    yraw = torch.empty(M, D, device=x.device, dtype=torch.bfloat16)
    stats = torch.empty(M, nt, 2, device=x.device, dtype=torch.float32)
    check(lib.mmu_layernorm_raw_stats(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(yraw), ptr(stats), nt, M, D,
                                      stream_ptr()), "mmu_layernorm_raw_stats")
    return y, yraw, stats


def mask_gather_tokens(src, idx=None, keep=None, modality=0, dtype=torch.float32, pos_major=False):
    """(B, l_src, d) -> (B, n_sel, d), or (n_sel, B, d) with ``pos_major`` (the engine's row order)."""
    _cuda(src, idx, keep)
    Bn, l_src, d = src.shape
    n_sel = l_src if idx is None else idx.numel()
    shape = (n_sel, Bn, d) if pos_major else (Bn, n_sel, d)
    out = torch.empty(*shape, device=src.device, dtype=dtype)
    check(lib.mmu_mask_gather_tokens(ptr(src), ptr(out), BF16 if dtype == torch.bfloat16 else F32,
                                     Bn, l_src, d, ptr(idx), n_sel, ptr(keep), modality, int(pos_major),
                                     stream_ptr()), "mmu_mask_gather_tokens")
    return out


def layernorm_fwd(x, gamma, beta, out_dtype=torch.float32):
    _cuda(x, gamma, beta)
    M, D = x.shape
    y = torch.empty(M, D, device=x.device, dtype=out_dtype)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    check(lib.mmu_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), _dt(y), ptr(mean), ptr(rstd),
                                M, D, stream_ptr()), "mmu_layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dx=None, want_lp=False, want_colsum=False):
    _cuda(dy, x, mean, rstd, gamma, dx)
    M, D = x.shape
    accumulate = dx is not None
    if dx is None:
        dx = torch.empty(M, D, device=x.device, dtype=torch.float32)
    dgamma = torch.zeros(D, device=x.device, dtype=torch.float32)
    dbeta = torch.zeros(D, device=x.device, dtype=torch.float32)
    colsum = torch.zeros(D, device=x.device, dtype=torch.float32) if want_colsum else None
    dx_lp = torch.empty(M, D, device=x.device, dtype=torch.bfloat16) if want_lp else None
    check(lib.mmu_layernorm_bwd(ptr(dy), _dt(dy), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dx),
                                int(accumulate), ptr(dx_lp), BF16, ptr(dgamma), ptr(dbeta),
                                ptr(colsum), M, D, stream_ptr()), "mmu_layernorm_bwd")
    return dx, dgamma, dbeta, dx_lp, colsum


def _tc_attention(qkv, D, H):
    return qkv.dtype == torch.bfloat16 and (D // H) % 64 == 0


def attention_fwd(qkv, B, L, D, H, pos_major=False, keep_probs=True):
    """Returns (out, saved): saved = lse (SIMT path) or the bf16 probabilities (tensor-core path).
    Rows of ``qkv`` / ``out``: ``b*L + l``, or ``l*B + b`` with ``pos_major``.
    ``keep_probs=False`` (eval: no backward follows) lets bf16 problems with head_dim 256 and
    B <= 128 run as ONE fused kernel; ``saved`` is then not written."""
    _cuda(qkv)
    out = torch.empty(B * L, D, device=qkv.device, dtype=qkv.dtype)
    if _tc_attention(qkv, D, H):
        Bp = (B + 7) // 8 * 8
        probs = torch.empty(L * H, B, Bp, device=qkv.device, dtype=torch.bfloat16)
        scores = torch.empty(L * H, B, Bp, device=qkv.device, dtype=torch.float32)
        flags = int(bool(pos_major)) | (0 if keep_probs else 2)
        check(lib.mmu_batchaxis_attention_fwd(ptr(qkv), ptr(out), 0, ptr(probs), ptr(scores),
                                              _dt(qkv), B, L, D, H, flags, stream_ptr()),
              "mmu_batchaxis_attention_fwd")
        return out, probs
    lse = torch.empty(L * H * B, device=qkv.device, dtype=torch.float32)
    check(lib.mmu_batchaxis_attention_fwd(ptr(qkv), ptr(out), ptr(lse), 0, 0, _dt(qkv), B, L, D, H,
                                          int(pos_major), stream_ptr()), "mmu_batchaxis_attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, saved, B, L, D, H, pos_major=False):
    _cuda(qkv, out, dout, saved)
    dqkv = torch.empty_like(qkv)
    if _tc_attention(qkv, D, H):
        scores = torch.empty(saved.shape, device=qkv.device, dtype=torch.float32)
        dprobs = torch.empty_like(saved)
        check(lib.mmu_batchaxis_attention_bwd(ptr(qkv), ptr(out), ptr(dout), 0, 0, ptr(saved),
                                              ptr(scores), ptr(dprobs), ptr(dqkv), _dt(qkv), B, L, D,
                                              H, int(pos_major), stream_ptr()), "mmu_batchaxis_attention_bwd")
        return dqkv
    delta = torch.empty(L * H * B, device=qkv.device, dtype=torch.float32)
    check(lib.mmu_batchaxis_attention_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(saved), ptr(delta), 0, 0,
                                          0, ptr(dqkv), _dt(qkv), B, L, D, H, int(pos_major), stream_ptr()),
          "mmu_batchaxis_attention_bwd")
    return dqkv


def new_accum(device):
    """Zeroed device ``mmu_metric_accum`` as an int64 tensor of ACC_WORDS 8-byte words."""
    return torch.zeros(_lib.ACC_WORDS, device=device, dtype=torch.int64)


def accum_to_dict(accum):
    """Host copy of an accumulator (ONE device->host transfer) as named numpy arrays."""
    import numpy as np
    raw = accum.detach().cpu().numpy()
    ints, dbl = raw.view(np.uint64), raw.view(np.float64)
    o = _lib.ACC_OFF
    return {
        "conf_count": ints[o["conf_count"]:o["conf_count"] + 15].astype(np.int64),
        "conf_correct": ints[o["conf_correct"]:o["conf_correct"] + 15].astype(np.int64),
        "hpred_count": ints[o["hpred_count"]:o["hpred_count"] + 32].astype(np.int64),
        "mi_count": ints[o["mi_count"]:o["mi_count"] + 32].astype(np.int64),
        "n_samples": int(ints[o["n_samples"]]), "n_rows": int(ints[o["n_rows"]]),
        "n_correct_rows": int(ints[o["n_correct_rows"]]),
        "n_correct_prob": int(ints[o["n_correct_prob"]]),
        "conf_sum": dbl[o["conf_sum"]:o["conf_sum"] + 15].copy(),
        "loss_sum": float(dbl[o["loss_sum"]]), "sum_h_pred": float(dbl[o["sum_h_pred"]]),
        "sum_h_exp": float(dbl[o["sum_h_exp"]]), "sum_mi": float(dbl[o["sum_mi"]]),
    }


def heads_uncertainty_epilogue(logits, labels, mode, *, grad_scale=0.0, want_grad=False,
                               want_pred=False, want_scores=False, accum=None):
    """Fused CE / accuracy / uncertainty / histogram epilogue over logits (N, E, C).
    ``labels``: int64 (N,) or (N, E).  Returns (dlogits, pred[N,2], scores[N,4], accum)."""
    _cuda(logits, labels)
    N, E, Cn = logits.shape
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    if labels.dim() == 1 or labels.shape[1] == 1:
        ls, les = 1, 0
    else:
        ls, les = labels.stride(0), labels.stride(1)
    dl = torch.empty_like(logits) if want_grad else None
    pred = torch.empty(N, 2, device=logits.device, dtype=torch.int32) if want_pred else None
    scores = torch.empty(N, 4, device=logits.device, dtype=torch.float32) if want_scores else None
    if accum is None:
        accum = new_accum(logits.device)
    check(lib.mmu_heads_uncertainty_epilogue(ptr(logits), ptr(labels), ls, les, N, E, Cn, mode,
                                             grad_scale, ptr(dl), ptr(pred), ptr(scores),
                                             ptr(accum), stream_ptr()),
          "mmu_heads_uncertainty_epilogue")
    return dl, pred, scores, accum


def adamw_flat_step(p, g, m, v, step, lr, *, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-3,
                    grad_scale=1.0, p_bf16=None):
    _cuda(p, g, m, v, p_bf16)
    check(lib.mmu_adamw_flat_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_bf16), p.numel(), lr,
                                  betas[0], betas[1], eps, weight_decay, step, grad_scale,
                                  stream_ptr()), "mmu_adamw_flat_step")


def posthoc_scoring(logits, labels, n_repeats, accum=None, want_p_true=False):
    """Post-hoc robustness statistics from packed-variant logits (V, B, E, C); see
    ``mmu_posthoc_scoring``.  Returns (accum uint8 tensor viewing ``mmu_posthoc_accum``, p_true
    (B, V) or None)."""
    _cuda(logits, labels)
    V, B, E, Cn = logits.shape
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    if accum is None:
        accum = torch.zeros(C.sizeof(_lib.PosthocAccum), dtype=torch.uint8, device=logits.device)
    p_true = torch.empty(B, V, device=logits.device) if want_p_true else None
    check(lib.mmu_posthoc_scoring(ptr(logits), ptr(labels), V, B, E, Cn, n_repeats, ptr(p_true),
                                  ptr(accum), stream_ptr()), "mmu_posthoc_scoring")
    return accum, p_true


def pair_concordance(x, y):
    """Exact pair counts of (x, y): x (n,) or (batch, n), y likewise (a 1-D argument is shared by
    every problem of the batch).  Returns int64 (batch, 4) = concordant, discordant, tied in x,
    tied in y; see ``mmu_pair_concordance``."""
    _cuda(x, y)
    if x.dtype != torch.float32 or y.dtype != torch.float32:
        raise TypeError("x and y must be fp32")
    n = x.shape[-1]
    if y.shape[-1] != n:
        raise ValueError("x and y differ in length")
    batch = max(x.shape[0] if x.dim() == 2 else 1, y.shape[0] if y.dim() == 2 else 1)
    for t in (x, y):
        if t.dim() == 2 and t.shape[0] != batch:
            raise ValueError("batch sizes differ")
    counts = torch.empty(batch, 4, dtype=torch.int64, device=x.device)
    check(lib.mmu_pair_concordance(ptr(x), ptr(y), n, batch, n if x.dim() == 2 else 0,
                                   n if y.dim() == 2 else 0, ptr(counts), stream_ptr()),
          "mmu_pair_concordance")
    return counts


def top_truncate(pred, labels=None, top=5, mute_true=False):
    """``trunk_pred_top`` of notebooks/analysis_round_1.py:74-85 on device; pred fp32 (N, C)."""
    _cuda(pred, labels)
    if pred.dtype != torch.float32 or (labels is not None and labels.dtype != torch.int64):
        raise TypeError("pred must be fp32 and labels int64")
    N, Cn = pred.shape
    out = torch.empty_like(pred)
    check(lib.mmu_top_truncate(ptr(pred), ptr(labels), N, Cn, top, int(bool(mute_true)), ptr(out),
                               stream_ptr()), "mmu_top_truncate")
    return out


def ragged_pad(packed, offsets, max_len):
    """Zero-padded (B, max_len, d) batch from packed rows (sum_len, d) + int32 offsets (B+1,):
    the device twin of ``pad_sequence(batch_first=True)`` (reference src/dataset.py:216-226)."""
    _cuda(packed, offsets)
    if offsets.dtype != torch.int32 or packed.dtype != torch.float32:
        raise TypeError("packed must be fp32 and offsets int32")
    B = offsets.numel() - 1
    out = torch.empty(B, max_len, packed.shape[1], device=packed.device, dtype=torch.float32)
    check(lib.mmu_ragged_pad(ptr(packed), ptr(offsets), ptr(out), B, max_len, packed.shape[1],
                             stream_ptr()), "mmu_ragged_pad")
    return out


def modality_keep_mask(u, r, p_drop, mode="random", score_img=None, score_txt=None):
    """Keep mask int32 (B, 2) on device from host-drawn uniforms ``u``, ``r`` (fp32 (B,), already
    on the device) and -- ``mode='guided'`` -- device-resident per-sample scores (1-D views with
    any equal stride, e.g. ``scores[:, 0]`` of the epilogue's per-sample output); see
    ``mmu_modality_keep_mask``."""
    if not (u.is_cuda and u.dtype == torch.float32 and u.is_contiguous()):
        raise ValueError("u must be a contiguous fp32 CUDA vector")
    B = u.numel()
    m = {"random": 0, "guided": 1}[mode]
    stride = 1
    if m == 1:
        if score_img.dtype != torch.float32 or score_txt.dtype != torch.float32 or \
                score_img.stride(0) != score_txt.stride(0) or score_img.numel() != B or score_txt.numel() != B:
            raise ValueError("guided scores must be fp32 (B,) views with equal strides")
        stride = score_img.stride(0)
    keep = torch.empty(B, 2, dtype=torch.int32, device=u.device)
    check(lib.mmu_modality_keep_mask(ptr(u), ptr(r), ptr(score_img), ptr(score_txt), stride, B,
                                     float(p_drop), m, ptr(keep), stream_ptr()), "mmu_modality_keep_mask")
    return keep
