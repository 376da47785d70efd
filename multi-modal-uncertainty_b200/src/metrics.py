"""Metrics of the hot path, computed by the fused CUDA epilogue.

``acc`` keeps the reference's metric protocol ``metric(y_pred, y_true, eval, dummy_dim)``
(reference ``train.py:119-130``, called from ``src/framework.py:111-112``).  ``UncertaintyMeter``
accumulates the north-star extension scores (predictive entropy, expected entropy, mutual
information, ECE / confidence histograms) on device across batches and ranks.
"""
import numpy as np
import torch

from ._backend import _lib, ops


def _owner_of(y_pred):
    fn = getattr(y_pred, "grad_fn", None)
    return getattr(fn, "model", None) if fn is not None else None


def acc(y_pred, y_true, eval, dummy_dim=False, model=None):
    """Accuracy in percent.  With ``dummy_dim`` (logits (B, E, C)): train -> over the (B*E) head
    rows against the tiled labels; eval -> argmax of the head-mean logits.  Without it the input
    is (B, C).  First-index argmax on ties, like ``Tensor.max(1)``."""
    if not dummy_dim:
        y_pred = y_pred.unsqueeze(1)
        mode = 1
    else:
        mode = 1 if eval else 0
    if model is None:
        model = _owner_of(y_pred)
    y_pred = y_pred.detach()
    if not y_pred.is_contiguous():
        y_pred = y_pred.contiguous()
    accum = model.cached_epilogue(y_pred, mode) if model is not None else None
    if accum is None:
        labels = y_true.reshape(y_pred.shape[0], -1) if mode == 0 else y_true.reshape(-1)
        _, _, _, accum = ops.heads_uncertainty_epilogue(y_pred, labels.contiguous(), mode)
    o = _lib.ACC_OFF
    correct = accum[o["n_correct_rows"]].to(torch.float32)
    rows = accum[o["n_rows"]].to(torch.float32)
    return correct / rows * 100


class UncertaintyMeter:
    """Device-side accumulator of the fused epilogue's metrics over a sweep.

    ``update(logits, labels)`` enqueues one kernel; nothing is read back until ``compute()``
    (one D2H copy of 936 bytes).  ``all_reduce()`` sums the accumulators across ranks: integer
    bins are order-independent, hence bit-exact."""

    def __init__(self, device, num_classes, num_heads, mode="eval"):
        self.accum = ops.new_accum(device)
        self.C, self.E = num_classes, num_heads
        self.mode = 1 if mode == "eval" else 0

    def reset(self):
        self.accum.zero_()

    def update(self, logits, labels, want_scores=False):
        labels = labels.reshape(-1) if self.mode == 1 else labels.reshape(logits.shape[0], -1)
        _, pred, scores, _ = ops.heads_uncertainty_epilogue(
            logits.detach().contiguous(), labels.contiguous(), self.mode, accum=self.accum,
            want_pred=want_scores, want_scores=want_scores)
        return pred, scores

    def all_reduce(self, group=None):
        from . import parallel
        parallel.all_reduce_accum(self.accum, group)

    def compute(self):
        d = ops.accum_to_dict(self.accum)
        n = max(d["n_samples"], 1)
        cnt = d["conf_count"].astype(np.float64)
        nz = cnt > 0
        ece = float(np.sum(cnt[nz] / cnt.sum() * np.abs(d["conf_correct"][nz] / cnt[nz] -
                                                      d["conf_sum"][nz] / cnt[nz]))) if nz.any() else 0.0
        d.update(loss=d["loss_sum"] / max(d["n_rows"], 1),
                 acc=100.0 * d["n_correct_rows"] / max(d["n_rows"], 1),
                 acc_prob=100.0 * d["n_correct_prob"] / n, ece=ece,
                 h_pred=d["sum_h_pred"] / n, h_exp=d["sum_h_exp"] / n, mi=d["sum_mi"] / n)
        return d


class PosthocMeter:
    """Device-side accumulator of the notebooks' post-hoc robustness scores (reference
    ``notebooks/utils.py:22-34``, ``notebooks/food101_robustness.py:24-77``): fed with the
    packed-variant logits of each batch, it keeps the Pearson sufficient statistics and the
    per-variant accuracy counts on the GPU; ``compute()`` reads 1.1 KB back."""

    def __init__(self, device, n_repeats=20):
        import ctypes as C
        self.n_repeats = n_repeats
        self.accum = torch.zeros(C.sizeof(_lib.PosthocAccum), dtype=torch.uint8, device=device)

    def reset(self):
        self.accum.zero_()

    def update(self, logits_vbec, labels, want_p_true=False):
        _, p_true = ops.posthoc_scoring(logits_vbec.contiguous(), labels.reshape(-1).contiguous(),
                                        self.n_repeats, accum=self.accum, want_p_true=want_p_true)
        return p_true

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            words = self.accum.view(torch.int64)
            dbl = words[:10].view(torch.float64).clone()
            cnt = words[10:].clone()
            dist.all_reduce(dbl, group=group)
            dist.all_reduce(cnt, group=group)
            words[:10] = dbl.view(torch.int64)
            words[10:] = cnt

    def compute(self):
        import ctypes as C
        raw = self.accum.cpu().numpy().tobytes()
        a = _lib.PosthocAccum.from_buffer_copy(raw)
        n = max(int(a.n_samples), 1)
        out = {"n_samples": int(a.n_samples)}
        for i, name in enumerate(("image", "text")):
            cov = a.sxy[i] - a.sx[i] * a.sy[i] / n
            vx = a.sxx[i] - a.sx[i] ** 2 / n
            vy = a.syy[i] - a.sy[i] ** 2 / n
            out["corr_" + name] = float(cov / np.sqrt(vx * vy)) if vx > 0 and vy > 0 else float("nan")
        r = self.n_repeats
        acc = np.array([a.correct[v] for v in range(3 + 2 * r)], dtype=np.float64) / n
        out.update(acc_full=100 * acc[0], acc_image=100 * acc[1], acc_text=100 * acc[2],
                   acc_image_control=float(acc[3:3 + r].mean()) if r else float("nan"),
                   acc_text_control=float(acc[3 + r:].mean()) if r else float("nan"),
                   acc_per_variant=acc)
        return out
