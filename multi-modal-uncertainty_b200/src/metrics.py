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
