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
    owner = getattr(y_pred, "_mmu_owner", None)   # set by model.forward_backward (no autograd graph)
    if owner is not None:
        return owner
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
    labels = y_true.reshape(y_pred.shape[0], -1) if mode == 0 else y_true.reshape(-1)
    # the loss epilogue already counted the correct rows of exactly these logits AND labels
    accum = model.cached_epilogue(y_pred, mode, labels) if model is not None else None
    if accum is None:
        _, _, _, accum = ops.heads_uncertainty_epilogue(y_pred, labels.contiguous(), mode)
    o = _lib.ACC_OFF
    correct = accum[o["n_correct_rows"]].to(torch.float32)
    rows = accum[o["n_rows"]].to(torch.float32)
    return correct / rows * 100


class AsyncScalars:
    """Device -> host read of per-step scalars (loss, metrics) that does NOT drain the work queued
    behind them.  ``tensor.item()`` / ``float(tensor)`` copies on the CURRENT stream, i.e. behind
    every kernel already enqueued -- with one-step-lagged logging the host then waits for the step
    it has just launched, and the GPU idles while the next step is being prepared (0.37 ms per
    13.8 ms step, ``tools/e2e_probe.py``).  ``push`` records an event behind the producing step,
    copies on a side stream into pinned memory and returns a ticket; ``pop`` waits for that copy
    only.  Replaces the two per-step ``.item()`` reads of reference ``src/framework.py:305-312``
    for a logger that may lag by a step."""

    def __init__(self, device, slots=4, width=8):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.host = [torch.empty(width, dtype=torch.float32).pin_memory() for _ in range(slots)]
        self.stage = [torch.empty(width, dtype=torch.float32, device=self.device) for _ in range(slots)]
        self.done = [None] * slots
        self.n = [0] * slots
        self.k = -1

    def push(self, scalars):
        """``scalars``: 0-dim device tensors.  Returns a ticket for :meth:`pop`."""
        self.k = (self.k + 1) % len(self.host)
        k = self.k
        if self.done[k] is not None:
            self.done[k].synchronize()          # the slot's previous copy has landed
        vals = torch.stack([s.detach().reshape(()).float() for s in scalars])
        self.n[k] = vals.numel()
        self.stage[k][:vals.numel()].copy_(vals)   # on the producing stream, behind the step
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self.host[k].copy_(self.stage[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self.done[k] = ev
        return k

    def pop(self, ticket):
        self.done[ticket].synchronize()
        return self.host[ticket][:self.n[ticket]].tolist()


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

    def __init__(self, device, n_repeats=20, auc=False):
        """``auc=True`` (binary heads: hateful memes) also keeps every sample's head-mean p(class 1)
        per variant on the device -- 4 bytes per sample-variant -- so that ``compute()`` can rank
        them: the ``AUC_table`` of reference ``notebooks/hatefulmeme_robustness.py:22-41``."""
        import ctypes as C
        self.n_repeats = n_repeats
        self.accum = torch.zeros(C.sizeof(_lib.PosthocAccum), dtype=torch.uint8, device=device)
        self.auc = auc
        self._p1, self._labels = [], []

    def reset(self):
        self.accum.zero_()
        self._p1, self._labels = [], []

    def update(self, logits_vbec, labels, want_p_true=False):
        logits_vbec, labels = logits_vbec.contiguous(), labels.reshape(-1).contiguous()
        _, p_true = ops.posthoc_scoring(logits_vbec, labels, self.n_repeats, accum=self.accum,
                                        want_p_true=want_p_true)
        if self.auc:
            if logits_vbec.shape[-1] != 2:
                raise ValueError("the AUROC table is defined for binary heads (C = 2)")
            self._p1.append(self.class_prob(logits_vbec, 1))
            self._labels.append(labels)
        return p_true

    def class_prob(self, logits_vbec, cls=1):
        """Head-mean probability of class ``cls`` per (sample, variant), (B, V): the scores of
        ``process_predictions_hatefulmeme`` (reference ``notebooks/hatefulmeme_robustness.py:
        105-112``) that ``metrics.auc_table`` ranks.  Same kernel as ``update`` with every label
        set to ``cls``; the accumulator of this meter is not touched."""
        B = logits_vbec.shape[1]
        lab = torch.full((B,), int(cls), dtype=torch.int64, device=logits_vbec.device)
        _, p = ops.posthoc_scoring(logits_vbec.contiguous(), lab, self.n_repeats, accum=None,
                                   want_p_true=True)
        return p

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
            if self.auc and self._p1:   # rank statistics need every sample: gather the scores
                from .parallel import all_gather_rows
                self._p1 = [all_gather_rows(torch.cat(self._p1), group)]
                self._labels = [all_gather_rows(torch.cat(self._labels), group)]

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
        if self.auc and self._p1:
            tab = auc_table(torch.cat(self._labels), torch.cat(self._p1))
            out.update(auc_per_variant=tab["AUC"], auc_full=tab["full"], auc_image=tab["image"],
                       auc_text=tab["text"], auc_image_control=tab["image_control"],
                       auc_text_control=tab["text_control"])
        return out


# ------------------------------------------------------------------------------------------------
# Rank statistics of the post-hoc analysis, from exact device-side pair counts
# (``mmu_pair_concordance``).  Nothing is sorted and nothing leaves the GPU but 32 bytes per problem.
def _as_f32_rows(t, device):
    t = torch.as_tensor(t)
    return t.to(device=device, dtype=torch.float32).contiguous()


def pair_counts(x, y):
    """int64 numpy (batch, 4): concordant, discordant, tied in x, tied in y (joint ties in both)."""
    if not torch.cuda.is_available():
        raise _lib.MMUError("rank statistics run on the GPU only (no CUDA device, no CPU fallback)")
    dev = y.device if isinstance(y, torch.Tensor) and y.is_cuda else torch.device("cuda")
    return ops.pair_concordance(_as_f32_rows(x, dev), _as_f32_rows(y, dev)).cpu().numpy()


def auroc(labels, scores):
    """Area under the ROC curve of binary ``labels`` (N,) against ``scores`` (N,) or (V, N): what
    ``sklearn.metrics.roc_auc_score`` returns at reference ``src/framework.py:195-198`` and
    ``notebooks/hatefulmeme_robustness.py:22-41``, as the Mann-Whitney ratio of exact pair counts.
    Returns a float (1-D scores) or a float64 array (V,).  Raises ValueError when only one class
    is present, like sklearn."""
    scores_t = torch.as_tensor(scores)
    n = scores_t.shape[-1]
    cnt = pair_counts(torch.as_tensor(labels).reshape(-1), scores_t)
    conc, disc, tx, ty = (cnt[:, k].astype(object) for k in range(4))   # Python ints: no overflow
    joint = conc + disc + tx + ty - n * (n - 1) // 2
    ty_only = ty - joint
    den = conc + disc + ty_only
    if (den == 0).any():
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    out = np.array([float((c + 0.5 * t) / d) for c, t, d in zip(conc, ty_only, den)])
    return float(out[0]) if scores_t.dim() == 1 else out


def kendalltau(x, y):
    """Kendall tau-b of the flattened inputs (``scipy.stats.kendalltau`` as called at reference
    ``notebooks/analysis_round_1.py:87-90``); nan when one side is constant."""
    x, y = torch.as_tensor(x).reshape(-1), torch.as_tensor(y).reshape(-1)
    n = x.numel()
    conc, disc, tx, ty = (int(v) for v in pair_counts(x, y)[0])
    tot = n * (n - 1) // 2
    if tot - tx == 0 or tot - ty == 0:
        return float("nan")
    return (conc - disc) / np.sqrt(float(tot - tx)) / np.sqrt(float(tot - ty))


def head_diversity_kendalltau(predictions, labels, top=5, mute_true=True):
    """Prediction-diversity score of reference ``notebooks/analysis_round_1.py:74-113``: per head,
    ``trunk_pred_top`` (keep the top-``top`` entries of each row, true class muted), then tau-b
    between every pair of heads (``itertools.combinations`` order) on the flattened arrays.
    predictions: CUDA fp32 (S, E, C); labels int64 (S).  Returns float64 (E(E-1)/2,); the notebook
    prints the mean."""
    import itertools
    S, E, C = predictions.shape
    labels = labels.reshape(-1).to(predictions.device, torch.int64).contiguous()
    muted = [ops.top_truncate(predictions[:, k, :].contiguous().float(), labels, top, mute_true)
             for k in range(E)]
    return np.array([kendalltau(a, b) for a, b in itertools.combinations(muted, 2)])


def auc_table(labels, scores):
    """``AUC_table`` of reference ``notebooks/hatefulmeme_robustness.py:22-41``: scores (S, V) of
    p(class 1) per variant (full, image, text, n image controls, n text controls) -> AUROC per
    variant (V,) plus the notebook's group means."""
    scores = torch.as_tensor(scores)
    auc = auroc(labels, scores.t().contiguous())
    n = (len(auc) - 3) // 2
    return {"AUC": auc, "full": auc[0], "image": auc[1], "text": auc[2],
            "image_control": float(auc[3:3 + n].mean()) if n else float("nan"),
            "text_control": float(auc[3 + n:].mean()) if n else float("nan")}
