"""Oracle: uncertainty scores, calibration histograms and post-hoc robustness scoring
(TEST INFRASTRUCTURE).

PARITY UNPINNED for everything in the first half of this file: the reference contains no
entropy / mutual-information / ECE / histogram code at all.  The
definitions below are the repo's own, the textbook ones, evaluated in fp64.  The second half
restates the notebooks' scoring (notebooks/utils.py:22-34, notebooks/food101_robustness.py:24-77),
which IS reference code.
"""
import math

import numpy as np
import torch

CONF_BINS = 15
SCORE_BINS = 32


def _f64(t):
    return t.detach().to(torch.float64)


def ensemble_scores(logits):
    """logits (N, E, C) -> dict of per-sample fp64 scores.

    p_k = softmax(z_k); pbar = mean_k p_k; H_pred = -sum pbar log pbar;
    H_exp = mean_k(-sum p_k log p_k); MI = H_pred - H_exp; conf = max pbar;
    pred_prob = argmax pbar (first index on ties); pred_logit = argmax mean_k z_k, the
    reference's own eval convention (src/model.py:302, train.py:126)."""
    z = _f64(logits)
    z = z - z.max(dim=-1, keepdim=True).values
    logp = z - torch.log(torch.exp(z).sum(dim=-1, keepdim=True))
    p = torch.exp(logp)
    h_k = -(p * logp).sum(dim=-1)  # (N, E)
    pbar = p.mean(dim=1)
    h_pred = -(pbar * torch.log(pbar.clamp_min(1e-300))).sum(dim=-1)
    h_exp = h_k.mean(dim=1)
    conf, pred_prob = pbar.max(dim=-1)
    return {
        "pbar": pbar,
        "h_pred": h_pred,
        "h_exp": h_exp,
        "mi": h_pred - h_exp,
        "conf": conf,
        "pred_prob": pred_prob,
        "pred_logit": _f64(logits).mean(dim=1).argmax(dim=-1),
    }


def bin_index(value, nbins, scale=1.0):
    """Equal-width bin of value/scale in [0, 1]: min(floor(v * nbins), nbins-1), negatives -> 0.
    Evaluated in fp32, the arithmetic the CUDA epilogue uses."""
    v = (value.to(torch.float32) * np.float32(1.0 / scale)) * np.float32(nbins)
    return v.floor().clamp(0, nbins - 1).to(torch.int64)


def calibration_histograms(logits, labels, conf_bins=CONF_BINS, score_bins=SCORE_BINS):
    """Per-batch accumulators the fused epilogue produces.

    Returns dict with integer tensors ``conf_count``/``conf_correct`` (conf_bins,),
    ``hpred_count``/``mi_count`` (score_bins,), fp64 ``conf_sum`` (conf_bins,), and scalar sums.
    H_pred is binned after division by log C, MI after division by log max(E, 2)."""
    s = ensemble_scores(logits)
    N, E, C = logits.shape
    y = labels.reshape(-1).to(torch.int64)
    correct = (s["pred_prob"] == y)
    cb = bin_index(s["conf"], conf_bins)
    hb = bin_index(s["h_pred"], score_bins, scale=math.log(C))
    mb = bin_index(s["mi"], score_bins, scale=math.log(max(E, 2)))
    out = {
        "conf_count": torch.bincount(cb, minlength=conf_bins),
        "conf_correct": torch.bincount(cb, weights=correct.to(torch.float64),
                                       minlength=conf_bins).to(torch.int64),
        "conf_sum": torch.bincount(cb, weights=s["conf"], minlength=conf_bins),
        "hpred_count": torch.bincount(hb, minlength=score_bins),
        "mi_count": torch.bincount(mb, minlength=score_bins),
        "n": N,
        "n_correct_prob": int(correct.sum()),
        "n_correct_logit": int((s["pred_logit"] == y).sum()),
        "sum_h_pred": float(s["h_pred"].sum()),
        "sum_h_exp": float(s["h_exp"].sum()),
        "sum_mi": float(s["mi"].sum()),
    }
    return out


def ece_from_bins(conf_count, conf_correct, conf_sum):
    """Expected calibration error: sum_b (n_b / N) |acc_b - conf_b|."""
    n = conf_count.to(torch.float64)
    N = n.sum().clamp_min(1.0)
    nz = n > 0
    acc_b = conf_correct.to(torch.float64)[nz] / n[nz]
    conf_b = conf_sum.to(torch.float64)[nz] / n[nz]
    return float(((n[nz] / N) * (acc_b - conf_b).abs()).sum())


# ------------------------------------------------ reference post-hoc scoring (notebooks)
def notebook_softmax(x):
    """notebooks/utils.py:22-23 (no max subtraction, numpy)."""
    x = np.asarray(x)
    return np.exp(x) / np.exp(x).sum(-1, keepdims=True)


def process_predictions(predictions, labels, n_repeats=20):
    """notebooks/food101_robustness.py:24-46 (non-MMBT branch): predictions (S, 3+2n, K, C).
    Returns p(true label) for: full, image-only, text-only, image controls (S, n), text
    controls (S, n), each after averaging PROBABILITIES over heads."""
    p = notebook_softmax(predictions).mean(2)  # (S, V, C)
    idx = np.arange(len(labels))
    pt = p[idx, :, labels]  # (S, V)
    return pt[:, 0], pt[:, 1], pt[:, 2], pt[:, 3:3 + n_repeats], pt[:, 3 + n_repeats:]


def pearson(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xm, ym = x - x.mean(), y - y.mean()
    return float((xm * ym).sum() / math.sqrt((xm * xm).sum() * (ym * ym).sum()))


def get_correlation(ori, image, text, image_corr, text_corr):
    """notebooks/utils.py:26-34: Pearson r between the experimental delta-p and the mean control
    delta-p, per modality."""
    return {
        "image": pearson(image - ori, (image_corr - ori[:, None]).mean(1)),
        "text": pearson(text - ori, (text_corr - ori[:, None]).mean(1)),
    }


def acc_table(predictions, labels, n_repeats=20):
    """notebooks/food101_robustness.py:48-77 (non-MMBT): accuracy of the head-mean LOGITS per
    variant; controls averaged over repeats.  Returns dict of percentages / fractions exactly as
    the notebook mixes them (full/image/text in %, control rows as fractions)."""
    pred = predictions.mean(2).argmax(-1)  # (S, V)
    lab = np.asarray(labels)
    out = {
        "full": float((pred[:, 0] == lab).mean() * 100),
        "image": float((pred[:, 1] == lab).mean() * 100),
        "text": float((pred[:, 2] == lab).mean() * 100),
        "image_control": (pred[:, 3:3 + n_repeats] == lab[:, None]).mean(-1),
        "text_control": (pred[:, 3 + n_repeats:] == lab[:, None]).mean(-1),
    }
    return out
