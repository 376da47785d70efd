"""Oracle: rank statistics of the reference's post-hoc analysis (TEST INFRASTRUCTURE -- only
tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this).

Restates, in plain numpy integer arithmetic:
  * `trunk_pred_top` and `subnetwork_wise_kendalltau` of notebooks/analysis_round_1.py:74-90
    (head-diversity score: scipy.stats.kendalltau of the flattened, top-5-truncated,
    true-class-muted predictions of every pair of heads, printed as the mean at :111-113);
  * the AUROC of src/framework.py:195-198 (`roc_auc_score(labels, preds[:, 1])`) and of
    notebooks/hatefulmeme_robustness.py:22-41 (`AUC_table`: full / image / text / 20 + 20 controls).

The arithmetic of both lives in third-party packages the reference imports unpinned
(`scipy.stats.kendalltau`, requirements.txt: scipy; `sklearn.metrics.roc_auc_score`,
requirements.txt: scikit_learn).  Both are installed in this image (scipy 1.x, scikit-learn 1.9), so
the restatement is PINNED twice: tests/test_oracle_golden.py holds it to goldens produced by
executing the reference's own notebook functions (tests/golden/make_golden.py::rank_case) and, live,
to scipy / sklearn on random vectors with heavy ties.

Published definitions restated here:
  tau-b = (P - Q) / sqrt((T - Tx)(T - Ty)),  T = n(n-1)/2, P / Q concordant / discordant pairs,
          Tx / Ty pairs tied in x / in y (joint ties counted in both)         [Kendall 1945]
  AUROC = (P + (Ty - Txy)/2) / (P + Q + Ty - Txy) with x the 0/1 labels: the Mann-Whitney U
          statistic, which is what the trapezoid rule over sklearn's ROC curve evaluates to.
"""
import itertools
import math

import numpy as np


def pair_counts(x, y, block=2048):
    """Exact counts over unordered pairs {i, j}: (concordant, discordant, tied in x, tied in y);
    both tie counts include jointly tied pairs.  O(n^2) in blocks; NaNs compare as ties."""
    x = np.asarray(x, dtype=np.float32).ravel()
    y = np.asarray(y, dtype=np.float32).ravel()
    n = x.size
    conc = disc = tx = ty = 0
    for i0 in range(0, n, block):
        xi, yi = x[i0:i0 + block, None], y[i0:i0 + block, None]
        ii = np.arange(i0, min(i0 + block, n))[:, None]
        for j0 in range(i0, n, block):
            xj, yj = x[None, j0:j0 + block], y[None, j0:j0 + block]
            jj = np.arange(j0, min(j0 + block, n))[None, :]
            a = (xi > xj).astype(np.int8) - (xi < xj).astype(np.int8)
            b = (yi > yj).astype(np.int8) - (yi < yj).astype(np.int8)
            up = jj > ii
            p = a * b
            conc += int(np.count_nonzero((p > 0) & up))
            disc += int(np.count_nonzero((p < 0) & up))
            tx += int(np.count_nonzero((a == 0) & up))
            ty += int(np.count_nonzero((b == 0) & up))
    return conc, disc, tx, ty


def tau_b_from_counts(conc, disc, tx, ty, n):
    """scipy.stats.kendalltau(variant='b') as a function of the pair counts; nan when either
    vector is constant (scipy returns nan there too)."""
    tot = n * (n - 1) // 2
    if tot - tx == 0 or tot - ty == 0:
        return float("nan")
    return (conc - disc) / math.sqrt(tot - tx) / math.sqrt(tot - ty)


def auroc_from_counts(conc, disc, tx, ty, n):
    """roc_auc_score for binary labels in x (0/1 floats) and scores in y."""
    joint = conc + disc + tx + ty - n * (n - 1) // 2
    ty_only = ty - joint
    den = conc + disc + ty_only  # = n_pos * n_neg
    if den == 0:
        return float("nan")  # sklearn raises "Only one class present"
    return (conc + 0.5 * ty_only) / den


def kendalltau(x, y):
    x, y = np.asarray(x).ravel(), np.asarray(y).ravel()
    return tau_b_from_counts(*pair_counts(x, y), x.size)


def auroc(labels, scores):
    labels = np.asarray(labels).ravel()
    return auroc_from_counts(*pair_counts(labels.astype(np.float32), scores), labels.size)


def trunk_pred_top(pred, labels, top, mute_true=False):
    """notebooks/analysis_round_1.py:74-85: per row keep the entries >= the top-th largest value
    of the ORIGINAL row (np.partition(row, -top)[-top]), after zeroing the true class when
    mute_true; everything else becomes 0."""
    pred = np.asarray(pred)
    out = pred.copy()
    if mute_true:
        out[np.arange(len(pred)), np.asarray(labels)] = 0
    value = np.sort(pred, axis=1)[:, -top][:, None]
    return np.where(out >= value, out, 0)


def subnetwork_wise_kendalltau(predictions, labels, top=5, mute_true=True):
    """notebooks/analysis_round_1.py:87-113: predictions (S, E, C) -> tau-b of every pair of heads
    (itertools.combinations order) on the flattened truncated arrays; the notebook reports the mean."""
    predictions = np.asarray(predictions)
    muted = [trunk_pred_top(predictions[:, k, :], labels, top, mute_true)
             for k in range(predictions.shape[1])]
    return np.array([kendalltau(a, b) for a, b in itertools.combinations(muted, 2)])


def auc_table(labels, scores):
    """notebooks/hatefulmeme_robustness.py:22-41 `AUC_table` on the (S, 3 + 2n) matrix of
    p(class 1): one AUROC per variant (full, image, text, image controls, text controls)."""
    scores = np.asarray(scores)
    return np.array([auroc(labels, scores[:, v]) for v in range(scores.shape[1])])
