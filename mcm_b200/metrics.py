"""OOD detection metrics consumed by the hot path's callers (host side, numpy).

Same results as the reference's metric layer, ``utils/detection_util.py:47-119`` (which leans on
``sklearn.metrics.roc_auc_score`` / ``average_precision_score``), re-derived from the definitions
so the package needs nothing but numpy:

* AUROC   = P(score_pos > score_neg) + 0.5 P(tie), via sorted-unique counting;
* AUPR    = sum over distinct thresholds of (recall step) x precision;
* FPR@r   = false-positive rate at the distinct threshold whose recall is closest to ``r`` (ties
            resolved towards the higher recall, as ``fpr_and_fdr_at_recall`` does, :66-106).

Positives are the ID samples: callers pass ``-in_score, -out_score`` (``:259``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["stable_cumsum", "fpr_and_fdr_at_recall", "get_measures", "auroc", "aupr"]


def stable_cumsum(arr, rtol=1e-05, atol=1e-08):
    """float64 cumulative sum with a consistency check against the plain sum (``:47-63``)."""
    a = np.asarray(arr)
    out = np.cumsum(a, dtype=np.float64)
    if out.size and not np.isclose(out[-1], np.sum(a, dtype=np.float64), rtol=rtol, atol=atol):
        raise RuntimeError("cumsum was found to be unstable: its last element does not correspond to sum")
    return out


def _threshold_counts(y_true, y_score):
    """Descending distinct thresholds -> (true positives, false positives) at each."""
    y_true = np.asarray(y_true).astype(bool).ravel()
    y_score = np.asarray(y_score).ravel()
    order = np.argsort(y_score, kind="mergesort")[::-1]
    s = y_score[order]
    t = y_true[order]
    ends = np.flatnonzero(s[1:] != s[:-1])          # last index of every run of equal scores
    ends = np.append(ends, t.size - 1)
    tps = stable_cumsum(t)[ends]
    fps = (ends + 1) - tps
    return tps, fps


def fpr_and_fdr_at_recall(y_true, y_score, recall_level=0.95, pos_label=None):
    """FPR at the operating point whose recall is nearest ``recall_level`` (``:66-106``)."""
    y_true = np.asarray(y_true)
    if pos_label is None:
        classes = np.unique(y_true)
        ok = (classes.size <= 2) and set(classes.tolist()) <= {0, 1, -1, True, False}
        if not ok:
            raise ValueError("Data is not binary and pos_label is not specified")
        pos_label = 1.0
    pos = (y_true == pos_label)
    tps, fps = _threshold_counts(pos, y_score)
    n_pos = tps[-1]
    n_neg = pos.size - int(pos.sum())
    # thresholds beyond the first one that already recalls every positive add nothing
    last = int(np.searchsorted(tps, n_pos))
    recall = tps[: last + 1] / n_pos
    # scan from high recall to low; first minimum wins -> ties go to the higher recall
    rev = recall[::-1]
    cut = int(np.argmin(np.abs(rev - recall_level)))
    return fps[: last + 1][::-1][cut] / n_neg


def auroc(labels, scores) -> float:
    """Area under the ROC curve (Mann-Whitney statistic with half credit for ties)."""
    labels = np.asarray(labels).astype(bool).ravel()
    scores = np.asarray(scores, dtype=np.float64).ravel()
    pos = np.sort(scores[labels])
    neg = np.sort(scores[~labels])
    if pos.size == 0 or neg.size == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    below = np.searchsorted(neg, pos, side="left")       # negatives strictly below each positive
    upto = np.searchsorted(neg, pos, side="right")       # negatives <= each positive
    wins = below.sum(dtype=np.float64) + 0.5 * (upto - below).sum(dtype=np.float64)
    return float(wins / (pos.size * neg.size))


def aupr(labels, scores) -> float:
    """Average precision: sum_n (R_n - R_{n-1}) P_n over distinct thresholds."""
    tps, fps = _threshold_counts(labels, scores)
    if tps[-1] == 0:
        return 0.0
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.concatenate(([0.0], recall))) * precision))


def get_measures(_pos, _neg, recall_level=0.95):
    """``(auroc, aupr, fpr)`` with ``_pos`` the positive (ID) scores (``:108-119``)."""
    pos = np.asarray(_pos[:]).reshape(-1)
    neg = np.asarray(_neg[:]).reshape(-1)
    examples = np.concatenate((pos, neg))
    labels = np.zeros(examples.size, dtype=np.int32)
    labels[: pos.size] = 1
    return auroc(labels, examples), aupr(labels, examples), fpr_and_fdr_at_recall(labels, examples, recall_level)
