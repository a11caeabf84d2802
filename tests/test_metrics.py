"""Metric layer (host): known answers, agreement with the oracle restatement of
utils/detection_util.py:47-119 and with sklearn (what the reference itself calls)."""
import numpy as np
import pytest

from mcm_b200 import metrics
from oracle import clip_mcm_oracle as O


def test_perfect_separation():
    pos, neg = np.array([3.0, 4.0, 5.0]), np.array([0.0, 1.0, 2.0])
    auroc, aupr, fpr = metrics.get_measures(pos, neg)
    assert auroc == 1.0 and aupr == 1.0 and fpr == 0.0


def test_inverted_and_ties():
    auroc, aupr, fpr = metrics.get_measures(np.array([0.0, 1.0]), np.array([2.0, 3.0]))
    assert auroc == 0.0 and fpr == 1.0
    auroc, _, _ = metrics.get_measures(np.array([1.0, 1.0]), np.array([1.0, 1.0]))
    assert auroc == 0.5
    # hand-computed: pos {0.9, 0.5, 0.5}, neg {0.5, 0.1}: wins = 2 + (0.5 + 1) * 2 = 5 of 6
    auroc, _, _ = metrics.get_measures(np.array([0.9, 0.5, 0.5]), np.array([0.5, 0.1]))
    assert auroc == pytest.approx(5.0 / 6.0)


def test_fpr_known_answer():
    # 20 positives 1..20, negatives at 0.5, 1.5, 2.5, 30.  Distinct thresholds 2 and 1.5 both recall
    # 19/20 = 0.95; the reference scans from high recall down and keeps the FIRST minimum
    # (utils/detection_util.py:100-106), i.e. the lower threshold 1.5, where 3 of 4 negatives pass.
    pos = np.arange(1, 21, dtype=np.float64)
    neg = np.array([0.5, 1.5, 2.5, 30.0])
    y = np.r_[np.ones(20), np.zeros(4)]
    fpr = metrics.fpr_and_fdr_at_recall(y, np.r_[pos, neg], 0.95)
    assert fpr == pytest.approx(3 / 4)
    assert fpr == pytest.approx(O.fpr_at_recall(y, np.r_[pos, neg], 0.95))


def test_stable_cumsum():
    a = np.full(1000, 0.1, dtype=np.float32)
    out = metrics.stable_cumsum(a)
    assert out.dtype == np.float64 and out[-1] == pytest.approx(100.0, rel=1e-6)


@pytest.mark.parametrize("seed", range(6))
def test_against_oracle_and_sklearn(seed):
    import sklearn.metrics as sk
    rng = np.random.default_rng(seed)
    n1, n2 = rng.integers(5, 400, 2)
    pos = rng.normal(0.4, 1.0, n1).astype(np.float32)
    neg = rng.normal(0.0, 1.0, n2).astype(np.float32)
    if seed % 2:
        pos, neg = np.round(pos, 1), np.round(neg, 1)     # heavy ties
    got = metrics.get_measures(pos, neg)
    ref = O.get_measures(pos, neg)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    y = np.r_[np.ones(n1), np.zeros(n2)]
    s = np.r_[pos, neg]
    assert got[0] == pytest.approx(sk.roc_auc_score(y, s), abs=1e-12)
    assert got[1] == pytest.approx(sk.average_precision_score(y, s), abs=1e-12)


def test_single_class_raises():
    with pytest.raises(ValueError):
        metrics.auroc(np.ones(4), np.arange(4.0))
