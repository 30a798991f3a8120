"""CPU oracle: ranking metrics.  TEST INFRASTRUCTURE.

Restates /root/reference/src/metrics/ranking.py:11-120 including its quirks:
NDCG's "ideal" DCG sorts the predicted list's own labels (:34-39), MAP divides by |truth| (:53),
HR/ARHR are sums over users, not means (:81,:91), unknown metric names yield None (:107,:118).

PINNED: tests/test_oracle_ranking.py checks it against the golden vectors captured from the
reference (tests/golden/ranking_golden.json) and, when /root/reference is present, against the
reference functions run live on random inputs.
"""
import numpy as np


def _check(a, b, k):
    if len(a) != len(b) or len(a) == 0 or k <= 0:
        raise ValueError('len(yss_true) != len(yss_pred) or len(yss_true)==0 or k<=0!')


def _hits(truth, pred, k):
    head = list(pred[:k])
    return head, np.array([p in truth for p in head], dtype=bool)


def per_user_cv(truth, pred, k):
    """(pre, recall, ndcg, map, mrr) for one user -- the summands of ranking.py:11-67."""
    head, hit = _hits(truth, pred, k)
    n_common = len(set(head) & set(truth))                       # set semantics as in :16,:25
    pre = n_common / float(k)
    rec = n_common / max(float(len(truth)), 1.0)
    disc = 1.0 / np.log2(np.arange(len(head)) + 2.0)
    gains = hit.astype(np.float64)                               # 2**label - 1
    dcg = float(np.sum(gains * disc)) if len(head) else 0.0
    ideal = float(np.sum(np.sort(gains)[::-1] * disc)) if len(head) else 0.0
    ndcg = dcg / max(ideal, 1.0)
    pos = np.flatnonzero(hit)
    ap = float(np.sum((np.arange(len(pos)) + 1.0) / (pos + 1.0))) / len(truth) if len(pos) else (
        0.0 / len(truth))                                        # ZeroDivisionError on empty truth, like :53
    rr = 1.0 / (pos[0] + 1.0) if len(pos) else 0.0
    return pre, rec, ndcg, ap, rr


_CV_COL = {'pre': 0, 'recall': 1, 'ndcg': 2, 'map': 3, 'mrr': 4}


def _cv_score(yss_true, yss_pred, k, col):
    _check(yss_true, yss_pred, k)
    tot = 0.0
    for t, p in zip(yss_true, yss_pred):
        tot += per_user_cv(t, p, k)[col]
    return tot / len(yss_true)


def precision_k_score(yss_true, yss_pred, k=5):
    return _cv_score(yss_true, yss_pred, k, 0)


def recall_k_score(yss_true, yss_pred, k=5):
    return _cv_score(yss_true, yss_pred, k, 1)


def ndcg_k_score(yss_true, yss_pred, k=5):
    return _cv_score(yss_true, yss_pred, k, 2)


def map_k_score(yss_true, yss_pred, k=5):
    return _cv_score(yss_true, yss_pred, k, 3)


def mrr_k_score(yss_true, yss_pred, k=5):
    return _cv_score(yss_true, yss_pred, k, 4)


def hr_k_score(ys_true, yss_pred, k=5):
    _check(ys_true, yss_pred, k)
    return float(sum(1.0 for y, p in zip(ys_true, yss_pred) if y in list(p[:k])))


def arhr_k_score(ys_true, yss_pred, k=5):
    _check(ys_true, yss_pred, k)
    tot = 0.0
    for y, p in zip(ys_true, yss_pred):
        head = list(p[:k])
        if y in head:
            tot += 1.0 / (head.index(y) + 1)
    return tot


def evaluateCV(yss_true, yss_pred, eval_metrics, k=5):
    return [(_cv_score(yss_true, yss_pred, k, _CV_COL[m]) if m in _CV_COL else None) for m in eval_metrics]


def evaluateLOOV(ys_true, yss_pred, eval_metrics, k=5):
    fn = {'hr': hr_k_score, 'arhr': arhr_k_score}
    return [(fn[m](ys_true, yss_pred, k) if m in fn else None) for m in eval_metrics]
