"""CPU restatement of the rating path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* metrics: reference src/metrics/rating.py:4-29 (numpy float64) -- PINNED against the reference module run live
  (tests/golden/rating_golden.json, made by oracle/gen_golden.py rating).
* MF: reference src/models/basic/models/mf.py:54-83 -- the step is ``oracle.steps.wrmf_step`` with weight = 1 (mf.py's
  loss is wrmf.py's with a unit weight); prediction + clip as mf.py:66-72,81.  PARITY UNPINNED against TensorFlow itself
  (same status as oracle.steps).
"""
import numpy as np

from . import steps


def mean_absolute_error(ys_true, ys_pred):      # rating.py:4-6
    ys_true, ys_pred = np.asarray(ys_true, dtype=np.float64), np.asarray(ys_pred)
    return 1 / ys_true.shape[0] * np.sum(np.fabs(ys_true - ys_pred))


def mean_squared_error(ys_true, ys_pred):       # rating.py:9-11
    ys_true, ys_pred = np.asarray(ys_true, dtype=np.float64), np.asarray(ys_pred)
    return 1 / ys_true.shape[0] * np.sum(np.power(ys_true - ys_pred, 2))


def root_mean_squared_error(ys_true, ys_pred):  # rating.py:14-16
    return np.sqrt(mean_squared_error(ys_true, ys_pred))


def evaluate(ys_true, ys_pred, eval_metrics):   # rating.py:18-29
    fn = dict(mae=mean_absolute_error, mse=mean_squared_error, rmse=root_mean_squared_error)
    return [(fn[m](ys_true, ys_pred) if m in fn else None) for m in eval_metrics]


def mf_predict(U, V, useritem, range_of_ratings=None):
    """mf.py:66-72 (+ the clip of :81 when a range is given): float32 predictions of the (user, item) rows."""
    u, i = useritem[:, 0].astype(np.int64), useritem[:, 1].astype(np.int64)
    p = np.sum(U[u].astype(np.float64) * V[i].astype(np.float64), axis=1).astype(np.float32)
    if range_of_ratings is not None:
        p = np.clip(p, np.float32(range_of_ratings[0]), np.float32(range_of_ratings[1]))
    return p


def mf_train(U, V, accU, accV, tra_tuple, tst_tuple, eval_metrics, range_of_ratings, reg, batch_size, max_iter, lr=0.1):
    """mf.py:86-110 fed by sampler_rating with negRatio = 0 (testmf.py:44): the positives in file order, batch after
    batch, the tail dropped; no randomness at all.  Returns the per-epoch (mean loss, scores)."""
    n_batches = int(len(tra_tuple) / batch_size)
    out = []
    for _ in range(max_iter):
        losses = []
        for k in range(n_batches):
            losses.append(steps.wrmf_step(U, V, accU, accV, tra_tuple[k * batch_size:(k + 1) * batch_size], lr, reg, 1.0))
        pred = mf_predict(U, V, tst_tuple[:, :2], range_of_ratings)
        out.append((float(np.mean(losses)), evaluate(tst_tuple[:, 2], pred, eval_metrics)))
    return out


def svd_predict(U, V, K, useritem, range_of_ratings=None):
    """svd.py:66-72 (+ the clip of :81): float32 predictions ``sum((U_u @ K) * V_i)`` of the (user, item) rows."""
    u, i = useritem[:, 0].astype(np.int64), useritem[:, 1].astype(np.int64)
    p = np.sum((U[u].astype(np.float64) @ K.astype(np.float64)) * V[i].astype(np.float64), axis=1).astype(np.float32)
    if range_of_ratings is not None:
        p = np.clip(p, np.float32(range_of_ratings[0]), np.float32(range_of_ratings[1]))
    return p
