"""Ranking metrics with the reference's signatures and outputs (reference src/metrics/ranking.py:11-120), computed by
the ``cf_rank_metrics`` CUDA kernel (one thread per user); only the final mean / sum over users is done on the host,
exactly where ranking.py does it.  Quirks kept on purpose: NDCG's ideal DCG comes from the predicted
list's own labels, MAP divides by |truth|, HR / ARHR are sums, unknown metric names give ``None``.
"""
import numpy as np

from .. import _lib
from ..engine import rank_metrics_device
from ..sparse import DeviceCSR

_CV_COL = {'pre': 0, 'recall': 1, 'ndcg': 2, 'map': 3, 'mrr': 4}
_LOOV_COL = {'hr': 5, 'arhr': 6}
_ERR = 'len(yss_true) != len(yss_pred) or len(yss_true)==0 or k<=0!'


def _pred_tensor(yss_pred, k, device):
    torch = _lib.require_cuda()
    if torch.is_tensor(yss_pred):
        return yss_pred.to(device=device, dtype=torch.int32)[:, :max(1, k)].contiguous()
    width = max(1, min(k, max((len(p) for p in yss_pred), default=1)))
    arr = np.full((len(yss_pred), width), -1, dtype=np.int32)
    for t, p in enumerate(yss_pred):
        head = list(p[:k])
        arr[t, :len(head)] = head
    return torch.from_numpy(arr).to(device)


def _truth_csr(truth, device, loov=False):
    torch = _lib.require_cuda()
    if isinstance(truth, DeviceCSR):
        return truth
    if loov:
        rows = [[int(y)] for y in truth]
    else:
        rows = [sorted(int(y) for y in t) for t in truth]
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    flat = np.fromiter((y for r in rows for y in r), dtype=np.int32, count=int(indptr[-1]))
    if len(flat) == 0:
        flat = np.zeros(1, dtype=np.int32)
    return DeviceCSR(torch.from_numpy(indptr).to(device), torch.from_numpy(flat).to(device), None, None,
                     (len(rows), 1 << 30))


def per_user(yss_true, yss_pred, k, loov=False, device=None):
    """[T, 8] float64 CUDA tensor of per-user {pre, recall, ndcg, map, mrr, hit, rr, n_pred}."""
    torch = _lib.require_cuda()
    if len(yss_true) != len(yss_pred) or len(yss_true) == 0 or k <= 0:
        raise ValueError(_ERR)
    if device is None:
        device = yss_pred.device if torch.is_tensor(yss_pred) else torch.device('cuda', torch.cuda.current_device())
    truth = _truth_csr(yss_true, device, loov)
    return rank_metrics_device(_pred_tensor(yss_pred, k, device), truth, k), truth


def _cv(yss_true, yss_pred, k, name):
    vals, truth = per_user(yss_true, yss_pred, k)
    if name == 'map' and int((truth.row_lengths() == 0).sum().item()):
        raise ZeroDivisionError('float division by zero')     # ranking.py:53 divides by len(yss_true[ind])
    return float(vals[:, _CV_COL[name]].sum().item()) / len(yss_true)


def precision_k_score(yss_true, yss_pred, k=5):
    return _cv(yss_true, yss_pred, k, 'pre')


def recall_k_score(yss_true, yss_pred, k=5):
    return _cv(yss_true, yss_pred, k, 'recall')


def ndcg_k_score(yss_true, yss_pred, k=5):
    return _cv(yss_true, yss_pred, k, 'ndcg')


def map_k_score(yss_true, yss_pred, k=5):
    return _cv(yss_true, yss_pred, k, 'map')


def mrr_k_score(yss_true, yss_pred, k=5):
    return _cv(yss_true, yss_pred, k, 'mrr')


def hr_k_score(ys_true, yss_pred, k=5):
    vals, _ = per_user(ys_true, yss_pred, k, loov=True)
    return float(vals[:, 5].sum().item())


def arhr_k_score(ys_true, yss_pred, k=5):
    vals, _ = per_user(ys_true, yss_pred, k, loov=True)
    return float(vals[:, 6].sum().item())


def evaluateCV(yss_true, yss_pred, eval_metrics, k=5):
    """ranking.py:94-109.  One kernel launch serves every requested metric."""
    known = [m for m in eval_metrics if m in _CV_COL]
    sums = None
    if known:
        vals, truth = per_user(yss_true, yss_pred, k)
        if 'map' in known and int((truth.row_lengths() == 0).sum().item()):
            raise ZeroDivisionError('float division by zero')
        sums = vals.sum(0).cpu().numpy() / len(yss_true)
    return [(float(sums[_CV_COL[m]]) if m in _CV_COL else None) for m in eval_metrics]


def evaluateLOOV(ys_true, yss_pred, eval_metrics, k=5):
    """ranking.py:111-120."""
    known = [m for m in eval_metrics if m in _LOOV_COL]
    sums = None
    if known:
        vals, _ = per_user(ys_true, yss_pred, k, loov=True)
        sums = vals.sum(0).cpu().numpy()
    return [(float(sums[_LOOV_COL[m]]) if m in _LOOV_COL else None) for m in eval_metrics]
