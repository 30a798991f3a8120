"""Rating metrics with the reference's signatures and outputs (reference src/metrics/rating.py:4-31), computed by the
``cf_rating_metrics`` CUDA kernel (fp64 sums of |t - p| and (t - p)^2); the division by n and the square root are done
on the host exactly where rating.py does them.  Inputs may be numpy arrays / sequences (uploaded) or CUDA tensors.
``evaluate`` returns ``None`` for unknown metric names, like the reference; an empty input raises ZeroDivisionError
(rating.py evaluates ``1 / ys_true.shape[0]``)."""
import math

import numpy as np

from .. import _lib


def _device_arrays(ys_true, ys_pred, device=None):
    torch = _lib.require_cuda()
    if device is None:
        device = next((t.device for t in (ys_pred, ys_true) if torch.is_tensor(t) and t.is_cuda),
                      torch.device('cuda', torch.cuda.current_device()))
    if torch.is_tensor(ys_true):
        t = ys_true.to(device=device, dtype=torch.float64).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(ys_true, dtype=np.float64))).to(device)
    if torch.is_tensor(ys_pred):
        p = ys_pred.to(device=device)
        if p.dtype not in (torch.float32, torch.float64):
            p = p.to(torch.float64)
        p = p.contiguous()
    else:
        p = torch.from_numpy(np.ascontiguousarray(np.asarray(ys_pred, dtype=np.float64))).to(device)
    return t.reshape(-1), p.reshape(-1)


def error_sums(ys_true, ys_pred, clip=None, device=None):
    """(sum |t - p|, sum (t - p)^2, n) with p clipped to ``clip = (lo, hi)`` first (mf.py:81) when given."""
    torch = _lib.require_cuda()
    t, p = _device_arrays(ys_true, ys_pred, device)
    n = int(t.numel())
    if n == 0:
        raise ZeroDivisionError('division by zero')
    if int(p.numel()) != n:
        raise ValueError('operands could not be broadcast together with shapes (%d,) (%d,)' % (n, int(p.numel())))
    lo, hi = (-math.inf, math.inf) if clip is None else (float(clip[0]), float(clip[1]))
    sums = torch.zeros(2, dtype=torch.float64, device=t.device)
    _lib.check(_lib.lib().cf_rating_metrics(_lib.ptr(p), int(p.dtype == torch.float64), _lib.ptr(t), n, lo, hi, _lib.ptr(sums),
                                            torch.cuda.current_stream(t.device).cuda_stream), 'cf_rating_metrics')
    a, q = sums.tolist()
    return a, q, n


def mean_absolute_error(ys_true, ys_pred):
    """MAE: mean of |t - p| over the rating pairs (same name and value as rating.py:4-6)."""
    a, _, n = error_sums(ys_true, ys_pred)
    return 1 / n * a


def mean_squared_error(ys_true, ys_pred):
    """MSE: mean of (t - p)^2 (same name and value as rating.py:9-11)."""
    _, q, n = error_sums(ys_true, ys_pred)
    return 1 / n * q


def root_mean_squared_error(ys_true, ys_pred):
    """RMSE: square root of the MSE (same name and value as rating.py:14-16)."""
    _, q, n = error_sums(ys_true, ys_pred)
    return float(np.sqrt(1 / n * q))


def evaluate(ys_true, ys_pred, eval_metrics, clip=None):
    """rating.py:18-29; one kernel launch serves every requested metric."""
    known = [m for m in eval_metrics if m in ('mae', 'mse', 'rmse')]
    a = q = n = None
    if known:
        a, q, n = error_sums(ys_true, ys_pred, clip)
    scores = []
    for eval_metric in eval_metrics:
        if eval_metric == 'mae':
            scores.append(1 / n * a)
        elif eval_metric == 'mse':
            scores.append(1 / n * q)
        elif eval_metric == 'rmse':
            scores.append(float(np.sqrt(1 / n * q)))
        else:
            scores.append(None)
    return scores
