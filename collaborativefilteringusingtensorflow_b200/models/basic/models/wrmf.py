"""WRMF with the reference's constructor and train/close entry points (reference src/models/basic/models/wrmf.py:11-163).

``solver='sgd'`` (default) is what the reference implements: minibatch Adagrad on sampled (user, item, rating) rows with a
uniform weight (wrmf.py:58-62), fed by sampler_rating."""
import numpy as np

from ..._base import RankingModelBase


class WRMF(RankingModelBase):
    _kind = 'wrmf'

    def __init__(self, n_users, n_items, topN=10,
                 split_method='cv', eval_metrics=['pre', 'recall', 'map', 'mrr', 'ndcg'],
                 weight=1, reg=0.02, n_factors=10, batch_size=500,
                 max_iter=50, lr=.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', update='sync', seed=None, verbose=True):
        self.weight, self.reg = weight, reg
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, update, seed, verbose, reg=float(reg), weight=float(weight))

    def _train_arrays(self, batch, rows_per_batch):
        torch = self.engine.torch
        if len(batch) == 1:        # the reference's float64 [rows, 3] array (wrmf.py:145-147)
            uir = batch[0]
            if torch.is_tensor(uir):
                ids, ratings = uir[:, :2].to(torch.int32), uir[:, 2].to(torch.float32)
            else:
                uir = np.asarray(uir)
                ids, ratings = uir[:, :2].astype(np.int32), uir[:, 2].astype(np.float32)
        else:                      # device sampler chunk: (ids int32 [rows, 2], ratings float32 [rows])
            ids, ratings = batch
        return self.engine.train_batches(ids, ratings=ratings, batch_size=rows_per_batch)

    def _log_line(self, fold, it, aveloss, scores):
        # wrmf.py:153-156
        return ("fold=%d iter=%2d: " % (fold, it + 1) + " TraLoss=%.4f lr=%.4f" % (aveloss, self._printed_lr) +
                ' \tTst:' + ' '.join([m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]))
