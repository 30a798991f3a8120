"""WRMF with the reference's constructor and train/close entry points (reference src/models/basic/models/wrmf.py:11-163).

``solver='sgd'`` (default) is what the reference implements: minibatch Adagrad on sampled (user, item, rating) rows with a
uniform weight (wrmf.py:58-62), fed by sampler_rating.  ``solver='als'`` is the weighted-ALS solver of the
model the reference's README cites (README.md:29): confidence ``weight`` on observed pairs, 1 elsewhere, L2 ``reg``; one
epoch = a user half-sweep + an item half-sweep (tensor-core Gram + per-row Cholesky, csrc/cf_als.cu); the sampler is not
used.  It cannot be step-compared with the reference (different algorithm); tests compare it with a dense fp64 solve."""
import numpy as np

from ..._base import RankingModelBase


class WRMF(RankingModelBase):
    _kind = 'wrmf'

    def __init__(self, n_users, n_items, topN=10,
                 split_method='cv', eval_metrics=['pre', 'recall', 'map', 'mrr', 'ndcg'],
                 weight=1, reg=0.02, n_factors=10, batch_size=500,
                 max_iter=50, lr=.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', update='sync', seed=None, verbose=True, solver='sgd'):
        if solver not in ('sgd', 'als'):
            raise ValueError("solver must be 'sgd' or 'als'")
        self.weight, self.reg, self.solver = weight, reg, solver
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, update, seed, verbose, reg=float(reg), weight=float(weight))

    def _train_arrays(self, batch, rows_per_batch):
        torch = self.engine.torch
        if len(batch) == 1:        # the reference's float64 [rows, 3] array (wrmf.py:145-147)
            uir = batch[0]
            if torch.is_tensor(uir):
                ids, ratings = uir[:, :2].to(torch.int32).contiguous(), uir[:, 2].to(torch.float32).contiguous()
            else:
                uir = np.asarray(uir)
                ids, ratings = uir[:, :2].astype(np.int32), uir[:, 2].astype(np.float32)
        else:                      # device sampler chunk: (ids int32 [rows, 2], ratings float32 [rows])
            ids, ratings = batch
        return self.engine.train_batches(ids, ratings=ratings, batch_size=rows_per_batch)

    def _epoch(self, sampler, n_batches):
        if self.solver == 'sgd':
            return super(WRMF, self)._epoch(sampler, n_batches)
        tra = self._train_csr
        self.engine.als_half_sweep('users', tra)
        self.engine.als_half_sweep('items', tra.transpose())
        return self.engine.torch.full((1,), float('nan'), dtype=self.engine.torch.float64, device=self.device)

    def _log_line(self, fold, it, aveloss, scores):
        # wrmf.py:153-156
        return ("fold=%d iter=%2d: " % (fold, it + 1) + " TraLoss=%.4f lr=%.4f" % (aveloss, self._printed_lr) +
                ' \tTst:' + ' '.join([m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]))
