"""GBPRMF with the reference's constructor and train/close entry points (reference src/models/pl/models/gbprmf.py:13-186):
BPR with a group-averaged positive score (rho mix), item bias and Adagrad on U, V and b."""
import datetime as dt

from ..._base import RankingModelBase


class GBPRMF(RankingModelBase):
    _kind = 'gbpr'

    def __init__(self, n_users, n_items, topN=10, rho=.5, gsize=2,
                 split_method='cv', eval_metrics=['pre', 'recall', 'mrr', 'ndcg'],
                 reg=0.02, n_factors=20, batch_size=100,
                 max_iter=30, lr=0.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', update='sync', seed=None, verbose=True):
        self.rho, self.gsize, self.reg = rho, gsize, reg
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, update, seed, verbose, reg=float(reg), rho=float(rho))

    def _train_arrays(self, batch, rows_per_batch):
        pairs, negs, group = batch[0], batch[1], batch[2]
        if int(group.shape[1]) != int(self.gsize):
            raise ValueError('group has %d columns, model was built with gsize=%d' % (group.shape[1], self.gsize))
        return self.engine.train_batches(pairs, negs, group, batch_size=rows_per_batch)

    def _format_epoch(self, fold, it, aveloss, scores, t0, t1):
        # gbprmf.py:171-176 prints a timestamp and the epoch's wall time as well
        return (dt.datetime.now().strftime('%m-%d %H:%M:%S') + ' ' +
                "%s_fold=%d iter=%2d:" % (self.split_method, fold, it + 1) +
                " TraLoss=%.4f lr=%.4f" % (aveloss, self._printed_lr) +
                ' \tTst@' + str(self.topN) + ':' + ' '.join([m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]) +
                " \t\ttimecost=%d(s)" % (t1 - t0).seconds)
