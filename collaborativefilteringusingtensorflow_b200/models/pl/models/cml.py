"""CML with the reference's constructor and train/close entry points (reference src/models/pl/models/cml.py:12-214):
hinge on squared distances against the closest of W negatives, WARP-style rank weight, the L2 "covariance" term and
the unit-norm clip, all inside the fused step kernel; scoring is -||u - v||^2."""
from ..._base import RankingModelBase


class CML(RankingModelBase):
    _kind = 'cml'

    def __init__(self, n_users, n_items, topN=5,
                 split_method='cv', eval_metrics=['pre', 'recall', 'mrr', 'ndcg'],
                 reg_cov=1., margin=1.5, use_rank_weight=True, clip_norm=1.0,
                 n_factors=20, batch_size=100,
                 max_iter=50, lr=0.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', update='sync', seed=None, verbose=True):
        self.reg_cov, self.margin, self.use_rank_weight, self.clip_norm = reg_cov, margin, use_rank_weight, clip_norm
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, update, seed, verbose,
                    reg=float(reg_cov), margin=float(margin), use_rank_weight=bool(use_rank_weight),
                    clip_norm=float(clip_norm))

    def _train_arrays(self, batch, rows_per_batch):
        return self.engine.train_batches(batch[0], batch[1], batch_size=rows_per_batch)

    def _after_training(self, fold, tra, test_users, truth, scores):
        # cml.py:203-211: re-recommend once at topN = 1000 and score the prefixes
        topNs = [5, 10, 20, 50, 100, 200, 500, 1000]
        pred = self.recommend_device(test_users, min(topNs[-1], self.n_items), tra)
        for topN in topNs:
            self.topN = topN
            scores = self._eval(truth, pred, topN)
            if self.verbose:
                print("%s_fold=%d: " % (self.split_method, fold) + ' \tTst@' + str(self.topN) + ':' + ' '.join(
                    [m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]))
        return scores
