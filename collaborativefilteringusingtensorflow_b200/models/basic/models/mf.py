"""MF (rating prediction) with the reference's constructor and train/close entry points
(reference src/models/basic/models/mf.py:12-113; driver basic/testmf.py).

The minibatch objective ``l2_loss(<U_u, V_i> - r) + reg * (l2_loss(U_u) + l2_loss(V_i))`` (mf.py:54-64) is WRMF's with
``weight = 1`` (wrmf.py:52-75), so the fused step kernel is the same instantiation (CF_MODEL_WRMF); what is new is the
evaluation: predictions of the test tuples (cf_predict_pairs), clipped to ``range_of_ratings`` (mf.py:81), scored by
metrics/rating.py (cf_rating_metrics)."""
import numpy as np

from ..._base import RankingModelBase
from ....metrics import rating


class MF(RankingModelBase):
    _kind = 'wrmf'

    def __init__(self, n_users, n_items, eval_metrics=['rmse', 'mae'],
                 range_of_ratings=(.5, 5), reg=0.02, n_factors=10, batch_size=500,
                 max_iter=50, lr=.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', update='sync', seed=None, verbose=True):
        self.range_of_ratings, self.reg = range_of_ratings, reg
        self._setup(n_users, n_items, 1, 'cv', eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, update, seed, verbose, reg=float(reg), weight=1.0)

    def _train_arrays(self, batch, rows_per_batch):
        torch = self.engine.torch
        if len(batch) == 1:        # the reference's float64 [rows, 3] array (mf.py:97-98)
            uir = batch[0]
            if torch.is_tensor(uir):
                ids, ratings = uir[:, :2].to(torch.int32).contiguous(), uir[:, 2].to(torch.float32).contiguous()
            else:
                uir = np.asarray(uir)
                ids, ratings = uir[:, :2].astype(np.int32), uir[:, 2].astype(np.float32)
        else:                      # device sampler chunk: (ids int32 [rows, 2], ratings float32 [rows])
            ids, ratings = batch
        return self.engine.train_batches(ids, ratings=ratings, batch_size=rows_per_batch)

    def predict_pairs(self, useritem):
        """Public form of ``__predict`` (mf.py:66-72) for explicit (user, item) rows: numpy float32 [n], unclipped."""
        out = self.engine.predict_pairs(useritem)
        self.engine.check_flags()
        return out.cpu().numpy()

    def evaluate(self, tst_tuple):
        """``__eval`` (mf.py:80-83): clip the predictions of ``tst_tuple[:, :2]`` to range_of_ratings, score against
        ``tst_tuple[:, 2]``."""
        torch = self.engine.torch
        if torch.is_tensor(tst_tuple):
            ids, truth = tst_tuple[:, :2], tst_tuple[:, 2]
        else:
            tst_tuple = np.asarray(tst_tuple)
            ids, truth = tst_tuple[:, :2].astype(np.int32), tst_tuple[:, 2].astype(np.float64)
        pred = self.engine.predict_pairs(ids)
        scores = rating.evaluate(truth, pred, self.eval_metrics, clip=self.range_of_ratings)
        self.engine.check_flags()
        return scores

    def train(self, fold, tra_tuple, tst_tuple, sampler):
        """mf.py:86-110: max_iter epochs of int(len(tra_tuple) / batch_size) minibatches from the rating sampler,
        evaluation + one printed line per epoch, returns the last epoch's scores."""
        n_batches = int(len(tra_tuple) / self.batch_size)
        scores = None
        for it in range(self.max_iter):
            losses = self._epoch(sampler, n_batches)
            aveloss = float(losses.mean().item())
            self.engine.check_flags()
            if hasattr(sampler, 'check_flags'):
                sampler.check_flags()
            scores = self.evaluate(tst_tuple)
            if self.verbose:
                print("fold=%d iter=%2d: " % (fold, it + 1),
                      "TraLoss=%.4f lr=%.4f" % (aveloss, self._printed_lr),
                      '\tTst:' + ' '.join([m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]))
            self._printed_lr *= .98       # mf.py:106 decays a float the optimizer never sees again
        return scores
