"""BPRMF with the reference's constructor and train/close entry points (reference src/models/pl/models/bprmf.py:12-173).
The TF1 graph is replaced by the fused sm_100a step kernel (csrc/cf_step_impl.cuh) and the masked top-K kernel."""
from ..._base import RankingModelBase


class BPRMF(RankingModelBase):
    _kind = 'bpr'

    def __init__(self, n_users, n_items, topN=5,
                 split_method='cv', eval_metrics=['pre', 'recall', 'mrr', 'ndcg'],
                 reg=0.02, n_factors=20, batch_size=100,
                 max_iter=50, lr=0.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', update='sync', seed=None, verbose=True):
        self.reg = reg
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, update, seed, verbose, reg=float(reg))

    def _train_arrays(self, batch, rows_per_batch):
        pairs, negs = batch[0], batch[1]
        return self.engine.train_batches(pairs, negs, batch_size=rows_per_batch)
