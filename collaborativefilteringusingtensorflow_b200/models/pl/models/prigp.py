"""PRIGP with the reference's constructor and train/close entry points (reference src/models/pl/models/prigp.py:17-231;
driver pl/testprigp.py): BPR on (u, i, j) plus an alpha-weighted BPR term on a collaborative pair (t, k) drawn from the
user's neighbour-count coefficients; item bias in the score, Adagrad on the two embedding tables only (:134)."""
from .... import _lib
from ._tuple import TupleModelBase


class PRIGP(TupleModelBase):
    _tuple_model = _lib.TUPLE_PRIGP
    _weighted_coef = False

    def __init__(self, n_users, n_items,
                 topK=50, topN=5,
                 split_method='cv', eval_metrics=['pre', 'recall', 'map', 'mrr', 'ndcg'],
                 alpha=1.,
                 reg=0.01, n_factors=20, batch_size=1000,
                 max_iter=50, lr=0.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', seed=None, verbose=True):
        self.alpha, self.reg = alpha, reg
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, 'sync', seed, verbose, reg=float(reg))
        self._init_tuple(topK)
        self._seed = seed

    def _make_sampler(self, tra, coef_csr_):
        from ....samplers.sampler_prigp import Sampler
        return Sampler(tra, coef_csr_, self.batch_size, seed=self._seed or 0, device=self.device)
