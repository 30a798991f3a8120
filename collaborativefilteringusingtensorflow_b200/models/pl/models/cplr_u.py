"""CPLR with the reference's constructor and train/close entry points (reference src/models/pl/models/cplr_u.py:17-294;
driver pl/testcplr_u.py): three coefficient-weighted BPR terms over (positive i, collaborative t, negative j) with
alpha / beta / gamma mixing; item bias in the score; Adagrad on both embedding tables and the bias (:141).  The
coefficient matrix is the similarity-weighted neighbour sum of cplr_u.py:89-97, divided per user by the mean of its
non-zero entries (:196-199)."""
from .... import _lib
from ._tuple import TupleModelBase


class CPLR(TupleModelBase):
    _tuple_model = _lib.TUPLE_CPLR
    _weighted_coef = True

    def __init__(self, n_users, n_items,
                 topK=50, topN=5,
                 split_method='cv', eval_metrics=['pre', 'recall', 'map', 'mrr', 'ndcg'],
                 alpha=1., beta=1., gamma=1.,
                 reg=0.01, n_factors=20, batch_size=1000,
                 max_iter=50, lr=0.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', seed=None, verbose=True):
        self.alpha, self.beta, self.gamma, self.reg = alpha, beta, gamma, reg
        self._setup(n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
                    init_mean, init_stddev, device, optimizer, 'sync', seed, verbose, reg=float(reg))
        self._init_tuple(topK)
        self._seed = seed

    def _normalise(self, coef):
        """cplr_u.py:196-199: every user's row divided by the mean of its non-zero entries."""
        nnz = (coef != 0).sum(1)
        ave = coef.sum(1) / nnz.clamp(min=1)
        scale = self.engine.torch.where(ave > 0, 1.0 / ave.clamp(min=1e-300), self.engine.torch.ones_like(ave))
        return coef * scale[:, None]

    def _make_sampler(self, tra, coef_csr_):
        from ....samplers.sampler_uitj_ranking import Sampler
        return Sampler(tra, coef_csr_, self.batch_size, seed=self._seed or 0, device=self.device)
