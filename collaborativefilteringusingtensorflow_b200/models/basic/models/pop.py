"""PopRank with the reference's constructor and ``train`` entry point (reference src/models/basic/models/pop.py:10-63;
driver basic/testpop.py).

The reference ranks items by the number of training interactions (``sorted(..., reverse=True)``: a stable sort, so ties
keep ascending item order, pop.py:19-21) and walks that list per test user skipping the user's training items
(:23-35).  That is the masked top-N of the hot path with the same tie rule (score desc, id asc), so it runs on the same
kernel: a 1-factor engine whose user rows are 1 and whose item rows hold the popularity count (exact in fp32) gives
``score(u, i) = pop_i``.  The reference is numpy-only, so this model is pinned end to end against the reference itself
(tests/golden/pop_golden.json: its recommended lists and metric values on ml-100k fold 1)."""
from ..._base import RankingModelBase


class PopRank(RankingModelBase):
    _kind = 'bpr'

    def __init__(self, n_users, n_items,
                 topN=5, split_method='cv', eval_metrics=['rmse', 'mae'], *, device='GPU'):
        self._setup(n_users, n_items, topN, split_method, eval_metrics, 1, 1, 0, 0.0, 0.0, 0.1, device, 'adagrad', 'sync', 0,
                    False)
        self.engine.U.zero_()
        self.engine.V.zero_()
        self.engine.U[:, 0] = 1.0

    def _fit_popularity(self, tra):
        """pop.py:19-21: interactions per item (column nnz of the training matrix)."""
        torch = self.engine.torch
        pop = torch.bincount(tra.indices.to(torch.int64), minlength=self.n_items)
        if int(pop.max().item()) >= (1 << 24):
            raise ValueError('an item with more than 2^24 interactions does not fit the exact fp32 count')
        self.engine.V[:, 0] = pop.to(torch.float32)

    def train(self, fold, trasR, tstsR, sampler=None):
        """pop.py:45-62: count, recommend topN unseen items per test user, evaluate."""
        tra, test_users, truth = self._prepare_eval(trasR, tstsR)
        self._fit_popularity(tra)
        pred = self.recommend_device(test_users, min(self.topN, self.n_items), tra)
        return self._eval(truth, pred)
