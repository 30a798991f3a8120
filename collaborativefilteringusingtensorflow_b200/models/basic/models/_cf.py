"""Shared driver of the two neighbourhood models (reference src/models/basic/models/itemcf.py:72-94, usercf.py:69-90):
similarity on the training matrix, scores of the test users, masked top-N, evaluation."""
from .... import neighbors
from ....engine import resolve_device
from ....metrics import ranking
from ....sparse import DeviceCSR


class NeighborhoodModel(object):
    _mode = None     # 'item' | 'user'

    def __init__(self, n_users, n_items, topK=50,
                 topN=5, split_method='cv', eval_metrics=['rmse', 'mae'], *, device='GPU'):
        self.n_users, self.n_items, self.topK = int(n_users), int(n_items), int(topK)
        self.topN, self.split_method, self.eval_metrics = int(topN), split_method, list(eval_metrics)
        self.device = resolve_device(device)
        self.nbr_idx = self.nbr_sim = None
        self._train_csr = None

    def _as_csr(self, m):
        return m if isinstance(m, DeviceCSR) else DeviceCSR.from_scipy(m, self.device, with_values=True)

    def fit(self, trasR):
        """``__calsim__`` + ``__topk__`` (itemcf.py:19-40) / the neighbour choice of usercf.py:36-37: the topK most similar
        items (users) of every item (user)."""
        tra = self._as_csr(trasR)
        self._train_csr = tra
        ent = tra.transpose() if self._mode == 'item' else tra
        K = min(self.topK, ent.shape[0])
        self.nbr_idx, self.nbr_sim = neighbors.cosine_topk(ent, K)
        return self

    def predict(self, users):
        """``__predict__`` (itemcf.py:42-50, usercf.py:31-44): dense scores [len(users), n_items] (CUDA float64)."""
        import torch
        users = torch.as_tensor(users, dtype=torch.int32, device=self.device)
        return neighbors.neighbor_scores(self._train_csr, users, self.nbr_idx, self.nbr_sim, self._mode)

    def recommend_device(self, users, topN=None):
        import torch
        users = torch.as_tensor(users, dtype=torch.int32, device=self.device)
        return neighbors.topk_dense(self.predict(users), topN or self.topN, users, self._train_csr)

    def recommend(self, users, topN=None):
        """Public form of the private ``__recommend`` (itemcf.py:52-66): list of lists of item ids."""
        return [[int(x) for x in row if x >= 0] for row in self.recommend_device(users, topN).cpu().numpy()]

    def train(self, fold, trasR, tstsR):
        """Reference entry point (itemcf.py:72-94): returns the evaluation scores."""
        import torch
        self.fit(trasR)
        tst = self._as_csr(tstsR)
        test_users = torch.nonzero(tst.row_lengths() > 0).reshape(-1)               # itemcf.py:80 (ascending)
        truth = tst.select_rows(test_users)
        if self.split_method == 'loov':                                               # itemcf.py:84-85: the first test item
            truth = DeviceCSR(torch.arange(len(test_users) + 1, device=self.device, dtype=torch.int64),
                              truth.indices[truth.indptr[:-1]].contiguous(), None, None, truth.shape)
        test_users = test_users.to(torch.int32)
        pred = self.recommend_device(test_users, min(self.topN, self.n_items))
        if self.split_method == 'cv':
            return ranking.evaluateCV(truth, pred, self.eval_metrics, self.topN)
        elif self.split_method == 'loov':
            return ranking.evaluateLOOV(truth, pred, self.eval_metrics, self.topN)
        return None
