"""Drop-in for the reference's src/samplers/sampler_gbpr.py: ``Sampler(trasR, gsize=2, n_neg=5, batch_size=100,
n_workers=1)``, ``next_batch() -> (pairs[B,2] int32, negs[B,W] int64, group[B,G] int64)``; group members are drawn with
replacement from the users that have the pair's item (sampler_gbpr.py:15,41) through an item->user CSR on the device."""
import numpy as np

from .sampler_ranking import Sampler as _Ranking


class Sampler(_Ranking):
    def __init__(self, trasR, gsize=2, n_neg=5, batch_size=100, n_workers=1, seed=0, device='GPU'):
        super(Sampler, self).__init__(trasR, n_neg, batch_size, n_workers, seed, device)
        self.gsize = int(gsize)
        if self.gsize < 1:
            raise ValueError('gsize must be >= 1')
        self.train.transpose()

    def _to_host_batches(self, chunk, n):
        B = self.batch_size
        pairs = chunk[0].cpu().numpy()
        negs, group = chunk[1].cpu().numpy().astype(np.int64), chunk[2].cpu().numpy().astype(np.int64)
        return [(pairs[k * B:(k + 1) * B], negs[k * B:(k + 1) * B], group[k * B:(k + 1) * B]) for k in range(n)]
