"""Drop-in for the reference's src/samplers/sampler_rating.py: ``Sampler(trasR, negRatio=.0, batch_size=500,
n_workers=1)``, ``next_batch() -> float64 [B + int(B*negRatio), 3]`` rows (user, item, rating): B positives in file
order (never shuffled globally, :24-26) plus random (user, negative item, 0) rows, shuffled inside the batch (:38)."""
import numpy as np

from .. import _lib
from ._base import DeviceSamplerBase


class Sampler(DeviceSamplerBase):
    def __init__(self, trasR, negRatio=.0, batch_size=500, n_workers=1, seed=0, device='GPU'):
        super(Sampler, self).__init__(trasR, batch_size, seed, device, with_values=True)
        self.negRatio = negRatio
        self.num_neg = int(self.batch_size * negRatio)
        self.rows_per_batch = self.batch_size + self.num_neg
        self.n_workers = n_workers

    def next_chunk(self, n):
        """n minibatches as CUDA tensors: (ids[n*Bt,2] int32, ratings[n*Bt] float32)."""
        torch, Bt = self.torch, self.rows_per_batch
        ids = self._empty('ids', (n * Bt, 2), torch.int32)
        ratings = self._empty('ratings', (n * Bt,), torch.float32)
        off = 0
        for epoch, batch0, count in self._segments(n):
            a = self._args(epoch, batch0, count)
            a.W, a.G, a.n_neg_rows, a.shuffle = 0, 0, self.num_neg, 0
            a.out_pairs = ids.data_ptr() + off * 2 * 4
            a.out_ratings = ratings.data_ptr() + off * 4
            _lib.check(self.lib.cf_sample_rating(a, self._stream()), 'cf_sample_rating')
            self.launches += 1
            off += count * Bt
        return ids, ratings

    def _to_host_batches(self, chunk, n):
        Bt = self.rows_per_batch
        out = np.concatenate([chunk[0].cpu().numpy().astype(np.float64),
                              chunk[1].cpu().numpy().astype(np.float64)[:, None]], axis=1)
        return [out[k * Bt:(k + 1) * Bt] for k in range(n)]
