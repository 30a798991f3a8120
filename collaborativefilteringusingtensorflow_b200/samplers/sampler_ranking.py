"""Drop-in for the reference's src/samplers/sampler_ranking.py: ``Sampler(trasR, n_neg=5, batch_size=100, n_workers=1)``,
``next_batch() -> (pairs[B,2] int32, negs[B,W] int64)``; sampling itself runs on the GPU (cf_sample_ranking)."""
import numpy as np

from .. import _lib
from ._base import DeviceSamplerBase


class Sampler(DeviceSamplerBase):
    def __init__(self, trasR, n_neg=5, batch_size=100, n_workers=1, seed=0, device='GPU', shuffle=True):
        super(Sampler, self).__init__(trasR, batch_size, seed, device)
        self.n_neg = int(n_neg)
        self.gsize = 0
        self.shuffle = bool(shuffle)
        self.n_workers = n_workers          # accepted for signature compatibility; there are no worker threads

    def next_chunk(self, n):
        """n minibatches as CUDA int32 tensors: (pairs[n*B,2], negs[n*B,W][, group[n*B,G]])."""
        torch, B, W, G = self.torch, self.batch_size, self.n_neg, self.gsize
        pairs = self._empty('pairs', (n * B, 2), torch.int32)
        negs = self._empty('negs', (n * B, max(W, 1)), torch.int32) if W else None
        group = self._empty('group', (n * B, G), torch.int32) if G else None
        off = 0
        for epoch, batch0, count in self._segments(n):
            a = self._args(epoch, batch0, count)
            a.W, a.G, a.n_neg_rows, a.shuffle = W, G, 0, int(self.shuffle)
            if G:
                a.train_t = self.train.transpose().as_c(False)
            a.out_pairs = pairs.data_ptr() + off * 2 * 4
            a.out_negs = (negs.data_ptr() + off * W * 4) if W else None
            a.out_group = (group.data_ptr() + off * G * 4) if G else None
            _lib.check(self.lib.cf_sample_ranking(a, self._stream()), 'cf_sample_ranking')
            self.launches += 1
            off += count * B
        return (pairs, negs, group) if G else (pairs, negs)

    def _to_host_batches(self, chunk, n):
        B = self.batch_size
        pairs, negs = chunk[0].cpu().numpy(), chunk[1].cpu().numpy().astype(np.int64)
        return [(pairs[k * B:(k + 1) * B], negs[k * B:(k + 1) * B]) for k in range(n)]
