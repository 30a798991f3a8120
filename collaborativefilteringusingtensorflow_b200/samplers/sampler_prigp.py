"""Drop-in for the reference's src/samplers/sampler_prigp.py: ``Sampler(trasR, coefMat, batch_size=100, n_workers=1)``,
``next_batch() -> int64 [B, 5]`` rows (u, i, j, t, k); sampling runs on the GPU (cf_sample_tuples, sampler_prigp.py:22-52
restated: the epoch's shuffled positives, a uniform non-positive j, t from the user's coefficient row, k outside it or --
with probability Phi(nnz / n_items), the reference draws a standard normal there -- inside it with another coefficient
than t's, the pair ordered by coefficient)."""
import numpy as np

from .. import _lib
from ._tuple import TupleSamplerBase, coef_csr, to_host


class Sampler(TupleSamplerBase):
    _model, _width = _lib.TUPLE_PRIGP, 5

    def __init__(self, trasR, coefMat, batch_size=100, n_workers=1, seed=0, device='GPU'):
        super(Sampler, self).__init__(trasR, batch_size, seed, device)
        self.coef = coef_csr(coefMat, self.device)
        self.collab, self.eligible = None, None
        self.n_workers = n_workers          # accepted for signature compatibility; there are no worker threads

    def next_chunk(self, n):
        """n minibatches as ONE CUDA int32 tensor [n * B, 5]."""
        out = self._empty('tuples', (n * self.batch_size, 5), self.torch.int32)
        off = 0
        for epoch, batch0, count in self._segments(n):
            self._launch(count, epoch, batch0, out, None, off)
            off += count * self.batch_size
        return (out,)

    def _to_host_batches(self, chunk, n):
        t = to_host(chunk[0]).astype(np.int64)
        return [t[k * self.batch_size:(k + 1) * self.batch_size] for k in range(n)]
