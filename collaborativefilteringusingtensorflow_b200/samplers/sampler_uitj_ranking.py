"""Drop-in for the reference's src/samplers/sampler_uitj_ranking.py: ``Sampler(trasR, coefMat, batch_size=1000, n_workers=1)``,
``next_batch() -> (uitj int64 [B, 4], coefs float64 [B, 2])``; sampling runs on the GPU (cf_sample_tuples,
sampler_uitj_ranking.py:22-38 restated: a uniform user among those with positives, collaborative items and room for a
negative, a uniform positive i, a uniform collaborative item t, a uniform item j in neither; coefs = (coef[u, i], coef[u, t]))."""
import numpy as np

from .. import _lib
from ._tuple import TupleSamplerBase, coef_csr, collaborative_rows, to_host


class Sampler(TupleSamplerBase):
    _model, _width = _lib.TUPLE_CPLR, 4

    def __init__(self, trasR, coefMat, batch_size=1000, n_workers=1, seed=0, device='GPU'):
        super(Sampler, self).__init__(trasR, batch_size, seed, device)
        self.coef = coef_csr(coefMat, self.device)
        self.collab, self.eligible = collaborative_rows(self.train, self.coef)
        if int(self.eligible.numel()) == 0:
            raise ValueError('no user has positives, collaborative items and room for a negative '
                             '(the reference sampler would spin forever, sampler_uitj_ranking.py:27)')
        self.n_workers = n_workers
        self.batches_per_epoch = 1 << 40     # an i.i.d. stream: the cursor only counts minibatches

    def next_chunk(self, n):
        """n minibatches as CUDA tensors (uitj int32 [n * B, 4], coefs float32 [n * B, 2])."""
        torch = self.torch
        out = self._empty('tuples', (n * self.batch_size, 4), torch.int32)
        coefs = self._empty('coefs', (n * self.batch_size, 2), torch.float32)
        self._launch(n, 0, self.batch, out, coefs, 0)
        self.batch += n
        return out, coefs

    def _to_host_batches(self, chunk, n):
        t, c = to_host(chunk[0]).astype(np.int64), to_host(chunk[1]).astype(np.float64)
        B = self.batch_size
        return [(t[k * B:(k + 1) * B], c[k * B:(k + 1) * B]) for k in range(n)]
