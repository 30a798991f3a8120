"""Drop-in for the reference's src/samplers/sampler_uij_ranking.py: ``Sampler(trasR, batch_size=100, n_workers=1)``,
``next_batch() -> uij[B,3] int64`` (user, positive item, negative item)."""
import numpy as np

from .sampler_ranking import Sampler as _Ranking


class Sampler(_Ranking):
    def __init__(self, trasR, batch_size=100, n_workers=1, seed=0, device='GPU'):
        super(Sampler, self).__init__(trasR, 1, batch_size, n_workers, seed, device)

    def next_chunk_uij(self, n):
        pairs, negs = self.next_chunk(n)
        return self.torch.cat([pairs, negs], dim=1)

    def _to_host_batches(self, chunk, n):
        B = self.batch_size
        uij = np.concatenate([chunk[0].cpu().numpy(), chunk[1].cpu().numpy()], axis=1).astype(np.int64)
        return [uij[k * B:(k + 1) * B] for k in range(n)]
