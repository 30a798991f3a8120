"""Host side of the tuple samplers (csrc/cf_tuple.cu: cf_sample_tuples) shared by sampler_prigp and sampler_uitj_ranking."""
import numpy as np

from .. import _lib
from ..sparse import DeviceCSR, null_csr
from ._base import DeviceSamplerBase


def coef_csr(coefMat, device):
    """The coefficient matrix (prigp.py:83-90 / cplr_u.py:89-97) as a device CSR with float32 values."""
    if isinstance(coefMat, DeviceCSR):
        if coefMat.values is None:
            raise ValueError('the coefficient CSR needs values')
        return coefMat
    m = coefMat.tocsr()
    m.eliminate_zeros()
    return DeviceCSR.from_scipy(m, device, with_values=True)


class TupleSamplerBase(DeviceSamplerBase):
    _model = None
    _width = None

    def _launch(self, n, epoch, batch0, tuples, coefs, off):
        a = _lib.TupleSampleArgs()
        a.train, a.coef = self.train.as_c(False), self.coef.as_c(True)
        a.collab = self.collab.as_c(False) if self.collab is not None else null_csr()
        a.eligible, a.n_eligible = _lib.ptr(self.eligible), 0 if self.eligible is None else int(self.eligible.numel())
        a.seed, a.epoch, a.batch0, a.n_batches, a.B, a.model = self.seed, epoch, batch0, n, self.batch_size, self._model
        a.out_tuples = tuples.data_ptr() + off * self._width * 4
        a.out_coefs = (coefs.data_ptr() + off * 2 * 4) if coefs is not None else None
        a.flags = _lib.ptr(self.flags)
        _lib.check(self.lib.cf_sample_tuples(a, self._stream()), 'cf_sample_tuples')
        self.launches += 1


def collaborative_rows(train, coef):
    """``ut[u] = set(coefMat[u].nonzero()) - ui[u]`` (sampler_uitj_ranking.py:13) as a device CSR, plus the users that can
    be drawn (:27: positives, collaborative items and room for a negative)."""
    torch = _lib.require_cuda()
    n_users, n_items = train.shape
    ckey = coef.rows.to(torch.int64) * n_items + coef.indices.to(torch.int64)
    tkey = train.rows.to(torch.int64) * n_items + train.indices.to(torch.int64)
    pos = torch.searchsorted(tkey, ckey).clamp_(max=max(int(tkey.numel()) - 1, 0))
    keep = (tkey[pos] != ckey) if tkey.numel() else torch.ones_like(ckey, dtype=torch.bool)
    rows, cols = coef.rows[keep], coef.indices[keep]
    counts = torch.bincount(rows.to(torch.int64), minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=rows.device)
    indptr[1:] = torch.cumsum(counts, 0)
    collab = DeviceCSR(indptr, cols.contiguous(), rows.contiguous(), None, (n_users, n_items))
    npos = train.row_lengths()
    ok = (npos > 0) & (counts > 0) & (npos + counts < n_items)
    return collab, torch.nonzero(ok).reshape(-1).to(torch.int32)


def to_host(t):
    return np.ascontiguousarray(t.cpu().numpy())
