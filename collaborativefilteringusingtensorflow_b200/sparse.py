"""Device-resident CSR of the interaction matrix (what the reference keeps as ``lil_matrix.rows`` /
``dict user -> set(items)``, sampler_ranking.py:13-14, bprmf.py:117-128)."""
import numpy as np

from . import _lib


class DeviceCSR(object):
    """indptr int64 [n_rows+1], indices int32 [nnz] (sorted per row), rows int32 [nnz] (COO expansion),
    optional values float32 [nnz]; all torch CUDA tensors."""

    def __init__(self, indptr, indices, rows, values, shape):
        self.indptr, self.indices, self.rows, self.values = indptr, indices, rows, values
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(indices.numel())
        self._t = None
        self._host = None

    def __len__(self):
        return self.shape[0]

    @classmethod
    def from_scipy(cls, m, device, with_values=False):
        torch = _lib.require_cuda()
        csr = m.tocsr().astype(np.float32)
        csr.sum_duplicates()
        csr.sort_indices()
        if csr.shape[0] >= 2 ** 31 or csr.shape[1] >= 2 ** 31:
            raise ValueError('ids must fit int32')
        indptr = torch.from_numpy(csr.indptr.astype(np.int64)).to(device)
        indices = torch.from_numpy(csr.indices.astype(np.int32)).to(device)
        rows = torch.from_numpy(np.repeat(np.arange(csr.shape[0], dtype=np.int32), np.diff(csr.indptr))).to(device)
        values = torch.from_numpy(csr.data.astype(np.float32)).to(device) if with_values else None
        out = cls(indptr, indices, rows, values, csr.shape)
        out._host = csr
        return out

    @classmethod
    def from_device_coo(cls, rows, cols, shape, values=None):
        """Build from (already deduplicated) device COO arrays; sorts by (row, col) on the device."""
        torch = _lib.require_cuda()
        key = rows.to(torch.int64) * int(shape[1]) + cols.to(torch.int64)
        key, order = torch.sort(key)
        rows_s = (key // int(shape[1])).to(torch.int32)
        cols_s = (key % int(shape[1])).to(torch.int32)
        counts = torch.bincount(rows_s.to(torch.int64), minlength=int(shape[0]))
        indptr = torch.zeros(int(shape[0]) + 1, dtype=torch.int64, device=rows.device)
        indptr[1:] = torch.cumsum(counts, 0)
        vals = values[order].to(torch.float32) if values is not None else None
        return cls(indptr, cols_s, rows_s, vals, shape)

    def transpose(self):
        """item -> users CSR (the reference's ``item_posUserList``, sampler_gbpr.py:15)."""
        if self._t is None:
            self._t = DeviceCSR.from_device_coo(self.indices, self.rows, (self.shape[1], self.shape[0]), self.values)
        return self._t

    def select_rows(self, row_ids):
        """CSR holding only the listed rows (row t of the result = row row_ids[t]); host-side prep."""
        torch = _lib.require_cuda()
        ids = torch.as_tensor(row_ids, dtype=torch.int64, device=self.indptr.device)
        lo, hi = self.indptr[ids], self.indptr[ids + 1]
        lens = hi - lo
        indptr = torch.zeros(len(ids) + 1, dtype=torch.int64, device=ids.device)
        indptr[1:] = torch.cumsum(lens, 0)
        total = int(indptr[-1].item())
        seg = torch.repeat_interleave(torch.arange(len(ids), device=ids.device), lens)
        pos = torch.arange(total, device=ids.device) - indptr[seg] + lo[seg]
        return DeviceCSR(indptr, self.indices[pos].contiguous(), seg.to(torch.int32),
                         None if self.values is None else self.values[pos].contiguous(), (len(ids), self.shape[1]))

    def as_c(self, with_values=True):
        return _lib.Csr(_lib.ptr(self.indptr), _lib.ptr(self.indices), _lib.ptr(self.rows),
                        _lib.ptr(self.values) if with_values else None, self.shape[0], self.shape[1], self.nnz)

    def row_lengths(self):
        return self.indptr[1:] - self.indptr[:-1]


def null_csr():
    return _lib.Csr(None, None, None, None, 0, 0, 0)
