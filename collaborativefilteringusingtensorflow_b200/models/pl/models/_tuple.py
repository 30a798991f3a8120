"""Shared driver of PRIGP and CPLR (reference src/models/pl/models/prigp.py:174-228, cplr_u.py:178-290): user-user cosine
similarity, top-K neighbours, coefficient matrix, the tuple sampler, epochs of gradient-only steps (cf_tuple_grads) +
dense applies, full-catalogue evaluation with U V^T + b."""
import numpy as np

from .... import _lib, neighbors
from ....sparse import DeviceCSR
from ..._base import RankingModelBase


class TupleModelBase(RankingModelBase):
    _kind = 'gbpr'          # tables U, V, b with Adagrad state; scoring U V^T + b (prigp.py:124-128)
    _tuple_model = None
    _weighted_coef = None   # cplr_u.py:93 sums similarities, prigp.py:87 counts neighbours

    def _init_tuple(self, topK):
        eng, torch = self.engine, self.engine.torch
        self.topK = int(topK)
        self._gU, self._gV = torch.zeros_like(eng.U), torch.zeros_like(eng.V)
        self._gb = torch.zeros_like(eng.b)
        self.coefMat = None

    # ------------------------------------------------------------------ preprocessing (prigp.py:60-90, cplr_u.py:64-97)
    def coefficient_matrix(self, trasR):
        """``__topk__(__calsim__(trasR))`` + ``__calcoef__``: float64 dense on the device, rows of users with positives."""
        torch = self.engine.torch
        tra = trasR if isinstance(trasR, DeviceCSR) else DeviceCSR.from_scipy(trasR, self.device, with_values=True)
        K = min(self.topK, tra.shape[0])
        nbr_idx, nbr_sim = neighbors.cosine_topk(tra, K)
        users = torch.arange(tra.shape[0], dtype=torch.int32, device=self.device)
        if self._weighted_coef:
            coef = neighbors.neighbor_scores(tra, users, nbr_idx, nbr_sim, 'user')
        else:       # prigp.py:87: user_predict += trasR[nn_user] > 0
            ones = (nbr_idx >= 0).to(torch.float32)
            binary = DeviceCSR(tra.indptr, tra.indices, tra.rows, None, tra.shape)
            coef = neighbors.neighbor_scores(binary, users, nbr_idx, ones, 'user')
        coef[tra.row_lengths() == 0] = 0          # `for user in set(trasR.nonzero()[0])`
        return coef

    def _coef_to_csr(self, coef):
        torch = self.engine.torch
        nz = torch.nonzero(coef)
        rows, cols = nz[:, 0].to(torch.int32), nz[:, 1].to(torch.int32)
        counts = torch.bincount(nz[:, 0], minlength=coef.shape[0])
        indptr = torch.zeros(coef.shape[0] + 1, dtype=torch.int64, device=coef.device)
        indptr[1:] = torch.cumsum(counts, 0)
        return DeviceCSR(indptr, cols.contiguous(), rows.contiguous(), coef[nz[:, 0], nz[:, 1]].to(torch.float32).contiguous(), coef.shape)

    # ------------------------------------------------------------------ one sess.run(train_op)
    def _apply(self, table, acc, grad, n_rows, d, ld):
        eng = self.engine
        ap = _lib.ApplyArgs()
        ap.table, ap.acc, ap.n_rows, ap.d, ap.ld = _lib.ptr(table), _lib.ptr(acc), n_rows, d, ld
        ap.grads, ap.ldg = _lib.ptr(grad), ld
        ap.model, ap.optimizer = _lib.MODEL_BPR, 0 if eng.optimizer == 'adagrad' else 1
        ap.lr, ap.clip_norm = eng.hyper['lr'], eng.hyper['clip_norm']
        _lib.check(eng.lib.cf_apply_dense(ap, eng.torch.cuda.current_stream(eng.device).cuda_stream), 'cf_apply_dense')

    def _train_arrays(self, batch, rows_per_batch):
        eng, torch = self.engine, self.engine.torch
        tuples = eng._as_i32(batch[0])
        coefs = None
        if self._tuple_model == _lib.TUPLE_CPLR:
            coefs = batch[1]
            if not torch.is_tensor(coefs):
                coefs = torch.from_numpy(np.ascontiguousarray(np.asarray(coefs, dtype=np.float32)))
            coefs = coefs.to(device=eng.device, dtype=torch.float32).contiguous()
        width = 5 if self._tuple_model == _lib.TUPLE_PRIGP else 4
        if tuples.dim() != 2 or tuples.shape[1] != width:
            raise ValueError('tuples must be [rows, %d]' % width)
        rows = int(tuples.shape[0])
        B = int(rows_per_batch or rows)
        if rows == 0 or rows % B:
            raise ValueError('rows (%d) must be a positive multiple of batch_size (%d)' % (rows, B))
        nb = rows // B
        loss = torch.zeros(nb, dtype=torch.float64, device=eng.device)
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        for k in range(nb):
            a = _lib.TupleArgs()
            a.U, a.V, a.b = _lib.ptr(eng.U), _lib.ptr(eng.V), _lib.ptr(eng.b)
            a.n_users, a.n_items, a.d, a.ld, a.model = eng.n_users, eng.n_items, eng.d, eng.ld, self._tuple_model
            a.tuples = tuples.data_ptr() + 4 * width * B * k
            a.coefs = (coefs.data_ptr() + 8 * B * k) if coefs is not None else None
            a.B = B
            a.alpha, a.beta, a.gamma, a.reg = float(self.alpha), float(getattr(self, 'beta', 0.0)), float(getattr(self, 'gamma', 0.0)), float(self.reg)
            a.gradU, a.gradV, a.gradb = _lib.ptr(self._gU), _lib.ptr(self._gV), _lib.ptr(self._gb)
            a.loss, a.counters = loss.data_ptr() + 8 * k, _lib.ptr(eng.counters)
            _lib.check(eng.lib.cf_tuple_grads(a, stream), 'cf_tuple_grads')
            self._apply(eng.U, eng.accU, self._gU, eng.n_users, eng.d, eng.ld)
            self._apply(eng.V, eng.accV, self._gV, eng.n_items, eng.d, eng.ld)
            if self._tuple_model == _lib.TUPLE_CPLR:      # prigp.py:134 leaves item_bias out of var_list
                self._apply(eng.b, eng.accb, self._gb, eng.n_items, 1, 1)
            eng.launches += 3 + (1 if self._tuple_model == _lib.TUPLE_CPLR else 0)
        return loss

    def step(self, *batch):
        """One minibatch on explicit arrays: PRIGP step(uijtk[B, 5]); CPLR step(uitj[B, 4], coefs[B, 2])."""
        return super(TupleModelBase, self).step(*batch)

    def _make_sampler(self, tra, coef_csr_):
        raise NotImplementedError

    def train(self, fold, trasR, tstsR, sampler=None):
        """Reference entry point (prigp.py:174-228): the reference builds its sampler itself; one may be passed in."""
        tra = self._as_csr(trasR) if not isinstance(trasR, DeviceCSR) else trasR
        if sampler is None:
            if tra.values is None:
                tra = DeviceCSR(tra.indptr, tra.indices, tra.rows, self.engine.torch.ones(tra.nnz, device=self.device), tra.shape)
            coef = self._normalise(self.coefficient_matrix(tra))
            self.coefMat = self._coef_to_csr(coef)
            sampler = self._make_sampler(tra, self.coefMat)
        return super(TupleModelBase, self).train(fold, tra, tstsR, sampler)

    def _normalise(self, coef):
        return coef

    def _format_epoch(self, fold, it, aveloss, scores, t0, t1):
        import datetime as dt
        return (dt.datetime.now().strftime('%m-%d %H:%M:%S') + ' ' +
                "%s_fold=%d iter=%2d:" % (self.split_method, fold, it + 1) +
                " TraLoss=%.4f lr=%.4f" % (aveloss, self._printed_lr) +
                ' \tTst@' + str(self.topN) + ':' + ' '.join([m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]) +
                " \t\ttimecost=%d(s)" % (t1 - t0).seconds)
