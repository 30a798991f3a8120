"""SVD (rating prediction with a d x d kernel matrix) with the reference's constructor and train/close entry points
(reference src/models/basic/models/svd.py:12-113; driver basic/testsvd.py).

``pred = sum((U_u @ K) * V_i)`` (svd.py:66-72); loss ``l2_loss(pred - r) + reg * (l2_loss(U_u) + l2_loss(V_i))``
(:52-64, no L2 on K); Adagrad on U, K and V (:74-80).  Because every minibatch updates the dense matrix K, the step is
done in the gradient-only form: ``cf_svd_grads`` sums every pair's gradients into dense tables, ``cf_apply_dense``
applies them (rows with an all-zero gradient are skipped, which equals TF's sparse apply on the gathered rows)."""
import numpy as np

from .... import _lib
from ....metrics import rating
from .mf import MF

ADAGRAD_ACC0 = 0.1


class SVD(MF):
    def __init__(self, n_users, n_items, eval_metrics=['rmse', 'mae'],
                 range_of_ratings=(.5, 5), reg=0.02, n_factors=10, batch_size=500,
                 max_iter=50, lr=.1,
                 init_mean=0.0, init_stddev=0.1,
                 device='CPU', *, optimizer='adagrad', seed=None, verbose=True):
        if n_factors > 128:
            raise ValueError('SVD supports n_factors <= 128')
        super(SVD, self).__init__(n_users, n_items, eval_metrics, range_of_ratings, reg, n_factors, batch_size, max_iter, lr,
                                  init_mean, init_stddev, device, optimizer=optimizer, update='sync', seed=seed, verbose=verbose)
        eng, torch = self.engine, self.engine.torch
        gen = torch.Generator(device=eng.device)
        gen.manual_seed(int(seed) + 0x5D if seed is not None else int(np.random.SeedSequence().generate_state(1)[0]))
        self.K = torch.empty(eng.d, eng.ld, device=eng.device)
        self.K.zero_()
        k = torch.empty(eng.d, eng.d, device=eng.device)
        torch.nn.init.trunc_normal_(k, init_mean, init_stddev, init_mean - 2 * init_stddev, init_mean + 2 * init_stddev,
                                    generator=gen)                      # tf.truncated_normal_initializer (svd.py:31-33)
        self.K[:, :eng.d].copy_(k)
        self.accK = torch.full_like(self.K, ADAGRAD_ACC0)
        self._gU = torch.zeros_like(eng.U)
        self._gV = torch.zeros_like(eng.V)
        self._gK = torch.zeros_like(self.K)

    # ------------------------------------------------------------------ parameters
    def state_dict(self):
        sd = self.engine.state_dict()
        sd['K'], sd['accK'] = self.K[:, :self.engine.d].clone(), self.accK[:, :self.engine.d].clone()
        return sd

    def load_state_dict(self, sd):
        self.engine.load_state_dict(sd)
        torch, d = self.engine.torch, self.engine.d
        for name in ('K', 'accK'):
            if name in sd:
                src = torch.as_tensor(np.asarray(sd[name]) if not torch.is_tensor(sd[name]) else sd[name],
                                      dtype=torch.float32, device=self.engine.device)
                if tuple(src.shape) != (d, d):
                    raise ValueError('%s: expected shape %s, got %s' % (name, (d, d), tuple(src.shape)))
                getattr(self, name)[:, :d].copy_(src)

    # ------------------------------------------------------------------ one sess.run(train_op)
    def _args(self, ids, ratings, n):
        eng = self.engine
        a = _lib.SvdArgs()
        a.U, a.V, a.K = _lib.ptr(eng.U), _lib.ptr(eng.V), _lib.ptr(self.K)
        a.n_users, a.n_items, a.d, a.ld, a.ldk = eng.n_users, eng.n_items, eng.d, eng.ld, int(self.K.stride(0))
        a.pairs, a.ratings, a.B, a.reg = _lib.ptr(ids), _lib.ptr(ratings), n, float(self.reg)
        a.counters = _lib.ptr(eng.counters)
        return a

    def _apply(self, table, acc, grad, n_rows, d, ld):
        eng = self.engine
        ap = _lib.ApplyArgs()
        ap.table, ap.acc, ap.n_rows, ap.d, ap.ld = _lib.ptr(table), _lib.ptr(acc), n_rows, d, ld
        ap.grads, ap.ldg = _lib.ptr(grad), ld
        ap.model, ap.optimizer = eng.model_id, 0 if eng.optimizer == 'adagrad' else 1
        ap.lr, ap.clip_norm = eng.hyper['lr'], eng.hyper['clip_norm']
        _lib.check(eng.lib.cf_apply_dense(ap, eng.torch.cuda.current_stream(eng.device).cuda_stream), 'cf_apply_dense')

    def _train_arrays(self, batch, rows_per_batch):
        eng, torch = self.engine, self.engine.torch
        if len(batch) == 1:        # the reference's float64 [rows, 3] array (svd.py:97-98)
            uir = batch[0]
            if torch.is_tensor(uir):
                ids, ratings = uir[:, :2], uir[:, 2]
            else:
                uir = np.asarray(uir)
                ids, ratings = uir[:, :2].astype(np.int32), uir[:, 2].astype(np.float32)
        else:
            ids, ratings = batch
        ids = eng._as_i32(ids)
        if not torch.is_tensor(ratings):
            ratings = torch.from_numpy(np.ascontiguousarray(np.asarray(ratings, dtype=np.float32)))
        ratings = ratings.to(device=eng.device, dtype=torch.float32).contiguous()
        rows = int(ids.shape[0])
        B = int(rows_per_batch or rows)
        if rows == 0 or rows % B:
            raise ValueError('rows (%d) must be a positive multiple of batch_size (%d)' % (rows, B))
        nb = rows // B
        loss = torch.zeros(nb, dtype=torch.float64, device=eng.device)
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        for k in range(nb):
            a = self._args(ids[k * B:(k + 1) * B], ratings[k * B:(k + 1) * B], B)
            a.gradU, a.gradV, a.gradK = _lib.ptr(self._gU), _lib.ptr(self._gV), _lib.ptr(self._gK)
            a.loss = loss.data_ptr() + 8 * k
            _lib.check(eng.lib.cf_svd_grads(a, stream), 'cf_svd_grads')
            self._apply(eng.U, eng.accU, self._gU, eng.n_users, eng.d, eng.ld)
            self._apply(eng.V, eng.accV, self._gV, eng.n_items, eng.d, eng.ld)
            self._apply(self.K, self.accK, self._gK, eng.d, eng.d, eng.ld)
            eng.launches += 4
        return loss

    # ------------------------------------------------------------------ evaluation
    def _predict_device(self, ids):
        eng, torch = self.engine, self.engine.torch
        ids = eng._as_i32(ids)
        if ids.dim() != 2 or ids.shape[1] != 2:
            raise ValueError('pairs must be [n, 2]')
        n = int(ids.shape[0])
        out = torch.empty(n, dtype=torch.float32, device=eng.device)
        if n:
            a = self._args(ids, None, n)
            _lib.check(eng.lib.cf_svd_predict_pairs(a, _lib.ptr(out), torch.cuda.current_stream(eng.device).cuda_stream),
                       'cf_svd_predict_pairs')
            eng.launches += 1
        return out

    def predict_pairs(self, useritem):
        out = self._predict_device(useritem)
        self.engine.check_flags()
        return out.cpu().numpy()

    def evaluate(self, tst_tuple):
        torch = self.engine.torch
        if torch.is_tensor(tst_tuple):
            ids, truth = tst_tuple[:, :2], tst_tuple[:, 2]
        else:
            tst_tuple = np.asarray(tst_tuple)
            ids, truth = tst_tuple[:, :2].astype(np.int32), tst_tuple[:, 2].astype(np.float64)
        scores = rating.evaluate(truth, self._predict_device(ids), self.eval_metrics, clip=self.range_of_ratings)
        self.engine.check_flags()
        return scores

    def recommend_device(self, users, topN, train_csr=None):
        raise NotImplementedError('SVD is a rating model (svd.py has no recommend)')
