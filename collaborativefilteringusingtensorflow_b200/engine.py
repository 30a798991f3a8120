"""Device state + kernel launches shared by the four model classes (host side of the C ABI).

Holds what the reference keeps in TF variables and Adagrad slots (bprmf.py:29-34,86): the embedding tables,
their accumulators and -- new here -- the per-row occurrence workspace of the fused step kernel.
torch is used for device memory and streams only; all arithmetic is in libcf_b200.so.
"""
import numpy as np

from . import _lib
from .sparse import null_csr

_MODEL_IDS = {'bpr': _lib.MODEL_BPR, 'cml': _lib.MODEL_CML, 'gbpr': _lib.MODEL_GBPR, 'wrmf': _lib.MODEL_WRMF}
_SCORE_KIND = {'bpr': _lib.SCORE_DOT, 'cml': _lib.SCORE_NEG_SQDIST, 'gbpr': _lib.SCORE_DOT_BIAS, 'wrmf': _lib.SCORE_DOT}
_OPT = {'adagrad': _lib.OPT_ADAGRAD, 'sgd': _lib.OPT_SGD}
_UPD = {'sync': _lib.UPDATE_SYNC, 'hogwild': _lib.UPDATE_HOGWILD}

ADAGRAD_ACC0 = 0.1   # tf.train.AdagradOptimizer initial_accumulator_value


def resolve_device(device):
    """The reference's ``device='CPU'|'GPU'`` (bprmf.py:26-27) picks a TF device; here everything runs on CUDA."""
    torch = _lib.require_cuda()
    if isinstance(device, torch.device):
        return device
    s = str(device)
    if s.upper() in ('CPU', 'GPU'):
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device(s)


def pad_ld(d):
    return (int(d) + 3) // 4 * 4


class FactorEngine(object):
    def __init__(self, kind, n_users, n_items, n_factors, device='GPU', init_mean=0.0, init_stddev=0.1,
                 optimizer='adagrad', update='sync', seed=None, **hyper):
        torch = _lib.require_cuda()
        self.torch = torch
        self.lib = _lib.lib()
        self.kind = kind
        self.model_id = _MODEL_IDS[kind]
        self.device = resolve_device(device)
        self.n_users, self.n_items, self.d = int(n_users), int(n_items), int(n_factors)
        self.ld = pad_ld(n_factors)
        if self.ld > 512:
            raise ValueError('n_factors up to 512 are supported')
        self.optimizer, self.update = optimizer, update
        self.hyper = dict(lr=0.1, reg=0.0, margin=0.0, clip_norm=1.0, rho=0.0, weight=1.0, use_rank_weight=False)
        self.hyper.update(hyper)
        gen = torch.Generator(device=self.device)
        gen.manual_seed(int(seed) if seed is not None else int(np.random.SeedSequence().generate_state(1)[0]))
        self.U = self._init_table(self.n_users, init_mean, init_stddev, gen, truncated=(kind != 'cml'))
        self.V = self._init_table(self.n_items, init_mean, init_stddev, gen, truncated=(kind != 'cml'))
        self.b = None
        if kind == 'gbpr':   # gbprmf.py:36-38
            self.b = torch.empty(self.n_items, device=self.device)
            torch.nn.init.trunc_normal_(self.b, init_mean, init_stddev, init_mean - 2 * init_stddev,
                                        init_mean + 2 * init_stddev, generator=gen)
        self.accU = torch.full_like(self.U, ADAGRAD_ACC0)
        self.accV = torch.full_like(self.V, ADAGRAD_ACC0)
        self.accb = torch.full_like(self.b, ADAGRAD_ACC0) if self.b is not None else None
        self.counters = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._ws = None
        self._ws_rows = 0
        self._needs_full_clip = (kind == 'cml')
        self.launches = 0

    # ------------------------------------------------------------------ state
    def _init_table(self, n, mean, std, gen, truncated):
        torch = self.torch
        t = torch.zeros(n, self.ld, device=self.device)
        view = t[:, :self.d]
        if truncated:   # tf.truncated_normal_initializer (bprmf.py:30): resample beyond 2 sigma
            torch.nn.init.trunc_normal_(view, mean, std, mean - 2 * std, mean + 2 * std, generator=gen)
        else:           # tf.random_normal_initializer (cml.py:33)
            view.normal_(mean, std, generator=gen)
        return t

    def state_dict(self):
        out = {'U': self.U[:, :self.d].clone(), 'V': self.V[:, :self.d].clone(),
               'accU': self.accU[:, :self.d].clone(), 'accV': self.accV[:, :self.d].clone()}
        if self.b is not None:
            out['b'], out['accb'] = self.b.clone(), self.accb.clone()
        return out

    def load_state_dict(self, sd):
        torch = self.torch
        for name in ('U', 'V', 'accU', 'accV'):
            if name in sd:
                dst = getattr(self, name)
                src = torch.as_tensor(np.asarray(sd[name]) if not torch.is_tensor(sd[name]) else sd[name],
                                      dtype=torch.float32, device=self.device)
                if tuple(src.shape) != (dst.shape[0], self.d):
                    raise ValueError('%s: expected shape %s, got %s' % (name, (dst.shape[0], self.d), tuple(src.shape)))
                dst[:, :self.d].copy_(src)
        for name in ('b', 'accb'):
            if name in sd and getattr(self, name) is not None:
                getattr(self, name).copy_(torch.as_tensor(np.asarray(sd[name]) if not torch.is_tensor(sd[name])
                                                          else sd[name], dtype=torch.float32, device=self.device))
        self._needs_full_clip = (self.kind == 'cml')

    # ------------------------------------------------------------------ workspace
    def _workspace(self, B, W, G):
        torch = self.torch
        rows = int(self.lib.cf_step_staging_rows(self.model_id, B, W, G))
        if self._ws is None or self._ws_rows < rows:
            self._ws = dict(
                metaU=torch.zeros(self.n_users, dtype=torch.int32, device=self.device),
                metaV=torch.zeros(self.n_items, dtype=torch.int32, device=self.device),
                slot_row=torch.full((rows,), -1, dtype=torch.int32, device=self.device),   # CF_SLOT_EMPTY
                slotU=torch.zeros(self.n_users, dtype=torch.int32, device=self.device),
                slotV=torch.zeros(self.n_items, dtype=torch.int32, device=self.device),
                staging=torch.zeros(rows, self.ld + 4, device=self.device))
            self._ws_rows = rows
        return self._ws

    def reset_workspace(self):
        if self._ws is not None:
            for k in ('metaU', 'metaV', 'staging'):
                self._ws[k].zero_()
            self._ws['slot_row'].fill_(-1)
        self.counters.zero_()

    def check_flags(self):
        """Synchronises; raises on a device-side condition recorded by a previous launch."""
        f = int(self.counters[1].item())
        if f:
            self.reset_workspace()
            msgs = []
            if f & _lib.FLAG_INDEX_RANGE:
                msgs.append('batch index out of range (TF would raise InvalidArgumentError in the gather)')
            if f & _lib.FLAG_STAGING_FULL:
                msgs.append('gradient staging overflow')
            if f & _lib.FLAG_SAMPLER_GAVEUP:
                msgs.append('a user has (almost) every item as a positive: no negative found '
                            '(the reference sampler would spin forever, sampler_ranking.py:35)')
            raise RuntimeError('; '.join(msgs) or 'device flag %d' % f)

    # ------------------------------------------------------------------ training
    def _as_i32(self, x, cols=None):
        torch = self.torch
        if x is None:
            return None
        if not torch.is_tensor(x):
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
        x = x.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
        return x

    def train_batches(self, pairs, negs=None, group=None, ratings=None, batch_size=None, want_loss=True, profile=None,
                      grad_tables=None, after_step_event=None):
        """Run ``n = rows / batch_size`` consecutive minibatches (the inner loop of bprmf.py:143-148).
        Index arrays may be numpy or torch (any int dtype); returns the per-minibatch loss as a CUDA float64
        tensor (or None).  ``grad_tables=(gU, gV, gb)`` (dense, zeroed; gb only for GBPR) switches to the replicated
        data-parallel mode: one minibatch whose row gradients are summed into the tables, nothing is applied
        (dist.ReplicatedTrainer all-reduces and applies them)."""
        torch = self.torch
        pairs = self._as_i32(pairs)
        rows = int(pairs.shape[0])
        B = int(batch_size or rows)
        if rows == 0 or rows % B:
            raise ValueError('rows (%d) must be a positive multiple of batch_size (%d)' % (rows, B))
        nb = rows // B
        W = G = 0
        if self.kind != 'wrmf':
            negs = self._as_i32(negs)
            if negs is None or negs.dim() != 2 or negs.shape[0] != rows:
                raise ValueError('negs must be [rows, W]')
            W = int(negs.shape[1])
        if self.kind == 'gbpr':
            group = self._as_i32(group)
            if group is None or group.dim() != 2 or group.shape[0] != rows:
                raise ValueError('group must be [rows, G]')
            G = int(group.shape[1])
        if self.kind == 'wrmf':
            if ratings is None:
                raise ValueError('WRMF needs ratings')
            if not torch.is_tensor(ratings):
                ratings = torch.from_numpy(np.ascontiguousarray(np.asarray(ratings, dtype=np.float32)))
            ratings = ratings.to(device=self.device, dtype=torch.float32).contiguous()
        a = _lib.StepArgs()
        a.U, a.V, a.b = _lib.ptr(self.U), _lib.ptr(self.V), _lib.ptr(self.b)
        a.accU, a.accV, a.accb = _lib.ptr(self.accU), _lib.ptr(self.accV), _lib.ptr(self.accb)
        a.n_users, a.n_items, a.d, a.ld = self.n_users, self.n_items, self.d, self.ld
        a.pairs, a.negs, a.group = _lib.ptr(pairs), _lib.ptr(negs), _lib.ptr(group)
        a.ratings = _lib.ptr(ratings) if self.kind == 'wrmf' else None
        a.B, a.W, a.G, a.n_batches = B, W, G, nb
        a.model, a.optimizer, a.update = self.model_id, _OPT[self.optimizer], _UPD[self.update]
        h = self.hyper
        a.use_rank_weight = int(bool(h['use_rank_weight']))
        a.lr, a.reg, a.margin, a.clip_norm = h['lr'], h['reg'], h['margin'], h['clip_norm']
        a.rho, a.weight = h['rho'], h['weight']
        if self.update == 'sync':
            ws = self._workspace(B, W, G)
            a.metaU, a.metaV = _lib.ptr(ws['metaU']), _lib.ptr(ws['metaV'])
            a.slotU, a.slotV, a.slot_row = _lib.ptr(ws['slotU']), _lib.ptr(ws['slotV']), _lib.ptr(ws['slot_row'])
            a.staging, a.staging_rows = _lib.ptr(ws['staging']), ws['staging'].shape[0]
        a.counters = _lib.ptr(self.counters)
        loss = torch.zeros(nb, dtype=torch.float64, device=self.device) if want_loss else None
        a.loss = _lib.ptr(loss)
        if after_step_event is None:          # (or left by the epoch loop for the next call: models/_base.py::_epoch)
            after_step_event, self.pending_after_step_event = getattr(self, 'pending_after_step_event', None), None
        if after_step_event is not None:      # torch.cuda.Event recorded after the last minibatch's fused step kernel
            after_step_event.record(torch.cuda.current_stream(self.device))   # (creates the lazily-made handle; re-recorded by the library)
            a.event_after_step = after_step_event.cuda_event
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if grad_tables is not None:
            if nb != 1 or self.update != 'sync':
                raise ValueError('the replicated mode takes one minibatch per call, update="sync"')
            gU, gV, gb = grad_tables
            a.gradU, a.gradV, a.gradb = _lib.ptr(gU), _lib.ptr(gV), _lib.ptr(gb)
            _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
            self.launches += 1
            return loss
        if self._needs_full_clip and nb > 1:
            # the reference clips BOTH WHOLE tables after every step (cml.py:119-129); after the first such clip
            # every row has norm <= clip_norm and clipping only the touched rows (fused) is the same thing
            a.n_batches = 1
            _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
            self._full_clip(stream)
            a.n_batches = nb - 1
            a.pairs = pairs.data_ptr() + 4 * 2 * B
            a.negs = negs.data_ptr() + 4 * W * B
            a.loss = (loss.data_ptr() + 8) if want_loss else None
            _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
        elif profile is not None:   # bench.py: per-kernel CUDA-event times (synchronises)
            import ctypes as C
            mc, ms, ma = C.c_float(0), C.c_float(0), C.c_float(0)
            _lib.check(self.lib.cf_train_steps_profiled(a, stream, C.byref(mc), C.byref(ms), C.byref(ma)),
                       'cf_train_steps_profiled')
            profile['count_ms'] = profile.get('count_ms', 0.0) + mc.value
            profile['step_ms'] = profile.get('step_ms', 0.0) + ms.value
            profile['apply_ms'] = profile.get('apply_ms', 0.0) + ma.value
            profile['n_batches'] = profile.get('n_batches', 0) + nb
        else:
            _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
            if self._needs_full_clip:
                self._full_clip(stream)
        self.launches += nb * (3 if self.update == 'sync' else 1)
        return loss

    def _full_clip(self, stream):
        c = float(self.hyper['clip_norm'])
        _lib.check(self.lib.cf_clip_rows(_lib.ptr(self.U), self.n_users, self.d, self.ld, c, stream), 'cf_clip_rows')
        _lib.check(self.lib.cf_clip_rows(_lib.ptr(self.V), self.n_items, self.d, self.ld, c, stream), 'cf_clip_rows')
        self._needs_full_clip = False
        self.launches += 2

    # ------------------------------------------------------------------ evaluation
    def _topk_args(self, users, K, train_csr, item_range=None):
        torch = self.torch
        a = _lib.TopkArgs()
        a.U, a.V, a.b = _lib.ptr(self.U), _lib.ptr(self.V), _lib.ptr(self.b)
        a.n_users, a.n_items, a.d, a.ld = self.n_users, self.n_items, self.d, self.ld
        if users is None:
            users_t, T = None, self.n_users
        else:
            users_t = self._as_i32(users)
            T = int(users_t.numel())
            if T and (int(users_t.min()) < 0 or int(users_t.max()) >= self.n_users):
                raise ValueError('user id out of range')
        a.users, a.T, a.K, a.kind = _lib.ptr(users_t), T, int(K), _SCORE_KIND[self.kind]
        a.train = train_csr.as_c(False) if train_csr is not None else null_csr()
        a.flags = _lib.ptr(self.counters[1:2])
        a.item_lo, a.item_hi = item_range if item_range is not None else (0, 0)
        return a, users_t, T

    TENSOR_MIN_WORK = 1 << 31      # T * n_items above which method='auto' takes the tensor-core path

    def topk(self, users, K, train_csr=None, return_values=False, item_range=None, method='auto', debug_scores=False):
        """Masked top-K item ids [T, K] int32 (and fp64 scores): bprmf.py:90-103 in one pass.
        method: 'exact' (fp64 CUDA cores), 'tensor' (tcgen05 fp16 candidate pass + exact re-rank; identical results,
        K <= 1024 -- above 200 in rounds of 200 --, d <= 254) or 'auto' (tensor for large problems)."""
        torch = self.torch
        if K <= 0:
            raise ValueError('K must be positive')
        a, users_t, T = self._topk_args(users, K, train_csr, item_range)
        if T == 0:
            raise ValueError('no query users')
        tensor_ok = K <= 1024 and self.d <= 254 and item_range is None
        if method == 'auto':
            method = 'tensor' if tensor_ok and T * self.n_items >= self.TENSOR_MIN_WORK else 'exact'
        if method == 'tensor' and not tensor_ok:
            raise ValueError('the tensor-core top-K path needs K <= 1024, n_factors <= 254 and no item range')
        out_idx = torch.empty(T, K, dtype=torch.int32, device=self.device)
        out_val = torch.empty(T, K, dtype=torch.float64, device=self.device) if return_values else None
        a.out_idx, a.out_val = _lib.ptr(out_idx), _lib.ptr(out_val)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if method == 'tensor':
            need = int(self.lib.cf_topk_tc_workspace_bytes(a))
            if need < 0:
                raise RuntimeError(self.lib.cf_last_error().decode())
            if getattr(self, '_tc_ws', None) is None or self._tc_ws.numel() < need + 1024:
                self._tc_ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
            base = (self._tc_ws.data_ptr() + 1023) // 1024 * 1024
            dbg = None
            if debug_scores:
                dbg = torch.zeros(T, (self.n_items + 255) // 256 * 256, dtype=torch.float32, device=self.device)
            if getattr(self, 'tc_stats', None) is None:
                self.tc_stats = torch.zeros(4, dtype=torch.int32, device=self.device)   # see cf_b200.h: cf_topk_tc stats
            _lib.check(self.lib.cf_topk_tc(a, base, need, _lib.ptr(dbg), _lib.ptr(self.tc_stats), stream), 'cf_topk_tc')
            rounds = (K + 199) // 200
            self.launches += 3 + 2 * rounds + (3 * rounds - 1 if rounds > 1 else 0)
            if debug_scores:
                return out_idx, out_val, dbg
        else:
            _lib.check(self.lib.cf_topk_exact(a, stream), 'cf_topk_exact')
            self.launches += 1
        return (out_idx, out_val) if return_values else out_idx

    def als_half_sweep(self, side, csr):
        """Solve all user rows (side='users', csr = user -> items) or all item rows (side='items', csr = item -> users)
        of the weighted-ALS normal equations (cf_als_half_sweep); hyper: weight, reg."""
        torch = self.torch
        if self.d > 128:
            raise ValueError('the ALS solver supports n_factors <= 128')
        X, Y = (self.U, self.V) if side == 'users' else (self.V, self.U)
        n_x, n_y = X.shape[0], Y.shape[0]
        if csr.shape != (n_x, n_y):
            raise ValueError('CSR shape %s does not match (%d, %d)' % (csr.shape, n_x, n_y))
        need = int(self.lib.cf_als_workspace_bytes(n_y))
        if getattr(self, '_als_ws', None) is None or self._als_ws.numel() < need + 1024:
            self._als_ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        a = _lib.AlsArgs()
        a.X, a.Y, a.n_x, a.n_y, a.d, a.ldx, a.ldy = _lib.ptr(X), _lib.ptr(Y), n_x, n_y, self.d, self.ld, self.ld
        a.indptr, a.indices = _lib.ptr(csr.indptr), _lib.ptr(csr.indices)
        a.weight, a.reg = float(self.hyper['weight']), float(self.hyper['reg'])
        a.workspace = (self._als_ws.data_ptr() + 1023) // 1024 * 1024
        a.workspace_bytes = need
        _lib.check(self.lib.cf_als_half_sweep(a, torch.cuda.current_stream(self.device).cuda_stream), 'cf_als_half_sweep')
        self.launches += 3

    def als_gram(self, Y, G):
        """G[128, 128] += Y^T Y (cf_als_gram); the Gram stage of the half-sweep on a slice of rows."""
        torch = self.torch
        n_y = int(Y.shape[0])
        need = int(self.lib.cf_als_workspace_bytes(n_y))
        if getattr(self, '_als_ws', None) is None or self._als_ws.numel() < need + 1024:
            self._als_ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        ws = (self._als_ws.data_ptr() + 1023) // 1024 * 1024
        _lib.check(self.lib.cf_als_gram(_lib.ptr(Y), n_y, self.d, int(Y.stride(0)), _lib.ptr(G), ws, need,
                                        torch.cuda.current_stream(self.device).cuda_stream), 'cf_als_gram')
        self.launches += 2

    def als_solve_rows(self, X, Y, csr, G):
        """Solves the rows of X (a row-range view is fine) from the complete Gram G and their observed rows of Y."""
        torch = self.torch
        if self.d > 128:
            raise ValueError('the ALS solver supports n_factors <= 128')
        if csr.shape != (X.shape[0], Y.shape[0]):
            raise ValueError('CSR shape %s does not match (%d, %d)' % (csr.shape, X.shape[0], Y.shape[0]))
        a = _lib.AlsArgs()
        a.X, a.Y, a.n_x, a.n_y, a.d = _lib.ptr(X), _lib.ptr(Y), int(X.shape[0]), int(Y.shape[0]), self.d
        a.ldx, a.ldy = int(X.stride(0)), int(Y.stride(0))
        a.indptr, a.indices = _lib.ptr(csr.indptr), _lib.ptr(csr.indices)
        a.weight, a.reg = float(self.hyper['weight']), float(self.hyper['reg'])
        # with a workspace the rows with few observed columns take the low-rank (Woodbury) path of cf_als.cu
        need = int(self.lib.cf_als_workspace_bytes(int(Y.shape[0])))
        if getattr(self, '_als_ws', None) is None or self._als_ws.numel() < need + 1024:
            self._als_ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        a.workspace = (self._als_ws.data_ptr() + 1023) // 1024 * 1024
        a.workspace_bytes = need
        _lib.check(self.lib.cf_als_solve_rows(a, _lib.ptr(G), torch.cuda.current_stream(self.device).cuda_stream), 'cf_als_solve_rows')
        self.launches += 4

    def predict_pairs(self, pairs):
        """Scores of explicit (user, item) rows (``__predict`` of mf.py:66-72 fed with ``tst_tuple[:, :-1]``): float32
        CUDA tensor [n]; pairs is int [n, 2] (numpy or torch)."""
        torch = self.torch
        pairs = self._as_i32(pairs)
        if pairs.dim() != 2 or pairs.shape[1] != 2:
            raise ValueError('pairs must be [n, 2]')
        n = int(pairs.shape[0])
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        if n == 0:
            return out
        _lib.check(self.lib.cf_predict_pairs(_lib.ptr(self.U), _lib.ptr(self.V), _lib.ptr(self.b), self.n_users, self.n_items,
                                             self.d, self.ld, _SCORE_KIND[self.kind], _lib.ptr(pairs), n, _lib.ptr(out),
                                             _lib.ptr(self.counters), torch.cuda.current_stream(self.device).cuda_stream),
                   'cf_predict_pairs')
        self.launches += 1
        return out

    def scores(self, users):
        """Dense [T, n_items] fp64 score matrix of ``__predict__`` (small inputs only)."""
        torch = self.torch
        a, users_t, T = self._topk_args(users, 1, None)
        out = torch.empty(T, self.n_items, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.cf_scores(a, _lib.ptr(out), stream), 'cf_scores')
        self.launches += 1
        return out


def rank_metrics_device(pred_idx, truth_csr, k):
    """Per-user metric values [T, 8] (float64, CUDA) = {pre, recall, ndcg, map, mrr, hit, rr, n_pred}."""
    torch = _lib.require_cuda()
    lib = _lib.lib()
    pred_idx = pred_idx.contiguous()
    T, ldp = int(pred_idx.shape[0]), int(pred_idx.shape[1])
    if truth_csr.shape[0] != T or T == 0 or k <= 0:
        raise ValueError('len(yss_true) != len(yss_pred) or len(yss_true)==0 or k<=0!')
    out = torch.empty(T, 8, dtype=torch.float64, device=pred_idx.device)
    stream = torch.cuda.current_stream(pred_idx.device).cuda_stream
    _lib.check(lib.cf_rank_metrics(_lib.ptr(pred_idx), T, ldp, int(k), _lib.ptr(truth_csr.indptr),
                                   _lib.ptr(truth_csr.indices), _lib.ptr(out), stream), 'cf_rank_metrics')
    return out
