"""Shared driver logic of the four model classes: the epoch loop, full-catalog recommendation and evaluation of
reference src/models/pl/models/bprmf.py:90-170 (identical skeleton in cml.py, gbprmf.py, basic/models/wrmf.py)."""
import datetime as dt
import os

from ..engine import FactorEngine
from ..metrics import ranking
from ..sparse import DeviceCSR


class RankingModelBase(object):
    _kind = None
    _print_prefix_fold = True

    def _setup(self, n_users, n_items, topN, split_method, eval_metrics, n_factors, batch_size, max_iter, lr,
               init_mean, init_stddev, device, optimizer, update, seed, verbose, **hyper):
        self.n_users, self.n_items, self.topN = int(n_users), int(n_items), int(topN)
        self.split_method, self.eval_metrics = split_method, list(eval_metrics)
        self.n_factors, self.batch_size = int(n_factors), int(batch_size)
        self.max_iter, self.lr = int(max_iter), float(lr)
        self.verbose = verbose
        self.engine = FactorEngine(self._kind, n_users, n_items, n_factors, device, init_mean, init_stddev,
                                   optimizer=optimizer, update=update, seed=seed, lr=float(lr), **hyper)
        self.device = self.engine.device
        self._printed_lr = float(lr)   # the reference decays a float it only prints (bprmf.py:159)
        self._train_csr = None

    # ------------------------------------------------------------------ parameters
    def state_dict(self):
        return self.engine.state_dict()

    def load_state_dict(self, sd):
        self.engine.load_state_dict(sd)

    # ------------------------------------------------------------------ one sess.run(train_op)
    def step(self, *batch):
        """One minibatch update on explicit index arrays (numpy or torch); returns the loss as a float.
        BPRMF/CML: step(pairs[B,2], negs[B,W]); GBPRMF: step(pairs, negs, group[B,G]); WRMF: step(uir[B,3])."""
        loss = self._train_arrays(batch, rows_per_batch=None)
        self.engine.check_flags()
        return float(loss[0].item())

    def _train_arrays(self, batch, rows_per_batch):
        raise NotImplementedError

    # ------------------------------------------------------------------ evaluation
    def _as_csr(self, m):
        return m if isinstance(m, DeviceCSR) else DeviceCSR.from_scipy(m, self.device)

    def recommend_device(self, users, topN, train_csr=None):
        """[T, topN] int32 CUDA tensor: best unseen items per user by (score desc, id asc); -1 pads short rows."""
        K = int(topN)
        if K > 1024:
            raise ValueError('topN up to 1024 is supported')
        return self.engine.topk(users, K, train_csr)

    def recommend(self, users, topN=None, trasR=None):
        """Public form of the reference's private ``__recommend`` (bprmf.py:90-103): list of lists of item ids."""
        tr = self._as_csr(trasR) if trasR is not None else self._train_csr
        idx = self.recommend_device(users, topN or self.topN, tr).cpu().numpy()
        return [[int(x) for x in row if x >= 0] for row in idx]

    def predict(self, users):
        """Public form of ``__predict__`` (bprmf.py:77-81): dense scores [len(users), n_items] (numpy float64)."""
        return self.engine.scores(users).cpu().numpy()

    def _eval(self, truth, pred_idx, k=None):
        k = k or self.topN
        if self.split_method == 'cv':
            return ranking.evaluateCV(truth, pred_idx, self.eval_metrics, k)
        elif self.split_method == 'loov':
            return ranking.evaluateLOOV(truth, pred_idx, self.eval_metrics, k)
        return None

    # ------------------------------------------------------------------ the epoch loop
    def _epoch(self, sampler, n_batches):
        """All minibatches of one epoch; returns the per-minibatch losses (CUDA float64)."""
        torch = self.engine.torch
        losses = []
        done = 0
        if hasattr(sampler, 'next_chunk'):
            rows = getattr(sampler, 'rows_per_batch', self.batch_size)
            chunk = max(1, min(n_batches, (1 << 22) // max(1, rows * 8)))
            sizes = [min(chunk, n_batches - lo) for lo in range(0, n_batches, chunk)]
            # The sampler launches run on a side stream, ahead of the training stream: a batch is a pure function of (seed,
            # epoch, batch index) and reads nothing the steps write -- the reference's producer threads overlap the same way
            # (sampler_ranking.py:40-50).  Two schedules:
            #   'eager'  chunk j + 1 is sampled while chunk j trains (under its fused step kernel);
            #   'late'   chunk j + 2 is sampled once the fused step kernel of chunk j has finished (cf_step_args.
            #            event_after_step), i.e. under the staged apply of j and the counting kernel of j + 1.
            # Measured on configs[1]'s shape (B = 2^20): eager: BPR W=1 1.68 -> 1.62 ms per minibatch, GBPR 3.37 -> 3.17 ms, but
            # CML 2.68 -> 2.77 ms (its step kernel needs all four resident blocks per SM and the sampler's blocks take their
            # slots).  Default: eager, late for CML; CF_SAMPLE_OVERLAP=0 (one stream) / 1 (eager) / 2 (late) forces one.
            want = os.environ.get('CF_SAMPLE_OVERLAP', '')
            mode = {'0': None, '1': 'eager', '2': 'late'}.get(want, 'late' if self._kind == 'cml' else 'eager')
            if len(sizes) < 2 or self.engine.device.type != 'cuda':
                mode = None
            if mode is None:
                for n in sizes:
                    losses.append(self._train_arrays(sampler.next_chunk(n), rows))
                return torch.cat(losses)
            main = torch.cuda.current_stream(self.engine.device)
            if getattr(self, '_sample_stream', None) is None:
                # CF_SAMPLE_PRIORITY=1: a high-priority stream -- the sampler's blocks take the block slots the training
                # kernels free before those kernels' own next blocks do
                prio = -1 if os.environ.get('CF_SAMPLE_PRIORITY', '0') == '1' else 0
                self._sample_stream = torch.cuda.Stream(self.engine.device, priority=prio)
            side = self._sample_stream
            side.wait_stream(main)                       # (nothing sampled here may start before what precedes the epoch)

            # index buffers: a ring of `ahead + 1` persistent slots inside the sampler (no allocation per minibatch).  A slot is
            # rewritten only after the chunk that used it has trained: eager -- the side stream waits for that chunk's
            # `trained` event; late -- it waits for the step kernel of a LATER chunk anyway.
            ahead = 2 if mode == 'late' else 1
            ring = ahead + 1 if hasattr(sampler, 'ring_slot') else 0
            trained = [None] * max(ring, 1)

            def sample(j):
                with torch.cuda.stream(side):
                    if ring:
                        sampler.ring_slot = j % ring
                        if trained[j % ring] is not None:
                            side.wait_event(trained[j % ring])
                    try:
                        arrays = sampler.next_chunk(sizes[j])
                    finally:
                        if ring:
                            sampler.ring_slot = None
                    ev = torch.cuda.Event()
                    ev.record(side)
                return arrays, ev
            queue = [sample(j) for j in range(min(ahead, len(sizes)))]
            for j in range(len(sizes)):
                arrays, ev = queue.pop(0)
                if mode == 'eager' and j + 1 < len(sizes):
                    queue.append(sample(j + 1))
                main.wait_event(ev)
                if not ring:
                    for t in arrays:
                        if t is not None:
                            t.record_stream(main)        # allocated on the side stream, consumed on this one
                if mode == 'late' and j + 2 < len(sizes):
                    stepped = torch.cuda.Event()
                    self.engine.pending_after_step_event = stepped
                losses.append(self._train_arrays(arrays, rows))
                if ring:
                    trained[j % ring] = torch.cuda.Event()
                    trained[j % ring].record(main)
                if mode == 'late' and j + 2 < len(sizes):
                    side.wait_event(stepped)             # the fused step kernel of chunk j is done: sample chunk j + 2
                    queue.append(sample(j + 2))
            main.wait_stream(side)
        else:   # any object with the reference's next_batch() (numpy arrays): upload batch by batch
            while done < n_batches:
                batch = sampler.next_batch()
                batch = batch if isinstance(batch, tuple) else (batch,)
                losses.append(self._train_arrays(batch, None))
                done += 1
        return torch.cat(losses)

    def _prepare_eval(self, trasR, tstsR):
        torch = self.engine.torch
        tra, tst = self._as_csr(trasR), self._as_csr(tstsR)
        self._train_csr = tra
        test_users = torch.nonzero(tst.row_lengths() > 0).reshape(-1)          # bprmf.py:117 (ascending)
        truth = tst.select_rows(test_users)
        if self.split_method == 'loov':                                          # bprmf.py:121-122: first test item
            truth = DeviceCSR(torch.arange(len(test_users) + 1, device=self.device, dtype=torch.int64),
                              truth.indices[truth.indptr[:-1]].contiguous(), None, None, truth.shape)
        return tra, test_users.to(torch.int32), truth

    def _log_line(self, fold, it, aveloss, scores):
        return ("%s_fold=%d iter=%2d: " % (self.split_method, fold, it + 1) +
                " TraLoss=%.4f lr=%.4f" % (aveloss, self._printed_lr) +
                ' \tTst@' + str(self.topN) + ':' + ' '.join(
                    [m + '=%.4f' % s for m, s in zip(self.eval_metrics, scores)]))

    def train(self, fold, trasR, tstsR, sampler):
        """Reference entry point (bprmf.py:113-170): max_iter epochs of int(nnz/B) minibatches, evaluation + one
        printed line per epoch, returns the last epoch's scores."""
        tra, test_users, truth = self._prepare_eval(trasR, tstsR)
        n_batches = int(tra.nnz / self.batch_size)
        scores = None
        for it in range(self.max_iter):
            t0 = dt.datetime.now()
            losses = self._epoch(sampler, n_batches)
            aveloss = float(losses.mean().item())
            self.engine.check_flags()
            if hasattr(sampler, 'check_flags'):
                sampler.check_flags()
            t1 = dt.datetime.now()
            pred = self.recommend_device(test_users, min(self.topN, self.n_items), tra)
            scores = self._eval(truth, pred)
            if self.verbose:
                print(self._format_epoch(fold, it, aveloss, scores, t0, t1))
            self._printed_lr *= .98
        scores = self._after_training(fold, tra, test_users, truth, scores)
        return scores

    def _format_epoch(self, fold, it, aveloss, scores, t0, t1):
        return self._log_line(fold, it, aveloss, scores)

    def _after_training(self, fold, tra, test_users, truth, scores):
        return scores

    def close(self):
        """The reference closes its tf.Session here (bprmf.py:172-173); device memory is released with the object."""
        self.engine._ws = None
