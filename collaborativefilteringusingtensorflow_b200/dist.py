"""Multi-GPU (one process per GPU, torch.distributed) sharding of the hot path -- SURVEY.md section 8(e).

The reference is single-device (bprmf.py:131 uses DEVICES[0] only); this is new.

Training: users are range-sharded with their CSR rows (sampling and the user-row update are local), the item table is
row-sharded by ``item % P`` (local row ``item // P``).  Per minibatch every rank
  1. samples B local pairs + negatives (global item ids),
  2. dedupes the item ids it needs and asks their owners for the rows       (all_to_all: counts, ids, rows),
  3. runs the fused step kernel in exchange mode: user rows are updated in place, item-row gradients are summed into a
     compact buffer aligned with the fetched rows,
  4. returns the gradient rows to the owners                                (all_to_all: rows),
  5. owners sum what they received per row and apply it once (cf_apply_rows), i.e. exactly the single-GPU
     minibatch-synchronous semantics with global batch P*B.
Peer-pull variant (``item_transport='peer'``; NVLink / NVSwitch, one node): step 2's row transfer disappears -- every
rank maps the other ranks' item shards (CUDA IPC) and the fused kernel reads each item row straight from its owner's
memory; only ids and gradient rows still travel through NCCL.  It moves one row per OCCURRENCE instead of one per
unique id, so it wins when minibatches barely repeat items (huge catalogues: BASELINE configs[4]) and loses when they
do (configs[1]); ``'auto'`` measures the repeat ratio on the first minibatch and picks.
Small tables (GBPR's configs[2]: 42 MB) are not sharded at all: ``ReplicatedTrainer`` keeps every table on every GPU, each
rank turns its B pairs into dense gradient tables, ONE all_reduce sums them and every rank applies the same update.
Evaluation: every rank scores its item shard for all query users (their embeddings are all-gathered), keeps a local
top-K, and the [T, K] lists are all-gathered and merged (cf_topk_merge).

``ItemExchange`` is pure index bookkeeping + collectives on whatever device its tensors live on, so it is tested on CPU
with the gloo backend (tests/test_dist_gloo.py); the kernels plug in at steps 3 and 5.
"""
from . import _lib


class PeerMappingError(RuntimeError):
    """CUDA IPC mapping of a peer's buffer failed on some rank (raised on EVERY rank, after a collective vote)."""


class ExchangePlan(object):
    __slots__ = ('n_req', 'occ_local', 'send_counts', 'recv_counts', 'recv_local_rows', 'req_global', 'send_rows',
                 'peer_offsets')


class ItemExchange(object):
    """Routes item-row requests between ranks. Owner of item i is ``i % world``; its local row is ``i // world``."""

    def __init__(self, world, rank, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world, self.rank, self.group = int(world), int(rank), group

    def _a2a(self, out, inp, out_split, in_split):
        if self.world == 1:
            out.copy_(inp)
        else:
            self.dist.all_to_all_single(out, inp, out_split, in_split, group=self.group)

    def plan(self, item_ids, n_items_global=None):
        return self.plan_exchange(self.plan_local(item_ids, n_items_global))

    def plan_local(self, item_ids, n_items_global=None):
        """Collective-free half of the plan (dedupe + grouping by owner).  item_ids: int tensor (any shape) of GLOBAL item ids needed by this rank's minibatch.
        With ``n_items_global`` the dedupe is sort-free (flag array laid out [owner][local row] + prefix sum: O(items +
        occurrences)); without it falls back to torch.unique."""
        torch = self.torch
        flat = item_ids.reshape(-1).to(torch.int64)
        P = self.world
        if n_items_global is not None:
            L = (int(n_items_global) + P - 1) // P                         # rows per owner (padded)
            key = (flat % P) * L + flat // P                                # position in the owner-major layout
            flag = torch.zeros(P * L, dtype=torch.int32, device=flat.device)
            flag[key] = 1
            csum = torch.cumsum(flag, 0, dtype=torch.int32)
            occ = (csum[key] - 1).to(torch.int32)                           # occurrence -> row of the fetched buffer
            upos = torch.nonzero(flag).reshape(-1)                          # owner-major, ascending local row
            send_rows = (upos % L).to(torch.int32)
            ends = csum[torch.arange(1, P + 1, device=flat.device) * L - 1].to(torch.int64)
            send_counts = torch.diff(ends, prepend=torch.zeros(1, dtype=torch.int64, device=flat.device))
            req_global = (upos % L) * P + upos // L
        else:
            uniq, inv = torch.unique(flat, return_inverse=True)             # sorted unique ids + occurrence -> unique
            owner = uniq % P
            order = torch.sort(owner, stable=True).indices                  # group the requests by owner
            pos = torch.empty_like(order)
            pos[order] = torch.arange(order.numel(), device=order.device)
            send_counts = torch.bincount(owner, minlength=P)
            send_rows = (uniq[order] // P).to(torch.int32)                  # local row ids at the owners
            occ = pos[inv].to(torch.int32)
            req_global = uniq[order]
        p = ExchangePlan()
        p.n_req = int(send_rows.numel())
        p.occ_local = occ.reshape(item_ids.shape)
        p.send_counts = send_counts
        p.send_rows = send_rows
        p.req_global = req_global
        return p

    def plan_exchange(self, p, all_counts=False):
        """Collective half: tell every owner how many and which of its rows this rank needs.  With ``all_counts`` the whole
        P x P count matrix is all-gathered, which also tells an owner WHERE its segment starts in every requester's
        buffers (``peer_offsets``: what the owner-pull apply needs to read the gradient rows in place)."""
        torch = self.torch
        p.peer_offsets = None
        if all_counts:
            M = torch.empty(self.world * self.world, dtype=p.send_counts.dtype, device=p.send_counts.device)
            if self.world == 1:
                M[:] = p.send_counts
            else:
                self.dist.all_gather_into_tensor(M, p.send_counts.contiguous(), group=self.group)
            M = M.view(self.world, self.world).tolist()
            sc, rc = M[self.rank], [M[q][self.rank] for q in range(self.world)]
            p.peer_offsets = [sum(M[q][:self.rank]) for q in range(self.world)]
        else:
            recv_counts = torch.empty_like(p.send_counts)
            self._a2a(recv_counts, p.send_counts, None, None)
            sc, rc = p.send_counts.tolist(), recv_counts.tolist()
        recv_rows = torch.empty(sum(rc), dtype=torch.int32, device=p.send_rows.device)
        self._a2a(recv_rows, p.send_rows, rc, sc)
        p.send_counts, p.recv_counts = sc, rc
        p.recv_local_rows = recv_rows                                       # rows of MY shard that others (and I) asked for
        return p

    def fetch(self, plan, table_shard):
        """Returns the requested rows [n_req, ld], ordered like plan.occ_local indexes them."""
        torch = self.torch
        send = table_shard.index_select(0, plan.recv_local_rows.to(torch.int64))
        out = torch.empty(plan.n_req, table_shard.shape[1], dtype=table_shard.dtype, device=table_shard.device)
        self._a2a(out, send, plan.send_counts, plan.recv_counts)
        return out

    def push(self, plan, grad_rows):
        """Sends one gradient row per fetched row back to its owner; returns [n_recv, ld] aligned with
        plan.recv_local_rows (a row requested by several ranks appears once per requester)."""
        torch = self.torch
        out = torch.empty(int(plan.recv_local_rows.numel()), grad_rows.shape[1], dtype=grad_rows.dtype,
                          device=grad_rows.device)
        self._a2a(out, grad_rows.contiguous(), plan.recv_counts, plan.send_counts)
        return out


def item_shard_rows(n_items_global, world, rank):
    return (int(n_items_global) - rank + world - 1) // world


class DistributedTrainer(object):
    """Row-sharded BPRMF / CML training over ``world`` GPUs.  ``model`` is a model object built with
    n_users = this rank's users and n_items = this rank's item-shard rows; ``sampler`` samples this rank's CSR (columns =
    GLOBAL item ids).

    ``item_transport``:
      'fetch' / 'peer' / 'auto'  the DEVICE-SIDE exchange (csrc/cf_exchange.cu): hand-written kernels over CUDA-IPC peer
               memory; nothing returns to the host, NCCL only provides the two barriers of a minibatch.  'fetch' gathers
               every requested row ONCE into a local buffer (right when minibatches repeat items: configs[1]); 'peer' lets
               the fused step kernel read item rows from their owners per occurrence (right when they barely repeat:
               configs[4]); 'auto' measures the repeat ratio of the first minibatch and picks (same answer on every rank).
               In both, the owners read the gradient rows in place from the requesters' compact buffers.  'peer-push' is
               'peer' with the gradients going the other way: k_step red.adds every occurrence's gradient straight into
               the owner's dense gradient table over NVLink (both link directions busy at once).  Measured on configs[4]
               at 2 GPUs it LOSES (3.60 vs 3.31 ms per minibatch: the dense tables turn the L2-resident atomics of the
               compact buffer into DRAM read-modify-writes on both sides), so it is not what 'auto' picks.
      'replicate'  for catalogues that are SMALL against the minibatch (configs[1]: 6.3 M item occurrences per GPU and
               minibatch over 500 k items -- every rank requests almost every row every minibatch, so the row exchange moves
               the whole table and its gradients anyway, plus dedupe, routing, two barriers and an owner-side scatter): every
               rank keeps a replica of the whole item table (built once from the shards, laid out shard after shard; the
               engine's own table becomes a view of its block), the fused step red.adds item-row gradients into a dense table
               of the same layout, ONE NCCL reduce-scatter hands every rank the summed gradients of its shard,
               `cf_apply_dense` applies them there (the accumulators stay sharded) and ONE all-gather spreads the updated rows.
               Users stay sharded.  Measured on configs[1] at 8 GPUs with an all-reduce + full apply on every rank: 13.4 G triple
               updates/s against 11.4 G for the row exchange.  'auto' takes it when B * (1 + W) >= 2 * n_items_global.
      'nccl'   the first version, kept as the portable baseline: torch-op plan + three NCCL all-to-alls (ids, rows,
               gradients).  It is also what 'auto' falls back to, with a warning, when peer memory cannot be mapped."""

    def __init__(self, model, sampler, n_items_global, world, rank, group=None, item_transport='nccl'):
        if item_transport not in ('nccl', 'peer', 'peer-push', 'fetch', 'auto', 'replicate'):
            raise ValueError("item_transport must be 'nccl', 'peer', 'peer-push', 'fetch', 'replicate' or 'auto'")
        self.torch = _lib.require_cuda()
        self.lib = _lib.lib()
        self.model, self.eng, self.sampler = model, model.engine, sampler
        if self.eng.kind not in ('bpr', 'cml'):
            raise ValueError('sharded training supports BPRMF and CML')
        self.n_items_global = int(n_items_global)
        self.world, self.rank = int(world), int(rank)
        if self.eng.n_items != item_shard_rows(n_items_global, world, rank):
            raise ValueError('model.n_items must be the item-shard size %d' % item_shard_rows(n_items_global, world, rank))
        self.ex = ItemExchange(world, rank, group)
        # the routing (dedupe + grouping by owner) of minibatch k+1 runs on a side stream while minibatch k computes
        self.side = self.torch.cuda.Stream(device=self.eng.device)
        self._keep = []
        self._ows = None
        self._ows_rows = 0
        self.launches = 0
        self.bytes_sent = 0           # payload that went through NCCL ('nccl' transport)
        self.bytes_pulled = 0         # payload read from peer memory by our kernels: an upper bound counted on the host ...
        self.req_rows_dev = None      # ... and the exact number of unique rows requested so far (device scalar, int64)
        self.occurrences = 0
        self.step_events = None       # set to [] to collect one CUDA event per minibatch of step()
        self.phase_ms = None          # set to {} to collect per-phase CUDA-event times (synchronises every phase)
        self.item_transport = item_transport
        self.peer_ptrs = None         # device pointers of every rank's item shard (this rank's own included)
        self._opened = {}             # IPC handle bytes -> mapped base (an allocation is mapped once per process)
        self._pull = False
        self._push = False
        self._dev = None              # buffers of the device-side exchange (allocated on the first minibatch)
        self._k = 0
        self._bar = None              # preallocated 1-element tensor of the named cross-GPU barrier
        self._replicate = True if item_transport == 'replicate' else None if item_transport == 'auto' else False
        self._rep = None              # 'replicate': the item-table replica, its accumulators and the dense gradient table
        self.device_side = item_transport not in ('nccl', 'replicate')
        n_neg = getattr(sampler, 'n_neg', None)
        if n_neg is not None and self._use_replica(int(sampler.batch_size), int(n_neg)):
            self.device_side = False
        if self.device_side:
            if self.world > _lib.MAX_PEERS:
                raise ValueError('the device-side exchange supports up to %d GPUs on one node' % _lib.MAX_PEERS)
            self._pull = {'peer': True, 'peer-push': True, 'fetch': False, 'auto': None}[item_transport]
            self._push = item_transport == 'peer-push'
            if n_neg is not None:     # otherwise the buffers are shared on the first minibatch
                self._setup_or_fall_back(int(sampler.batch_size), int(n_neg))

    # ------------------------------------------------------------------ plumbing
    def _barrier(self, why):
        """Named cross-GPU barrier ON THE COMPUTE STREAM: a 1-element NCCL all_reduce enqueued on the current stream
        completes only after every rank has enqueued it, i.e. after everything each rank enqueued before it.  The peer
        transports rely on that order (remote NVLink reads of mailboxes / item rows / gradient rows against the owners'
        applies and the requesters' zeroing), so the process group must be NCCL -- a gloo collective synchronises the
        HOSTS and does not order the streams.  The one exception is opt-in (CF_DIST_HOST_BARRIER=1, set by the one-GPU
        form of tests/dist_check.py, where NCCL cannot run): synchronise the device, then a host barrier -- every rank's
        device work is complete before any rank goes on, which orders at least as much (and costs a pipeline drain)."""
        if self.world == 1:
            return
        dist = self.ex.dist
        if dist.get_backend(self.ex.group) != 'nccl':
            import os
            if os.environ.get('CF_DIST_HOST_BARRIER') != '1':
                raise RuntimeError('peer transport needs an NCCL process group (barrier "%s" orders CUDA streams)' % why)
            self.torch.cuda.synchronize()
            dist.barrier(group=self.ex.group)
            return
        if self._bar is None:
            self._bar = self.torch.zeros(1, device=self.eng.device)
        dist.all_reduce(self._bar, group=self.ex.group)

    def _share(self, t):
        """Exchanges CUDA IPC handles of tensor ``t`` (one per rank) and maps the peers' tensors into this process;
        returns the ``world`` device pointers (this rank's own pointer at its own index).  Collective."""
        import ctypes as C
        torch, dist = self.torch, self.ex.dist
        dev = self.eng.device
        if self.world == 1:
            return [t.data_ptr()]
        # every rank runs both collectives whatever happens locally; a failure anywhere is voted and raised everywhere
        err = None
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        try:
            _lib.check(self.lib.cf_ipc_export(_lib.ptr(t), C.addressof(handle), C.byref(off)), 'cf_ipc_export')
        except RuntimeError as e:
            err = e
        import os
        if os.environ.get('CF_IPC_FAIL_RANK') == str(self.rank):      # test hook: exercises the vote + NCCL fallback
            err = RuntimeError('CF_IPC_FAIL_RANK test hook')
        mine = torch.tensor(list(handle) + [(off.value >> (8 * k)) & 255 for k in range(8)], dtype=torch.uint8, device=dev)
        every = torch.empty(self.world * 72, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every, mine, group=self.ex.group)
        every = every.view(self.world, 72).cpu().numpy()
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(t.data_ptr())
                continue
            if err is not None:
                continue
            key = (r, bytes(every[r, :64].tolist()))
            base = self._opened.get(key)
            if base is None:
                h = (C.c_ubyte * 64)(*every[r, :64].tolist())
                b = C.c_void_p(0)
                try:
                    _lib.check(self.lib.cf_ipc_open(C.addressof(h), C.byref(b)), 'cf_ipc_open')
                except RuntimeError as e:
                    err = e
                    continue
                base = self._opened[key] = b.value
            ptrs.append(base + int.from_bytes(bytes(every[r, 64:72].tolist()), 'little'))
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.ex.group)
        if int(ok.item()) == 0:
            self.close()
            raise PeerMappingError(str(err) if err is not None else 'another rank could not map a peer buffer')
        return ptrs

    def close(self):
        """Unmaps the peers' buffers (call on every rank before the tables are freed)."""
        for b in self._opened.values():
            self.lib.cf_ipc_close(b)
        self._opened, self.peer_ptrs = {}, None
        if self._dev is not None:
            self._dev = None

    def _setup_or_fall_back(self, B, W):
        try:
            self._setup_device_exchange(B, W)
        except PeerMappingError as e:
            if self.item_transport != 'auto':
                raise
            import warnings
            warnings.warn('peer memory is not available (%s): item rows and gradients travel through NCCL' % e)
            self.close()
            self.device_side, self._pull, self._dev = False, False, None

    def _setup_device_exchange(self, B, W):
        """Allocates and IPC-shares the buffers of the device-side exchange, sized once for minibatches of B pairs with W
        negatives.  Collective."""
        torch, eng, P = self.torch, self.eng, self.world
        dev = eng.device
        occ = B * (1 + W)
        L = (self.n_items_global + P - 1) // P
        cap = min(occ, L)
        slots = min(occ, self.n_items_global)
        pad = lambda n: (n + 63) // 64 * 64
        # ONE allocation holds everything the peers read besides the item shard: two mailboxes (minibatch parity) of
        # [counts int32[64] | req int32[P, cap]] and the compact gradient buffer [slots, ld] (float32)
        mail = 64 + pad(P * cap)
        comm = torch.zeros(2 * mail + slots * eng.ld, dtype=torch.int32, device=dev)
        d = dict(B=B, W=W, L=L, cap=cap, slots=slots, comm=comm, mail=mail)
        d['gbuf'] = comm[2 * mail:].view(torch.float32).view(slots, eng.ld)
        d['slot_of'] = torch.full((P * L,), -1, dtype=torch.int32, device=dev)
        d['fetched'] = None            # allocated when the fetch transport is chosen
        d['slot_pairs'] = [torch.zeros(B, 2, dtype=torch.int32, device=dev) for _ in range(2)]
        d['slot_negs'] = [torch.zeros(B, W, dtype=torch.int32, device=dev) for _ in range(2)]
        d['slot_pos'] = [torch.zeros(B, dtype=torch.int32, device=dev) for _ in range(2)]
        srows = min(P * cap, 2 * occ + 1024)       # expected receipts of an owner: occ (items hash uniformly over owners)
        d['srows'] = srows
        d['meta'] = torch.zeros(eng.n_items, dtype=torch.int32, device=dev)
        d['slot'] = torch.zeros(eng.n_items, dtype=torch.int32, device=dev)
        d['slot_row'] = torch.full((srows,), -1, dtype=torch.int32, device=dev)
        d['staging'] = torch.zeros(srows, eng.ld + 4, device=dev)
        d['segs'] = torch.zeros(4 * _lib.MAX_PEERS + 4, dtype=torch.int64, device=dev)
        d['routed'] = [None, None]
        comm_ptrs = self._share(comm)
        self.peer_ptrs = self._share(eng.V)
        d['comm_ptrs'] = comm_ptrs
        self.req_rows_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._dev = d

    def _xargs(self, par, pairs=None, negs=None):
        d, eng, P = self._dev, self.eng, self.world
        a = _lib.ExchangeArgs()
        a.n_ranks, a.rank, a.n_items_global, a.cap = P, self.rank, self.n_items_global, d['cap']
        a.pairs, a.negs, a.B, a.W = _lib.ptr(pairs), _lib.ptr(negs), d['B'], d['W']
        a.slot_of = _lib.ptr(d['slot_of'])
        a.slot_pairs, a.slot_negs, a.slot_pos = _lib.ptr(d['slot_pairs'][par]), _lib.ptr(d['slot_negs'][par]), _lib.ptr(d['slot_pos'][par])
        for r in range(P):
            base = d['comm_ptrs'][r]
            a.counts[r] = base + 4 * par * d['mail']
            a.req[r] = base + 4 * (par * d['mail'] + 64)
            a.grads[r] = base + 4 * 2 * d['mail']
            a.tables[r] = self.peer_ptrs[r]
        a.fetched = _lib.ptr(d['fetched'])
        a.d, a.ld = eng.d, eng.ld
        a.table, a.acc, a.n_rows = _lib.ptr(eng.V), _lib.ptr(eng.accV), eng.n_items
        a.model, a.optimizer = eng.model_id, 0 if eng.optimizer == 'adagrad' else 1
        a.lr, a.clip_norm = eng.hyper['lr'], eng.hyper['clip_norm']
        a.meta, a.slot, a.slot_row = _lib.ptr(d['meta']), _lib.ptr(d['slot']), _lib.ptr(d['slot_row'])
        a.staging, a.staging_rows, a.segs = _lib.ptr(d['staging']), d['srows'], _lib.ptr(d['segs'])
        a.counters = _lib.ptr(eng.counters)
        return a

    def _route(self, pairs, negs, par):
        """Dedupe + routing of one minibatch on the CURRENT stream (fills mailbox ``par`` and the slot arrays ``par``)."""
        torch = self.torch
        lp, ln = pairs.to(torch.int32).contiguous(), negs.to(torch.int32).contiguous()
        a = self._xargs(par, lp, ln)
        _lib.check(self.lib.cf_exchange_route(a, torch.cuda.current_stream(self.eng.device).cuda_stream), 'cf_exchange_route')
        self._dev['routed'][par] = (lp, ln)       # keep the (possibly converted) index arrays alive until the step has run
        self.launches += 2

    def _decide_transport(self, par):
        """'auto': pull when the minibatches of all ranks together repeat items so rarely that one row per occurrence
        over NVLink is cheaper than gathering one row per unique id first (same answer on every rank).  One host
        synchronisation, on the first minibatch only."""
        torch, d = self.torch, self._dev
        counts = d['comm'][par * d['mail']: par * d['mail'] + self.world]
        t = torch.stack([counts.sum().double(), torch.tensor(float(d['B'] * (1 + d['W'])), dtype=torch.float64, device=counts.device)])
        if self.world > 1:
            self.ex.dist.all_reduce(t, group=self.ex.group)
        uniq, occ = t.tolist()
        self._pull = uniq > 0.5 * occ
        return self._pull

    def _owner_workspace(self, n):
        torch, eng = self.torch, self.eng
        if self._ows is None or self._ows_rows < n:
            rows = int(n * 1.25) + 1024
            self._ows = dict(meta=torch.zeros(eng.n_items, dtype=torch.int32, device=eng.device),
                             slot=torch.zeros(eng.n_items, dtype=torch.int32, device=eng.device),
                             slot_row=torch.full((rows,), -1, dtype=torch.int32, device=eng.device),
                             staging=torch.zeros(rows, eng.ld + 4, device=eng.device))
            self._ows_rows = rows
        return self._ows

    def _tick(self, name, ev0):
        if self.phase_ms is None:
            return None
        torch = self.torch
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        ev1.synchronize()
        if ev0 is not None:
            self.phase_ms[name] = self.phase_ms.get(name, 0.0) + ev0.elapsed_time(ev1)
        return ev1

    def make_plan(self, pairs, negs, local_only=False):
        torch = self.torch
        items = torch.cat([pairs[:, 1:2].to(torch.int64), negs.to(torch.int64)], dim=1)       # [B, 1 + W] global ids
        p = self.ex.plan_local(items, self.n_items_global)
        return p if local_only else self.ex.plan_exchange(p)

    def _step_args(self, B, W, want_loss):
        torch, eng = self.torch, self.eng
        a = _lib.StepArgs()
        a.U, a.accU, a.accV = _lib.ptr(eng.U), _lib.ptr(eng.accU), _lib.ptr(eng.accV)
        a.n_users, a.d, a.ld = eng.n_users, eng.d, eng.ld
        a.B, a.W, a.G, a.n_batches = B, W, 0, 1
        a.model, a.optimizer, a.update = eng.model_id, 0 if eng.optimizer == 'adagrad' else 1, _lib.UPDATE_SYNC
        h = eng.hyper
        a.use_rank_weight = int(bool(h['use_rank_weight']))
        a.lr, a.reg, a.margin, a.clip_norm, a.rho, a.weight = h['lr'], h['reg'], h['margin'], h['clip_norm'], h['rho'], h['weight']
        ws = eng._workspace(B, W, 0)
        a.metaU, a.metaV = _lib.ptr(ws['metaU']), _lib.ptr(ws['metaV'])
        a.slotU, a.slotV, a.slot_row = _lib.ptr(ws['slotU']), _lib.ptr(ws['slotV']), _lib.ptr(ws['slot_row'])
        a.staging, a.staging_rows = _lib.ptr(ws['staging']), ws['staging'].shape[0]
        a.counters = _lib.ptr(eng.counters)
        loss = torch.zeros(1, dtype=torch.float64, device=eng.device) if want_loss else None
        a.loss = _lib.ptr(loss)
        a.rank_items = self.n_items_global
        return a, loss

    # ------------------------------------------------------------------ one minibatch
    def step_chunk(self, pairs, negs, batch_size, want_loss=True, plan=None, routed=False, after_prepare=None):
        """One minibatch (rows == batch_size) of the sharded step on explicit local batches."""
        B = int(batch_size)
        if int(pairs.shape[0]) != B:
            raise ValueError('the sharded step takes one minibatch per call')
        if self._use_replica(B, int(negs.shape[1])):
            return self._step_chunk_replica(pairs, negs, B, want_loss)
        if self.device_side and self._dev is None:
            self._setup_or_fall_back(B, int(negs.shape[1]))
        if self.device_side:
            return self._step_chunk_device(pairs, negs, B, want_loss, routed, after_prepare)
        return self._step_chunk_nccl(pairs, negs, B, want_loss, plan)

    # ------------------------------------------------------------------ 'replicate': item table replicated, dense gradients all-reduced
    def _use_replica(self, B, W):
        """'auto' decides from sizes alone (so every rank decides alike, without a collective): the replica wins when a
        rank's minibatch holds at least twice as many item occurrences as the catalogue has rows."""
        if self._replicate is None:
            self._replicate = self.world > 1 and B * (1 + W) >= 2 * self.n_items_global
            if self._replicate:
                self.device_side = False
        return self._replicate

    def _setup_replica(self):
        """Builds the replica from the shards (collective, once).  Layout: the P shards one after the other (item i sits at
        row (i % P) * L + i // P, L = ceil(n_items / P)), so that a rank's shard is one contiguous block -- the unit of the
        reduce-scatter and of the all-gather -- and the engine's own item table becomes a VIEW of that block: evaluation and
        state_dict() read current rows without any copy."""
        torch, eng, P = self.torch, self.eng, self.world
        L = (self.n_items_global + P - 1) // P
        if P == 1:
            V = eng.V
        else:
            V = torch.zeros(P * L, eng.ld, device=eng.device)
            V[self.rank * L:self.rank * L + eng.n_items].copy_(eng.V)
            self.ex.dist.all_gather_into_tensor(V, V[self.rank * L:(self.rank + 1) * L], group=self.ex.group)
            eng.V = V[self.rank * L:self.rank * L + eng.n_items]
        self._rep = dict(V=V, L=L, g=torch.zeros(P * L, eng.ld, device=eng.device),
                         gblk=torch.zeros(L, eng.ld, device=eng.device) if P > 1 else None)

    def _replica_rows(self, ids):
        """global item id -> row of the replica"""
        P = self.world
        if P == 1:
            return ids
        q = ids // P
        return (ids - q * P) * self._rep['L'] + q

    def _wait_replica_comm(self):
        """The all-gather of the previous minibatch's updated item rows (and the zeroing of the gradient table) run on the
        communication stream; whoever reads the item table or starts the next fused step on the compute stream waits here."""
        if getattr(self, '_comm_pending', False):
            self.torch.cuda.current_stream(self.eng.device).wait_stream(self._comm)
            self._comm_pending = False

    def _step_chunk_replica(self, pairs, negs, B, want_loss, defer_gather_wait=False):
        """``defer_gather_wait``: leave the all-gather of the updated rows running on the communication stream when this
        call returns (step() does that between consecutive minibatches; the next call, or step()'s end, waits for it)."""
        torch, eng, P = self.torch, self.eng, self.world
        import os
        if self._rep is None:
            self._setup_replica()
        rep = self._rep
        L = rep['L']
        pairs, negs = eng._as_i32(pairs), eng._as_i32(negs)
        W = int(negs.shape[1])
        main = torch.cuda.current_stream(eng.device)
        stream = main.cuda_stream
        # Two streams (CF_REPLICA_OVERLAP=0 puts everything back on one): the reduce-scatter of the item gradients starts when
        # the fused step kernel has finished (cf_step_args.event_after_step) and runs under the staged apply of the USER rows;
        # the gradient table is zeroed and the updated rows are all-gathered on the communication stream while the compute
        # stream goes on to whatever precedes the next fused step.  The phase timer synchronises after every phase anyway.
        overlap = P > 1 and self.phase_ms is None and os.environ.get('CF_REPLICA_OVERLAP', '1') != '0'
        if overlap and getattr(self, '_comm', None) is None:
            self._comm, self._comm_pending = torch.cuda.Stream(eng.device), False
            self._ev_step, self._ev_rs, self._ev_applied = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        ev = self._tick('', None)
        if P > 1:
            pairs = torch.stack([pairs[:, 0], self._replica_rows(pairs[:, 1])], 1).contiguous()
            negs = self._replica_rows(negs).contiguous()
        a, loss = self._step_args(B, W, want_loss)
        # the fused step in exchange mode: user rows are counted / staged / applied locally, every item-row gradient is
        # red.added into the dense table (row = the item's row in the replica)
        a.V, a.n_items = _lib.ptr(rep['V']), P * L
        a.pairs, a.negs, a.gradV = _lib.ptr(pairs), _lib.ptr(negs), _lib.ptr(rep['g'])
        self._wait_replica_comm()                   # the replica is complete and the gradient table is zero again
        if overlap:
            self._ev_step.record(main)              # (creates the handle; the library re-records it after the fused step kernel)
            a.event_after_step = self._ev_step.cuda_event
        _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
        ev = self._tick('k_count + k_step + k_apply_staged', ev)
        g_mine = rep['g']
        if P > 1:       # every rank receives the summed gradients of ITS shard
            if overlap:
                self._comm.wait_event(self._ev_step)
                with torch.cuda.stream(self._comm):
                    self.ex.dist.reduce_scatter_tensor(rep['gblk'], rep['g'], group=self.ex.group)
                    self._ev_rs.record(self._comm)
                    rep['g'].zero_()                # after the reduce-scatter has read it; under the owner apply
                main.wait_event(self._ev_rs)
            else:
                self.ex.dist.reduce_scatter_tensor(rep['gblk'], rep['g'], group=self.ex.group)
            g_mine = rep['gblk']
        ev = self._tick('reduce-scatter of the dense item gradients (NCCL)', ev)
        ap = _lib.ApplyArgs()
        ap.table, ap.acc, ap.n_rows, ap.d, ap.ld = _lib.ptr(eng.V), _lib.ptr(eng.accV), eng.n_items, eng.d, eng.ld
        ap.grads, ap.ldg = _lib.ptr(g_mine), eng.ld
        ap.model, ap.optimizer = eng.model_id, 0 if eng.optimizer == 'adagrad' else 1
        ap.lr, ap.clip_norm = eng.hyper['lr'], eng.hyper['clip_norm']
        _lib.check(self.lib.cf_apply_dense(ap, stream), 'cf_apply_dense')      # applies the touched rows (and re-zeroes their gradients)
        if eng._needs_full_clip:
            # cml.py:119-129: step first, THEN the whole-table clip (engine.train_batches): this rank's
            # users and its item shard -- the all-gather below spreads the clipped rows
            eng._full_clip(stream)
        ev = self._tick('owner apply (k_apply_dense on the shard)', ev)
        if P > 1:
            if overlap:
                self._ev_applied.record(main)
                self._comm.wait_event(self._ev_applied)
                with torch.cuda.stream(self._comm):
                    self.ex.dist.all_gather_into_tensor(rep['V'], rep['V'][self.rank * L:(self.rank + 1) * L], group=self.ex.group)
                self._comm_pending = True
                if not defer_gather_wait:
                    self._wait_replica_comm()
            else:
                self.ex.dist.all_gather_into_tensor(rep['V'], rep['V'][self.rank * L:(self.rank + 1) * L], group=self.ex.group)
                rep['g'].zero_()
        ev = self._tick('all-gather of the updated item rows (NCCL) + zeroing of the gradient table', ev)
        self.launches += 3 + 1
        self.occurrences += B * (1 + W)
        self.bytes_sent += int(2 * (P - 1) * L * eng.ld * 4)   # reduce-scatter + all-gather volume per rank
        return loss

    def _step_chunk_device(self, pairs, negs, B, want_loss, routed, after_prepare):
        torch, eng, d = self.torch, self.eng, self._dev
        W = int(negs.shape[1])
        if B != d['B'] or W != d['W']:
            raise ValueError('minibatch shape (%d, %d) differs from the first one (%d, %d)' % (B, W, d['B'], d['W']))
        par = self._k & 1
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        ev = self._tick('', None)
        if not routed:
            self._route(pairs, negs, par)
        lp, ln = d['routed'][par]
        ev = self._tick('route (dedupe + group by owner: k_route_assign, k_route_fill)', ev)
        pull = self._pull if self._pull is not None else self._decide_transport(par)
        if not pull and d['fetched'] is None:
            d['fetched'] = torch.empty(d['slots'], eng.ld, device=eng.device)
        push = pull and self._push
        if push and d.get('gown') is None:
            # 'peer-push': item rows arrive over NVLink inside k_step, and the gradient of every occurrence LEAVES in the
            # same kernel -- red.added straight into the owner's dense gradient table (the owners then apply their tables
            # locally).  Shared once, on the first minibatch (collective).
            d['gown'] = torch.zeros(eng.n_items, eng.ld, device=eng.device)
            d['touched'] = torch.zeros(eng.n_items, dtype=torch.int32, device=eng.device)
            d['gown_ptrs'] = self._share(d['gown'])
        self.req_rows_dev += d['comm'][par * d['mail']: par * d['mail'] + self.world].sum()
        # barrier 1: every mailbox is complete, and every rank is past the owner-side apply of the previous minibatch (so
        # item rows may be read remotely and the gradient buffer may be zeroed)
        self._barrier('mailboxes complete')
        ev = self._tick('barrier 1', ev)
        x = self._xargs(par)
        if pull:
            x.fetched = None
        if push:
            x.dense_grads, x.touched = _lib.ptr(d['gown']), _lib.ptr(d['touched'])
        _lib.check(self.lib.cf_exchange_prepare(x, stream), 'cf_exchange_prepare')
        ev = self._tick('prepare (fetch rows over NVLink + zero grads + owner-side count: k_owner_segs, k_exchange_prepare)', ev)
        if after_prepare is not None:
            after_prepare()                      # the dedupe table is at rest again: the next minibatch may be routed
        a, loss = self._step_args(B, W, want_loss)
        if pull:    # item rows straight from their owners' shards inside the fused kernel (global ids)
            a.V, a.n_items = _lib.ptr(eng.V), self.n_items_global
            a.pairs, a.negs = _lib.ptr(lp), _lib.ptr(ln)
            for r, q in enumerate(self.peer_ptrs):
                a.peerV[r] = q
                if push:
                    a.peerG[r] = d['gown_ptrs'][r]
            a.n_peers = self.world
            if not push:    # the compact gradient buffer, one row per unique requested item (read in place by the owners)
                a.gradV, a.gslot_pos, a.gslot_neg = _lib.ptr(d['gbuf']), _lib.ptr(d['slot_pos'][par]), _lib.ptr(d['slot_negs'][par])
        else:       # the fetched copy of every requested row, ids = rows of the compact buffers
            a.V, a.n_items = _lib.ptr(d['fetched']), d['slots']
            a.pairs, a.negs = _lib.ptr(d['slot_pairs'][par]), _lib.ptr(d['slot_negs'][par])
            a.gradV = _lib.ptr(d['gbuf'])
        _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
        ev = self._tick('k_count + k_step + k_apply_staged', ev)
        self._barrier('gradient buffers complete')          # every rank is past its step kernel
        ev = self._tick('barrier 2', ev)
        _lib.check(self.lib.cf_exchange_apply(x, stream), 'cf_exchange_apply')
        if eng._needs_full_clip:
            # cml.py:119-129 clips BOTH whole tables after every step; after the first such clip every row has norm <= clip
            # and the touched-row clip fused into the applies is the same thing.  Like
            # engine.train_batches: step first, THEN clip -- the first minibatch's gradients see the unclipped init.
            eng._full_clip(stream)
        ev = self._tick('owner apply (k_owner_apply_dense)' if push else 'owner apply (k_owner_scatter + k_apply_staged)', ev)
        self._k += 1
        self.launches += 2 + 3 + (1 if push else 2)
        self.occurrences += B * (1 + W)
        self.bytes_pulled += (B * (1 + W) if pull else 0) * eng.ld * 4
        return loss

    def _step_chunk_nccl(self, pairs, negs, B, want_loss, plan):
        torch, eng = self.torch, self.eng
        ev = self._tick('', None)
        if plan is None:
            plan = self.make_plan(pairs, negs, local_only=True)
        if getattr(plan, 'recv_local_rows', None) is None:
            plan = self.ex.plan_exchange(plan)
        ev = self._tick('plan (dedupe + route ids)', ev)
        Vbuf = self.ex.fetch(plan, eng.V)                                                # [n_req, ld]
        ev = self._tick('fetch rows (gather + all_to_all)', ev)
        Gbuf = torch.zeros_like(Vbuf)
        lp = torch.stack([pairs[:, 0].to(torch.int32), plan.occ_local[:, 0]], dim=1).contiguous()
        ln = plan.occ_local[:, 1:].contiguous()
        a, loss = self._step_args(B, int(negs.shape[1]), want_loss)
        a.V, a.n_items = _lib.ptr(Vbuf), plan.n_req
        a.pairs, a.negs = _lib.ptr(lp), _lib.ptr(ln)
        a.gradV = _lib.ptr(Gbuf)
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        ev = self._tick('prep (remap ids, zero grads)', ev)
        _lib.check(self.lib.cf_train_steps(a, stream), 'cf_train_steps')
        ev = self._tick('k_count + k_step + k_apply_staged', ev)
        recv = self.ex.push(plan, Gbuf)
        ev = self._tick('push grads (all_to_all)', ev)                                                  # [n_recv, ld]
        n = int(recv.shape[0])
        if n:
            ows = self._owner_workspace(n)
            ap = _lib.ApplyArgs()
            ap.table, ap.acc, ap.n_rows, ap.d, ap.ld = _lib.ptr(eng.V), _lib.ptr(eng.accV), eng.n_items, eng.d, eng.ld
            ap.rows, ap.grads, ap.n, ap.ldg = _lib.ptr(plan.recv_local_rows), _lib.ptr(recv), n, eng.ld
            ap.model, ap.optimizer, ap.lr, ap.clip_norm = eng.model_id, a.optimizer, eng.hyper['lr'], eng.hyper['clip_norm']
            ap.meta, ap.slot, ap.slot_row = _lib.ptr(ows['meta']), _lib.ptr(ows['slot']), _lib.ptr(ows['slot_row'])
            ap.staging, ap.staging_rows, ap.counters = _lib.ptr(ows['staging']), ows['staging'].shape[0], _lib.ptr(eng.counters)
            _lib.check(self.lib.cf_apply_rows(ap, stream), 'cf_apply_rows')
        if eng._needs_full_clip:   # step first, then the one-time whole-table clip (cml.py:119-129)
            eng._full_clip(stream)
        ev = self._tick('owner apply (cf_apply_rows)', ev)
        self.launches += 3 + 3
        self.occurrences += B * (1 + int(negs.shape[1]))
        self.bytes_sent += (plan.n_req + n) * eng.ld * 4 + plan.n_req * 4
        return loss

    def step(self, n_minibatches=1, want_loss=True):
        """Sample + run n minibatches; returns their losses (CUDA float64).  The routing of minibatch k+1 (dedupe + grouping
        by owner) runs on a side stream while minibatch k computes."""
        torch = self.torch
        B = self.sampler.batch_size
        main = torch.cuda.current_stream(self.eng.device)
        W_known = getattr(self.sampler, 'n_neg', None)
        import os
        if (self.world > 1 and W_known is not None and n_minibatches > 1 and self.phase_ms is None and self.step_events is None
                and self._use_replica(B, int(W_known)) and os.environ.get('CF_REPLICA_OVERLAP', '1') != '0'):
            # 'replicate' transport, pipelined: minibatch k + 1 is SAMPLED on the side stream while the updated item rows of
            # minibatch k are all-gathered on the communication stream (the sampler waits for the owner apply of k, i.e. it
            # starts with the all-gather: both are idle time for the SMs otherwise).  A batch is a pure function of (seed,
            # epoch, batch index), so the minibatches are the ones next_chunk(n) would return; the index buffers live in a
            # ring of three persistent slots inside the sampler (models/_base.py::_epoch does the same on one GPU).
            side, sampler = self.side, self.sampler
            ring = 3 if hasattr(sampler, 'ring_slot') else 0
            used = [None] * 3
            side.wait_stream(main)

            def sample(j, after=None):
                if after is not None:
                    side.wait_event(after)
                with torch.cuda.stream(side):
                    if ring:
                        sampler.ring_slot = j % ring
                        if used[j % ring] is not None:
                            side.wait_event(used[j % ring])
                    try:
                        arrays = sampler.next_chunk(1)
                    finally:
                        if ring:
                            sampler.ring_slot = None
                    done = torch.cuda.Event()
                    done.record(side)
                return arrays, done
            nxt = sample(0)
            out = []
            for k in range(n_minibatches):
                arrays, done = nxt
                main.wait_event(done)
                if not ring:
                    for t in arrays:
                        if t is not None:
                            t.record_stream(main)
                out.append(self._step_chunk_replica(arrays[0], arrays[1], B, want_loss, defer_gather_wait=k + 1 < n_minibatches))
                if ring:
                    used[k % ring] = torch.cuda.Event()
                    used[k % ring].record(main)
                if k + 1 < n_minibatches:
                    nxt = sample(k + 1, after=self._ev_applied)
            main.wait_stream(side)
            return torch.cat(out) if want_loss else None
        chunk = self.sampler.next_chunk(n_minibatches)
        if self._use_replica(B, int(chunk[1].shape[1])):
            out = []
            for k in range(n_minibatches):
                out.append(self._step_chunk_replica(chunk[0][k * B:(k + 1) * B], chunk[1][k * B:(k + 1) * B], B, want_loss,
                                                    defer_gather_wait=k + 1 < n_minibatches))
                if self.step_events is not None:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(main)
                    self.step_events.append(ev)
            return torch.cat(out) if want_loss else None
        overlap = self.phase_ms is None and self.world > 1

        def batch(k):
            return chunk[0][k * B:(k + 1) * B], chunk[1][k * B:(k + 1) * B]

        out = []
        if self.device_side and self._dev is None:
            self._setup_or_fall_back(B, int(chunk[1].shape[1]))
        if self.device_side:
            ev_routed = [None, None]

            def route_ahead(k):
                # on the side stream, after prepare(k - 1) has returned the dedupe table to rest
                ev = torch.cuda.Event()
                ev.record(main)
                self.side.wait_event(ev)
                with torch.cuda.stream(self.side):
                    self._route(*batch(k), (self._k + 1) & 1)
                    done = torch.cuda.Event()
                    done.record(self.side)
                ev_routed[(self._k + 1) & 1] = done

            for k in range(n_minibatches):
                p, n = batch(k)
                routed = False
                if overlap and k > 0:
                    main.wait_event(ev_routed[self._k & 1])
                    routed = True
                nxt = (lambda kk=k + 1: route_ahead(kk)) if overlap and k + 1 < n_minibatches else None
                out.append(self.step_chunk(p, n, B, want_loss, routed=routed, after_prepare=nxt))
                if self.step_events is not None:               # per-minibatch device timeline (no synchronisation)
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record(main)
                    self.step_events.append(ev)
            return torch.cat(out) if want_loss else None

        sampled = torch.cuda.Event()
        sampled.record(main)
        self.side.wait_event(sampled)      # the side stream only depends on the sampled indices, not on the steps

        def plan_async(k):
            with torch.cuda.stream(self.side):
                return self.make_plan(*batch(k), local_only=True)

        nxt = plan_async(0) if overlap else None
        for k in range(n_minibatches):
            plan = None
            if overlap:
                plan = nxt
                main.wait_stream(self.side)
                self._keep = (self._keep + [plan])[-3:]     # plans were allocated on the side stream: keep them alive
            p, n = batch(k)
            out.append(self.step_chunk(p, n, B, want_loss, plan))
            if self.step_events is not None:               # per-minibatch device timeline (no synchronisation)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(main)
                self.step_events.append(ev)
            if overlap and k + 1 < n_minibatches:
                nxt = plan_async(k + 1)
        return torch.cat(out) if want_loss else None


def shard_mask_csr(sub_csr, world, rank):
    """Rows of ``sub_csr`` (global item ids) restricted to the items this rank owns, in LOCAL ids (item // world)."""
    torch = _lib.require_cuda()
    from .sparse import DeviceCSR
    keep = (sub_csr.indices % world) == rank
    rows = sub_csr.rows[keep]
    cols = (sub_csr.indices[keep] // world).to(torch.int32)
    counts = torch.bincount(rows.to(torch.int64), minlength=sub_csr.shape[0])
    indptr = torch.zeros(sub_csr.shape[0] + 1, dtype=torch.int64, device=rows.device)
    indptr[1:] = torch.cumsum(counts, 0)
    return DeviceCSR(indptr, cols.contiguous(), rows.contiguous(), None, (sub_csr.shape[0], item_shard_rows(sub_csr.shape[1], world, rank)))


def distributed_topk(engine, query_rows, K, train_local_csr, world, rank, group=None, method='auto', gather=True):
    """query_rows: [T, ld] embeddings of the query users (already all-gathered / replicated on every rank).
    engine.V is this rank's item shard (local row j = global item j * world + rank); train_local_csr masks in LOCAL item
    ids (row t = query t).  Every rank keeps a local top-K over its shard; then
      gather=True   the [T, K] lists are all-gathered and every rank merges all of them: returns the global top-K ids
                    [T, K] (int32) and fp64 scores on EVERY rank;
      gather=False  the lists are exchanged with ONE all-to-all so that rank r receives the P lists of users
                    [r * c, (r + 1) * c), c = ceil(T / P), and merges only those: returns (lo, ids [c', K], scores) for its own
                    slice -- 1/P of the traffic and of the merge work per rank, which is what an evaluation needs (every user's
                    metrics are computed once, their sums all-reduced: ``distributed_evaluate``)."""
    torch = _lib.require_cuda()
    import torch.distributed as dist
    lib = _lib.lib()
    T = int(query_rows.shape[0])
    saved_U, saved_n = engine.U, engine.n_users
    engine.U, engine.n_users = query_rows.contiguous(), T
    try:
        idx, val = engine.topk(None, K, train_local_csr, return_values=True, method=method)
    finally:
        engine.U, engine.n_users = saved_U, saved_n
    gidx = torch.where(idx >= 0, idx * world + rank, idx)
    stream = torch.cuda.current_stream(idx.device).cuda_stream
    if world == 1:
        return (gidx, val) if gather else (0, gidx, val)
    if gather:
        all_i = torch.empty(world * T, K, dtype=torch.int32, device=idx.device)       # rank p's [T, K] list at rows [p * T, (p + 1) * T)
        all_v = torch.empty(world * T, K, dtype=torch.float64, device=idx.device)
        dist.all_gather_into_tensor(all_i, gidx.contiguous(), group=group)
        dist.all_gather_into_tensor(all_v, val.contiguous(), group=group)
        out_i = torch.empty(T, K, dtype=torch.int32, device=idx.device)
        out_v = torch.empty(T, K, dtype=torch.float64, device=idx.device)
        _lib.check(lib.cf_topk_merge(all_i.data_ptr(), all_v.data_ptr(), world, T, K, out_i.data_ptr(), out_v.data_ptr(), stream),
                   'cf_topk_merge')
        return out_i, out_v
    c = (T + world - 1) // world
    if c * world != T:      # pad the user dimension so that every rank sends equal chunks
        gidx = torch.cat([gidx, torch.full((c * world - T, K), -1, dtype=torch.int32, device=idx.device)])
        val = torch.cat([val, torch.full((c * world - T, K), float('-inf'), dtype=torch.float64, device=idx.device)])
    recv_i = torch.empty(world * c, K, dtype=torch.int32, device=idx.device)          # rank p's list of MY users at rows [p * c, (p + 1) * c)
    recv_v = torch.empty(world * c, K, dtype=torch.float64, device=idx.device)
    dist.all_to_all_single(recv_i, gidx.contiguous(), group=group)
    dist.all_to_all_single(recv_v, val.contiguous(), group=group)
    out_i = torch.empty(c, K, dtype=torch.int32, device=idx.device)
    out_v = torch.empty(c, K, dtype=torch.float64, device=idx.device)
    _lib.check(lib.cf_topk_merge(recv_i.data_ptr(), recv_v.data_ptr(), world, c, K, out_i.data_ptr(), out_v.data_ptr(), stream),
               'cf_topk_merge')
    lo = rank * c
    n_mine = max(0, min(T, lo + c) - lo)
    return lo, out_i[:n_mine], out_v[:n_mine]


def _gather_item_rows(t, n_items_global, world, group=None):
    """Row-sharded [rows of item % P == rank, ...] tensors of every rank -> the full tensor in item order (collective)."""
    torch = _lib.require_cuda()
    import torch.distributed as dist
    P = int(world)
    L = (int(n_items_global) + P - 1) // P
    send = t if t.shape[0] == L else torch.cat([t, torch.zeros((L - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)])
    got = torch.empty((P * L,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)      # rank p's rows at [p * L, (p + 1) * L)
    dist.all_gather_into_tensor(got, send.contiguous(), group=group)
    got = got.view((P, L) + tuple(t.shape[1:]))
    return got.transpose(0, 1).reshape((L * P,) + tuple(t.shape[1:]))[:n_items_global].contiguous()   # item l * P + p sits at [p][l]


def gather_item_table(engine, n_items_global, world, rank, group=None):
    """The row-sharded item table (local row j = global item j * world + rank) all-gathered into the full
    [n_items_global, ld] table in item order (and the item bias, if the model has one).  10 M items x 512 B = 5 GB: a
    fraction of one B200's 180 GB, which is what makes the user-sharded evaluation below possible."""
    P = int(world)
    if P == 1:
        return engine.V, engine.b
    V = _gather_item_rows(engine.V, n_items_global, P, group)
    b = _gather_item_rows(engine.b, n_items_global, P, group) if getattr(engine, 'b', None) is not None else None
    return V, b


def user_sharded_topk(engine, users_local, K, train_csr_local, n_items_global, world, rank, group=None, method='auto',
                      tables=None, return_values=True):
    """Evaluation with the USERS sharded (they already are: user rows and their CSR rows live on their owner rank) and the
    item table gathered once: every rank scores ITS users against the whole catalogue -- no per-rank restart of the
    thresholds, no candidate lists to merge, and it scales with the number of GPUs (the item-sharded form of SURVEY 8e,
    ``distributed_topk``, makes every rank sweep all users: 8 GPUs gave 2.1-2.7x one GPU).  train_csr_local: rows = this
    rank's users, columns = GLOBAL item ids.  ``tables`` = a cached (V, b) from gather_item_table (V changes only when
    training steps run).  Returns the rank's own lists (ids int32 [T, K], fp64 scores); metrics over all users are
    ``distributed_evaluate``'s all-reduced sums."""
    V, b = tables if tables is not None else gather_item_table(engine, n_items_global, world, rank, group)
    saved = engine.V, engine.b, engine.n_items
    engine.V, engine.b, engine.n_items = V, b, int(n_items_global)
    try:
        out = engine.topk(users_local, K, train_csr_local, return_values=return_values, method=method)
    finally:
        engine.V, engine.b, engine.n_items = saved
    return out


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def distributed_evaluate(truth_local, pred_local, eval_metrics, k=5, split_method='cv', group=None):
    """``evaluateCV`` / ``evaluateLOOV`` (reference metrics/ranking.py:94-120) over users sharded across ranks: every rank
    runs cf_rank_metrics on ITS users' lists, the per-metric sums and the user count are all-reduced (one
    all_reduce of 8 doubles), and every rank returns the global values: the mean over all users for CV metrics, the sum
    for LOOV ones, ``None`` for unknown names -- as ranking.py does on one process.  A rank may hold no users."""
    torch = _lib.require_cuda()
    import torch.distributed as dist
    from .metrics import ranking as R
    world, _ = _world(group)
    loov = split_method == 'loov'
    cols = R._LOOV_COL if loov else R._CV_COL
    dev = pred_local.device if torch.is_tensor(pred_local) else torch.device('cuda', torch.cuda.current_device())
    acc = torch.zeros(10, dtype=torch.float64, device=dev)           # 8 column sums, #users, #users with no truth
    n_local = len(truth_local)
    if n_local != len(pred_local) or k <= 0:
        raise ValueError(R._ERR)
    if n_local:
        vals, truth = R.per_user(truth_local, pred_local, k, loov=loov, device=dev)
        acc[:8] = vals.sum(0)
        acc[8] = n_local
        acc[9] = (truth.row_lengths() == 0).sum()
    if world > 1:
        dist.all_reduce(acc, group=group)
    acc = acc.cpu().numpy()
    if acc[8] == 0:
        raise ValueError(R._ERR)
    if not loov and 'map' in eval_metrics and acc[9] > 0:
        raise ZeroDivisionError('float division by zero')             # ranking.py:53 divides by len(yss_true[ind])
    den = 1.0 if loov else acc[8]
    return [(float(acc[cols[m]] / den) if m in cols else None) for m in eval_metrics]


class DistributedALS(object):
    """Weighted-ALS sweeps of WRMF over ``world`` GPUs: both factor tables are replicated, the rows a
    half-sweep solves are range-sharded (rank r solves rows [r * chunk, (r + 1) * chunk) with chunk = ceil(n / world)).
    Per half-sweep: partial Gram of the rank's slice of the FIXED side (tcgen05) -> all_reduce of the 128 x 128 Gram (64
    KB) -> solve the local rows (cf_als_solve_rows) -> all_gather of the solved rows.  ``engine`` holds the full U / V;
    ``user_csr`` / ``item_csr`` are the CSRs of the rank's OWN row range (columns = global ids of the other side); one of
    them may be None when only the other side's half-sweep is run (half_sweep of the missing side raises)."""

    def __init__(self, engine, user_csr, item_csr, group=None):
        self.torch = _lib.require_cuda()
        self.eng, self.group = engine, group
        self.world, self.rank = _world(group)
        self.csr = dict(users=user_csr, items=item_csr)
        self.G = self.torch.zeros(128, 128, device=engine.device)
        for side, n in (('users', engine.n_users), ('items', engine.n_items)):
            lo, hi = self.row_range(n, self.world, self.rank)
            if self.csr[side] is not None and self.csr[side].shape[0] != hi - lo:
                raise ValueError('%s CSR must hold rows [%d, %d) of this rank' % (side, lo, hi))

    @staticmethod
    def row_range(n, world, rank):
        chunk = (int(n) + world - 1) // world
        return min(rank * chunk, int(n)), min((rank + 1) * chunk, int(n))

    def half_sweep(self, side):
        torch, eng = self.torch, self.eng
        import torch.distributed as dist
        if self.csr[side] is None:
            raise ValueError('no %s CSR was given to this DistributedALS' % side)
        X, Y = (eng.U, eng.V) if side == 'users' else (eng.V, eng.U)
        n_x, n_y = int(X.shape[0]), int(Y.shape[0])
        self.G.zero_()
        ylo, yhi = self.row_range(n_y, self.world, self.rank)
        if yhi > ylo:
            eng.als_gram(Y[ylo:yhi], self.G)
        if self.world > 1:
            dist.all_reduce(self.G, group=self.group)
        xlo, xhi = self.row_range(n_x, self.world, self.rank)
        if xhi > xlo:
            eng.als_solve_rows(X[xlo:xhi], Y, self.csr[side], self.G)
        if self.world > 1:
            chunk = (n_x + self.world - 1) // self.world
            mine = torch.zeros(chunk, X.shape[1], device=X.device)
            mine[:xhi - xlo] = X[xlo:xhi]
            every = torch.empty(self.world * chunk, X.shape[1], device=X.device)
            dist.all_gather_into_tensor(every, mine, group=self.group)
            X.copy_(every[:n_x])

    def sweep(self):
        self.half_sweep('users')
        self.half_sweep('items')


class ReplicatedTrainer(object):
    """Data-parallel training with REPLICATED tables for models whose tables are small (GBPR on BASELINE configs[2]: 138k
    x 27k x d=64 = 42 MB; sharding it would be all overhead, and GBPR's group users would need a second exchange on U).
    Works for all four models.  Per minibatch every rank
      1. samples its own B pairs (give every rank's sampler a different seed),
      2. runs the fused step in gradient-only mode: every row gradient (and GBPR's bias gradient) is red.added into
         dense tables gU / gV / gb that live in ONE flat buffer,
      3. all-reduces that buffer (NCCL, NVLS in-switch reduction where available),
      4. applies it with cf_apply_dense (rows with an all-zero gradient are skipped; the applied rows are re-zeroed),
    which is the single-GPU minibatch-synchronous step on the global batch of world * B pairs: the same summed gradient
    per row, applied once (TF1 sparse Adagrad semantics, bprmf.py:83-88).  The tables must start identical on every
    rank: ``__init__`` broadcasts rank 0's."""

    def __init__(self, model, sampler, group=None):
        self.torch = torch = _lib.require_cuda()
        self.lib = _lib.lib()
        self.model, self.eng, self.sampler, self.group = model, model.engine, sampler, group
        self.world, self.rank = _world(group)
        eng = self.eng
        if eng.update != 'sync':
            raise ValueError('replicated training needs update="sync"')
        nu, ni, ld = eng.n_users, eng.n_items, eng.ld
        nb = ni if eng.b is not None else 0
        self.flat = torch.zeros((nu + ni) * ld + nb, device=eng.device)
        self.gU = self.flat[:nu * ld].view(nu, ld)
        self.gV = self.flat[nu * ld:(nu + ni) * ld].view(ni, ld)
        self.gb = self.flat[(nu + ni) * ld:] if nb else None
        self.launches = 0
        self.bytes_reduced = 0
        if self.world > 1:
            import torch.distributed as dist
            for t in (eng.U, eng.V, eng.accU, eng.accV, eng.b, eng.accb):
                if t is not None:
                    dist.broadcast(t, 0, group=group)

    def _apply(self, table, acc, grad, ld, d):
        eng = self.eng
        ap = _lib.ApplyArgs()
        ap.table, ap.acc, ap.n_rows, ap.d, ap.ld = _lib.ptr(table), _lib.ptr(acc), int(table.shape[0]), d, ld
        ap.grads, ap.ldg = _lib.ptr(grad), ld
        ap.model, ap.optimizer = eng.model_id, 0 if eng.optimizer == 'adagrad' else 1
        ap.lr, ap.clip_norm = eng.hyper['lr'], eng.hyper['clip_norm']
        _lib.check(self.lib.cf_apply_dense(ap, self.torch.cuda.current_stream(eng.device).cuda_stream), 'cf_apply_dense')

    def step_chunk(self, pairs, negs=None, group=None, ratings=None, want_loss=True):
        """One minibatch on explicit local batches; returns the GLOBAL minibatch loss (CUDA float64 [1]) or None."""
        eng = self.eng
        loss = eng.train_batches(pairs, negs, group, ratings, batch_size=int(pairs.shape[0]), want_loss=want_loss,
                                 grad_tables=(self.gU, self.gV, self.gb))
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat, group=self.group)
            if want_loss:
                dist.all_reduce(loss, group=self.group)
            self.bytes_reduced += self.flat.numel() * 4
        self._apply(eng.U, eng.accU, self.gU, eng.ld, eng.d)
        self._apply(eng.V, eng.accV, self.gV, eng.ld, eng.d)
        if self.gb is not None:
            self._apply(eng.b, eng.accb, self.gb, 1, 1)
        if eng._needs_full_clip:   # step first, then the one-time whole-table clip (cml.py:119-129)
            eng._full_clip(self.torch.cuda.current_stream(eng.device).cuda_stream)
        self.launches += 3 + (1 if self.gb is not None else 0)
        return loss

    def step(self, n_minibatches=1, want_loss=True):
        """Sample + run n minibatches; returns their global losses."""
        torch = self.torch
        B = getattr(self.sampler, 'rows_per_batch', self.sampler.batch_size)     # the rating sampler adds negative rows
        chunk = self.sampler.next_chunk(n_minibatches)
        out = []
        for k in range(n_minibatches):
            part = [t[k * B:(k + 1) * B] for t in chunk]
            kind = self.eng.kind
            if kind == 'gbpr':
                out.append(self.step_chunk(part[0], part[1], group=part[2], want_loss=want_loss))
            elif kind == 'wrmf':
                out.append(self.step_chunk(part[0], ratings=part[1], want_loss=want_loss))
            else:
                out.append(self.step_chunk(part[0], part[1], want_loss=want_loss))
        return torch.cat(out) if want_loss else None
