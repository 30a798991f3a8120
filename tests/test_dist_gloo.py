"""world_size-2 gloo (CPU) test of the multi-GPU exchange plan (collaborativefilteringusingtensorflow_b200/dist.py):
request routing by item % P, row fetch, gradient return -- the host-side logic of SURVEY 8(e); the CUDA kernels plug in
between fetch and push on a GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, ld, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from collaborativefilteringusingtensorflow_b200.dist import ItemExchange, item_shard_rows
        ex = ItemExchange(world, rank)
        n_local = item_shard_rows(n_items, world, rank)
        assert sum(item_shard_rows(n_items, world, r) for r in range(world)) == n_items
        # shard row j holds global item j * world + rank; encode the global id in the row so fetches can be verified
        gid = torch.arange(n_local, dtype=torch.float32) * world + rank
        shard = gid[:, None] * 10 + torch.arange(ld, dtype=torch.float32)[None, :] * 0.001
        g = torch.Generator().manual_seed(100 + rank)
        for trial in range(3):
            ids = torch.randint(0, n_items, (64, 4), generator=g)
            if trial == 2:
                ids = torch.full((5, 2), int(rank), dtype=torch.int64)      # every occurrence the same item
            plan = ex.plan(ids, n_items if trial != 1 else None)      # sort-free dedupe and the torch.unique fallback
            rows = ex.fetch(plan, shard)
            assert rows.shape == (plan.n_req, ld) and plan.n_req == len(torch.unique(ids))
            got = rows[plan.occ_local.to(torch.int64).reshape(-1)]
            want = ids.reshape(-1).to(torch.float32)[:, None] * 10 + torch.arange(ld, dtype=torch.float32)[None, :] * 0.001
            assert torch.allclose(got, want), 'fetched rows do not match the requested global ids'
            # gradient return: rank r sends (1 + r) * id for every row it fetched; owners must see one row per requester
            grads = plan.req_global.to(torch.float32)[:, None].repeat(1, ld) * (1 + rank)
            recv = ex.push(plan, grads)
            assert recv.shape[0] == plan.recv_local_rows.numel()
            owned_global = plan.recv_local_rows.to(torch.float32) * world + rank
            ratio = recv[:, 0] / torch.clamp(owned_global, min=1)
            ok = (owned_global == 0) | torch.isin(ratio, torch.arange(1, world + 1, dtype=torch.float32))
            assert bool(ok.all()), 'a returned gradient row does not belong to the row it is aligned with'
            # per-row sums at the owner == sum over requesters
            summed = torch.zeros(n_local).index_add_(0, plan.recv_local_rows.to(torch.int64), recv[:, 0])
            req_mask = torch.zeros(n_items)
            req_mask[torch.unique(ids)] = 1 + rank
            allm = [torch.zeros(n_items) for _ in range(world)]
            dist.all_gather(allm, req_mask)
            expect_global = sum(allm) * torch.arange(n_items, dtype=torch.float32)
            assert torch.allclose(summed, expect_global[rank::world])
            # owner-pull variant: the whole count matrix is gathered, so the owner knows where its segment starts in every
            # requester's gradient buffer; reading those segments in place (here: from an all-gathered copy standing in
            # for NVLink peer memory) must give exactly what the push would have delivered
            plan2 = ex.plan_exchange(ex.plan_local(ids, n_items), all_counts=True)
            assert plan2.send_counts == plan.send_counts and plan2.recv_counts == plan.recv_counts
            assert torch.equal(plan2.recv_local_rows, plan.recv_local_rows)
            cap = 64 * 4
            mine = torch.zeros(cap, ld)
            mine[:plan.n_req] = grads
            every = [torch.zeros(cap, ld) for _ in range(world)]
            dist.all_gather(every, mine)
            pulled = torch.cat([every[r][plan2.peer_offsets[r]:plan2.peer_offsets[r] + plan2.recv_counts[r]] for r in range(world)])
            assert torch.equal(pulled, recv), 'segments read in place differ from the pushed rows'
        q.put((rank, 'ok'))
    except Exception as e:      # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_item_exchange_over_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1001, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)


def test_single_rank_exchange_is_identity():
    from collaborativefilteringusingtensorflow_b200.dist import ItemExchange
    ex = ItemExchange(1, 0)
    shard = torch.arange(50, dtype=torch.float32)[:, None].repeat(1, 4)
    ids = torch.tensor([[7, 3], [3, 49]])
    plan = ex.plan(ids)
    rows = ex.fetch(plan, shard)
    assert torch.equal(rows[plan.occ_local.to(torch.int64)][..., 0], ids.to(torch.float32))
    assert torch.equal(ex.push(plan, rows), rows)


def _layout_worker(rank, world, port, n_items, ld, q):
    """Host-side layout arithmetic of the N > 1 evaluation / 'replicate' / ALS paths on CPU tensors over gloo.  The product is
    CUDA-only (`_lib.require_cuda()` raises here); these functions only index and call collectives, so the TEST hands them
    torch itself -- nothing of this patch exists outside this process."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import types
        from collaborativefilteringusingtensorflow_b200 import _lib, dist as D
        from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
        _lib.require_cuda = lambda: torch
        n_local = D.item_shard_rows(n_items, world, rank)
        row_of = lambda ids: ids.to(torch.float32)[:, None] * 10 + torch.arange(ld, dtype=torch.float32)[None, :] * 0.001
        shard = row_of(torch.arange(n_local) * world + rank)          # local row j = global item j * world + rank
        # -- the gathered item table of the user-sharded evaluation: item order, also when world does not divide n_items
        full = D._gather_item_rows(shard, n_items, world)
        assert full.shape == (n_items, ld) and torch.equal(full, row_of(torch.arange(n_items)))
        bias = D._gather_item_rows(shard[:, 0].contiguous(), n_items, world)
        assert torch.equal(bias, row_of(torch.arange(n_items))[:, 0])
        # -- the 'replicate' transport's replica: shard after shard, the engine's own table a VIEW of its block
        eng = types.SimpleNamespace(V=shard.clone(), ld=ld, n_items=n_local, device=torch.device('cpu'))
        ex = types.SimpleNamespace(dist=dist, group=None)
        tr = types.SimpleNamespace(torch=torch, eng=eng, world=world, rank=rank, n_items_global=n_items, ex=ex)
        D.DistributedTrainer._setup_replica(tr)
        rep, L = tr._rep['V'], tr._rep['L']
        assert L == (n_items + world - 1) // world and rep.shape == (world * L, ld) and tr._rep['g'].shape == rep.shape
        ids = torch.randperm(n_items, generator=torch.Generator().manual_seed(5))
        assert torch.equal(rep[D.DistributedTrainer._replica_rows(tr, ids)], row_of(ids))
        assert eng.V.shape == (n_local, ld) and eng.V.data_ptr() == rep[rank * L:].data_ptr()      # no copy: a view of the block
        eng.V[0, 0] = -7.0
        assert float(rep[rank * L, 0]) == -7.0
        # every replica row is either one item's row or padding of a short shard (stays zero, never addressed)
        used = torch.zeros(world * L, dtype=torch.bool)
        used[D.DistributedTrainer._replica_rows(tr, torch.arange(n_items))] = True
        assert int(used.sum()) == n_items and bool((rep[~used] == 0).all())
        # -- 'auto' takes the replica exactly when a rank's minibatch holds >= 2 x n_items item occurrences (sizes only: the
        #    same answer on every rank without a collective)
        for B, W, want in ((n_items, 1, True), (n_items - 1, 1, False), (n_items // 3 + 1, 5, True), (1, 1, False)):
            t2 = types.SimpleNamespace(_replicate=None, world=world, n_items_global=n_items, device_side=True)
            assert D.DistributedTrainer._use_replica(t2, B, W) is want and t2.device_side is (not want)
        # -- the mask rows of the item-sharded top-K: owned items only, in local ids
        g = torch.Generator().manual_seed(11)
        dense = torch.rand(7, n_items, generator=g) < 0.05
        r, c = torch.nonzero(dense, as_tuple=True)
        indptr = torch.zeros(8, dtype=torch.int64)
        indptr[1:] = torch.cumsum(dense.sum(1), 0)
        sub = DeviceCSR(indptr, c.to(torch.int32), r.to(torch.int32), None, (7, n_items))
        loc = D.shard_mask_csr(sub, world, rank)
        assert loc.shape == (7, n_local)
        back = torch.zeros(7, n_local, dtype=torch.bool)
        back[loc.rows.to(torch.int64), loc.indices.to(torch.int64)] = True
        assert torch.equal(back, dense[:, rank::world])
        assert torch.equal(loc.indptr[1:] - loc.indptr[:-1], dense[:, rank::world].sum(1))
        # -- ALS row ranges: disjoint, ordered, covering, a rank may be empty
        for n in (1, world - 1, world, 10 * world + 1, n_items):
            rng_ = [D.DistributedALS.row_range(n, world, k) for k in range(world)]
            assert rng_[0][0] == 0 and rng_[-1][1] == n and all(a[1] == b[0] for a, b in zip(rng_, rng_[1:])) and all(lo <= hi for lo, hi in rng_)
        q.put((rank, 'ok'))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n_items', [(2, 1001), (3, 1001), (3, 999), (4, 10)])
def test_sharded_table_layouts_over_gloo(world, n_items):
    """dist._gather_item_rows, DistributedTrainer._setup_replica / _replica_rows / _use_replica, shard_mask_csr and
    DistributedALS.row_range (items `item % P`, users by range) on world_size 2 / 3 / 4."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_layout_worker, args=(r, world, port, n_items, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)


def _metrics_worker(rank, world, port, q):
    """dist.distributed_evaluate (SURVEY 8e "Metrics": per-user partial sums, one all-reduce) over gloo.  The per-user values
    come from the CUDA kernel in the product; here the TEST substitutes the oracle's per-user summands for it, so that what
    runs is the function's own arithmetic: which columns are summed, the user count, mean for CV / sum for LOOV, None for
    unknown names, the exceptions of ranking.py, a rank without users."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import types
        import numpy as np
        from collaborativefilteringusingtensorflow_b200 import _lib, dist as D
        from collaborativefilteringusingtensorflow_b200.metrics import ranking as R
        from oracle import ranking as O
        _lib.require_cuda = lambda: torch

        def per_user(truth, pred, k, loov=False, device=None):
            vals = torch.zeros(len(truth), 8, dtype=torch.float64)
            for t, (y, p) in enumerate(zip(truth, pred.tolist())):
                p = [x for x in p if x >= 0]
                if loov:
                    head = p[:k]
                    vals[t, 5] = float(y in head)
                    vals[t, 6] = 1.0 / (head.index(y) + 1) if y in head else 0.0
                else:
                    vals[t, :5] = torch.tensor(O.per_user_cv(y, p, k) if len(y) else (0.0, 0.0, 0.0, 0.0, 0.0), dtype=torch.float64)
            lens = torch.tensor([1 if loov else len(y) for y in truth])
            return vals, types.SimpleNamespace(row_lengths=lambda: lens)
        R.per_user = per_user
        rng = np.random.default_rng(3)                       # the same global problem on every rank
        n_items, k, T = 60, 7, 23
        truth = [set(rng.choice(n_items, size=int(rng.integers(1, 6)), replace=False).tolist()) for _ in range(T)]
        pred = np.stack([rng.permutation(n_items)[:k] for _ in range(T)]).astype(np.int32)
        owners = world - 1 if world > 2 else world           # world 3: the last rank holds no users
        cut = [T * r // owners for r in range(owners + 1)] + [T] * (world - owners)
        lo, hi = cut[rank], cut[rank + 1]
        names = ['pre', 'recall', 'auc', 'map', 'mrr', 'ndcg']
        got = D.distributed_evaluate(truth[lo:hi], torch.from_numpy(pred[lo:hi]), names, k, 'cv')
        want = O.evaluateCV(truth, [p.tolist() for p in pred], names, k)
        assert got[2] is None and want[2] is None            # unknown metric name -> None (ranking.py:94-109)
        assert np.allclose([g for g in got if g is not None], [w for w in want if w is not None], rtol=1e-12, atol=0), (got, want)
        ys = [int(rng.choice(sorted(t))) for t in truth]
        got = D.distributed_evaluate(ys[lo:hi], torch.from_numpy(pred[lo:hi]), ['hr', 'ndcg', 'arhr'], k, 'loov')
        want = O.evaluateLOOV(ys, [p.tolist() for p in pred], ['hr', 'ndcg', 'arhr'], k)
        assert got[1] is None and np.allclose([got[0], got[2]], [want[0], want[2]], rtol=1e-12)      # sums, not means
        # a user without test items anywhere makes 'map' raise on EVERY rank (ranking.py:53 divides by len(truth))
        truth2 = list(truth)
        truth2[0] = set()
        for m, exc in ((['map'], ZeroDivisionError), (['pre'], None)):
            try:
                D.distributed_evaluate(truth2[lo:hi], torch.from_numpy(pred[lo:hi]), m, k, 'cv')
                raised = None
            except ZeroDivisionError as e:
                raised = type(e)
            assert raised is exc, (m, raised)
        try:                                                 # k <= 0: ranking.py:12-13's ValueError, on every rank, before the collective
            D.distributed_evaluate(truth[lo:hi], torch.from_numpy(pred[lo:hi]), names, 0, 'cv')
            assert False, 'no ValueError'
        except ValueError:
            pass
        q.put((rank, 'ok'))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_metrics_over_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metrics_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)


def _als_worker(rank, world, port, q):
    """dist.DistributedALS (SURVEY 8e "ALS": partial Gram -> all_reduce -> solve the rank's row range -> all_gather) over gloo.
    The two kernels (cf_als_gram, cf_als_solve_rows) are replaced INSIDE THE TEST by dense float64 stand-ins, so what runs is
    the class's own sharding: row ranges, which rows feed the partial Gram, the padded all-gather, both sides of a sweep."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import types
        import numpy as np
        from collaborativefilteringusingtensorflow_b200 import _lib, dist as D
        from oracle import als as oals
        _lib.require_cuda = lambda: torch
        rng = np.random.default_rng(17)                      # the same problem on every rank
        nu, ni, d, weight, reg = 23, 17, 6, 3.0, 0.2
        Rm = rng.random((nu, ni)) < 0.25
        U0 = rng.standard_normal((nu, d)).astype(np.float32)
        V0 = rng.standard_normal((ni, d)).astype(np.float32)
        eng = types.SimpleNamespace(U=torch.from_numpy(U0.copy()), V=torch.from_numpy(V0.copy()), n_users=nu, n_items=ni, d=d,
                                    device=torch.device('cpu'))

        def als_gram(Y, G):
            G[:d, :d] += (Y.double().T @ Y.double()).float()

        def als_solve_rows(X, Y, csr, G):
            Yd, Gd = Y.double().numpy(), G[:d, :d].double().numpy()
            for r, cols in enumerate(csr.cols):
                Yp = Yd[cols]
                A = Gd + (weight - 1.0) * (Yp.T @ Yp) + reg * np.eye(d)
                X[r] = torch.from_numpy(np.linalg.solve(A, weight * Yp.sum(0))).float()
        eng.als_gram, eng.als_solve_rows = als_gram, als_solve_rows

        def local_csr(M, n):
            lo, hi = D.DistributedALS.row_range(n, world, rank)
            return types.SimpleNamespace(shape=(hi - lo, M.shape[1]), cols=[np.flatnonzero(M[r]) for r in range(lo, hi)])
        als = D.DistributedALS(eng, local_csr(Rm, nu), local_csr(Rm.T, ni))
        assert (als.world, als.rank) == (world, rank)
        als.half_sweep('users')
        want_u = oals.half_sweep(V0, [np.flatnonzero(r) for r in Rm], weight, reg)
        assert np.allclose(eng.U.numpy(), want_u, rtol=2e-4, atol=2e-5), np.abs(eng.U.numpy() - want_u).max()
        als.half_sweep('items')                              # uses the users just solved, on every rank the full table
        want_v = oals.half_sweep(eng.U.numpy(), [np.flatnonzero(c) for c in Rm.T], weight, reg)
        assert np.allclose(eng.V.numpy(), want_v, rtol=2e-4, atol=2e-5), np.abs(eng.V.numpy() - want_v).max()
        every = [torch.zeros_like(eng.U) for _ in range(world)]
        dist.all_gather(every, eng.U)
        assert all(torch.equal(every[0], e) for e in every)  # the replicated tables stay identical across ranks
        try:
            D.DistributedALS(eng, types.SimpleNamespace(shape=(nu + 1, ni), cols=[]), None)
            assert False, 'a CSR of the wrong row range was accepted'
        except ValueError:
            pass
        q.put((rank, 'ok'))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_sharded_als_over_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_als_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)


def _topk_worker(rank, world, port, q):
    """dist.distributed_topk (SURVEY 8e "Evaluation": local top-K over the rank's `item % P` shard -> exchange -> merge) over
    gloo, both exchange forms: all-gather + merge of every user on every rank, and the one all-to-all that hands rank r the P
    lists of ITS slice of the users.  The local top-K kernel and cf_topk_merge are replaced inside the test by numpy
    stand-ins that follow the kernels' contract (score desc, item id asc; -1 / -inf padding), so what runs is the function's
    own layout arithmetic: local -> global ids, the stacked gather, the padded slices, lo / n_mine."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import ctypes
        import types
        import numpy as np
        from collaborativefilteringusingtensorflow_b200 import _lib, dist as D
        _lib.require_cuda = lambda: torch
        torch.cuda.current_stream = lambda dev=None: types.SimpleNamespace(cuda_stream=0)

        def view(ptr, n, ct, dt):
            return np.frombuffer((ct * n).from_address(ptr), dtype=dt)

        def cf_topk_merge(pi, pv, P, T, K, po, pov, stream):       # [P, T, K] lists -> [T, K]: score desc, id asc, -1 padded
            ii = view(pi, P * T * K, ctypes.c_int32, np.int32).reshape(P, T, K)
            vv = view(pv, P * T * K, ctypes.c_double, np.float64).reshape(P, T, K)
            oi = view(po, T * K, ctypes.c_int32, np.int32).reshape(T, K)
            ov = view(pov, T * K, ctypes.c_double, np.float64).reshape(T, K)
            for t in range(T):
                ids, vals = ii[:, t].reshape(-1), vv[:, t].reshape(-1)
                keep = ids >= 0
                ids, vals = ids[keep], vals[keep]
                order = np.lexsort((ids, -vals))[:K]
                oi[t], ov[t] = -1, -np.inf
                oi[t, :len(order)], ov[t, :len(order)] = ids[order], vals[order]
            return 0
        _lib.lib = lambda: types.SimpleNamespace(cf_topk_merge=cf_topk_merge)
        rng = np.random.default_rng(23)                      # the same problem on every rank
        n_items, d, K = 41, 5, 6
        for T in (11, 2):                                    # 2 users over 3 ranks: a rank's slice is empty
            Vg = rng.standard_normal((n_items, d))
            Q = rng.standard_normal((T, d))
            Vg[7] = Vg[12]                                   # an exact tie across two shards: the lower item id first
            shard = torch.from_numpy(Vg[rank::world].copy())
            eng = types.SimpleNamespace(U=None, n_users=0, V=shard)

            def topk(users, k, csr, return_values=True, method='auto', eng=eng):
                s = eng.U.numpy() @ eng.V.numpy().T          # [T, local items], local ids
                idx = np.full((s.shape[0], k), -1, np.int32)
                val = np.full((s.shape[0], k), -np.inf)
                for t in range(s.shape[0]):
                    order = np.lexsort((np.arange(s.shape[1]), -s[t]))[:k]
                    idx[t, :len(order)], val[t, :len(order)] = order, s[t, order]
                return torch.from_numpy(idx), torch.from_numpy(val)
            eng.topk = topk
            full = Q @ Vg.T
            want = np.stack([np.lexsort((np.arange(n_items), -full[t]))[:K] for t in range(T)]).astype(np.int32)
            gi, gv = D.distributed_topk(eng, torch.from_numpy(Q), K, None, world, rank)
            assert eng.U is None and eng.n_users == 0        # the engine's own user table is put back
            assert np.array_equal(gi.numpy(), want), (gi, want)
            assert np.allclose(gv.numpy(), np.take_along_axis(full, want.astype(np.int64), 1), rtol=0, atol=1e-12)
            lo, si, sv = D.distributed_topk(eng, torch.from_numpy(Q), K, None, world, rank, gather=False)
            c = (T + world - 1) // world
            assert lo == rank * c and si.shape[0] == max(0, min(T, lo + c) - lo)
            assert np.array_equal(si.numpy(), want[lo:lo + c]) and np.allclose(sv.numpy(), gv.numpy()[lo:lo + c], rtol=0, atol=0)
        q.put((rank, 'ok'))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_item_sharded_topk_exchange_over_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_topk_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)
