#!/usr/bin/env python
"""Multi-GPU parity script (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Sharded training (users range-sharded, items by item % P, all-to-all of rows and gradients) must equal ONE GPU running the
same global minibatch (P * B pairs) with the plain fused step; sharded top-K + all-gather merge must equal single-GPU
top-K.  Rank 0 prints 'DIST_CHECK OK'."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stage_collectives_through_host():
    """ONE-GPU form of this script (CF_DIST_BACKEND=gloo): NCCL refuses two ranks on one device, so the ranks -- separate
    processes that share GPU 0 -- talk over gloo, and every collective on a CUDA tensor is staged through the host here:
    device synchronise, copy out, the CPU collective, copy back.  That keeps what the product relies on (a collective
    completes only after everything every rank enqueued before it) and changes nothing else: the device-side exchange
    still runs over CUDA-IPC mappings of the other PROCESS's buffers, the kernels and the plan are the ones of N GPUs."""
    real = {n: getattr(dist, n) for n in ('all_reduce', 'broadcast', 'all_gather', 'all_gather_into_tensor',
                                          'all_to_all_single', 'reduce_scatter_tensor')}

    def out_in(name, n_out, flat=False):
        def f(*a, **kw):
            a = list(a)
            cu = [k for k in range(len(a)) if torch.is_tensor(a[k]) and a[k].is_cuda]
            if not cu:
                return real[name](*a, **kw)
            torch.cuda.synchronize()
            dev_t = {k: a[k] for k in cu}
            fl = flat or (name == 'all_to_all_single' and len(a) == 2 and not kw.get('output_split_sizes') and not kw.get('input_split_sizes'))
            for k in cu:
                a[k] = dev_t[k].cpu().reshape(-1) if fl else dev_t[k].cpu()   # gloo checks shapes, NCCL only sizes
            real[name](*a, **kw)
            for k in cu[:n_out]:
                dev_t[k].copy_(a[k].view(dev_t[k].shape))
            torch.cuda.synchronize()
        return f
    dist.all_reduce = out_in('all_reduce', 1)
    dist.broadcast = out_in('broadcast', 1)
    dist.all_gather_into_tensor = out_in('all_gather_into_tensor', 1, flat=True)
    dist.all_to_all_single = out_in('all_to_all_single', 1)
    dist.reduce_scatter_tensor = out_in('reduce_scatter_tensor', 1, flat=True)

    def all_gather(outs, t, group=None):
        if not t.is_cuda:
            return real['all_gather'](outs, t, group=group)
        torch.cuda.synchronize()
        host = [torch.empty(o.shape, dtype=o.dtype) for o in outs]
        real['all_gather'](host, t.cpu(), group=group)
        for o, h in zip(outs, host):
            o.copy_(h)
        torch.cuda.synchronize()
    dist.all_gather = all_gather


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    backend = os.environ.get('CF_DIST_BACKEND', 'nccl')
    if backend == 'nccl':
        torch.cuda.set_device(local)
        dev = torch.device('cuda', local)
        dist.init_process_group('nccl', device_id=dev)
    else:
        local = local % torch.cuda.device_count()               # the ranks share the GPUs there are (one: all on cuda:0)
        torch.cuda.set_device(local)
        dev = torch.device('cuda', local)
        dist.init_process_group('gloo')
        stage_collectives_through_host()
        os.environ['CF_DIST_HOST_BARRIER'] = '1'                # dist.py: the named barrier = device synchronise + host barrier
        if rank == 0:
            print('backend gloo: %d ranks on %d GPU(s), collectives staged through the host' % (world, torch.cuda.device_count()))
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML
    from collaborativefilteringusingtensorflow_b200.dist import DistributedTrainer, distributed_topk, item_shard_rows
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    from scipy.sparse import lil_matrix
    ok = True
    transport = sys.argv[1] if len(sys.argv) > 1 else 'nccl'    # 'nccl' | 'peer' | 'fetch' | 'auto'
    for kind in ('bpr', 'cml'):
        nu_l, ni, d, B, W = 500, 1203, 128, 1024, 3
        nu = nu_l * world
        rng = np.random.default_rng(7)                                  # same stream on every rank
        U0 = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)    # row norms ~1.13 > clip_norm: the first step sees the
        V0 = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)    # unclipped init, then both whole tables are clipped (cml.py:119-129)
        mk = (lambda n_u, n_i: BPRMF(n_u, n_i, n_factors=d, reg=0.05, verbose=False, seed=1, device=dev)) if kind == 'bpr' else \
             (lambda n_u, n_i: CML(n_u, n_i, n_factors=d, reg_cov=1.0, margin=1.0, verbose=False, seed=1, device=dev))
        local_m = mk(nu_l, item_shard_rows(ni, world, rank))
        local_m.load_state_dict(dict(U=U0[rank * nu_l:(rank + 1) * nu_l], V=V0[rank::world]))

        class NoSampler(object):
            batch_size = B
        tr = DistributedTrainer(local_m, NoSampler(), ni, world, rank, item_transport=transport)
        ref = None
        if rank == 0:
            ref = mk(nu, ni)
            ref.load_state_dict(dict(U=U0, V=V0))
        for step in range(3):
            pairs = np.stack([rng.integers(0, nu_l, (world, B)), rng.integers(0, ni, (world, B))], 2).astype(np.int32)
            negs = rng.integers(0, ni, (world, B, W)).astype(np.int32)
            tr.step_chunk(torch.from_numpy(pairs[rank]).to(dev), torch.from_numpy(negs[rank]).to(dev), B)
            local_m.engine.check_flags()
            if rank == 0:
                gp = pairs.copy()
                gp[:, :, 0] += (np.arange(world) * nu_l)[:, None]        # local -> global user ids
                ref.step(gp.reshape(-1, 2), negs.reshape(-1, W))
        # gather the shards on rank 0 and compare
        st = local_m.state_dict()
        Us = [torch.empty_like(st['U']) for _ in range(world)]
        dist.all_gather(Us, st['U'].contiguous())
        n_max = item_shard_rows(ni, world, 0)
        Vpad = torch.zeros(n_max, d, device=dev)
        Vpad[:st['V'].shape[0]] = st['V']
        Vs = [torch.empty_like(Vpad) for _ in range(world)]
        dist.all_gather(Vs, Vpad)
        if rank == 0:
            r = ref.state_dict()
            Ug = torch.cat(Us).cpu().numpy()
            Vg = np.zeros((ni, d), np.float32)
            for p in range(world):
                Vg[p::world] = Vs[p][:item_shard_rows(ni, world, p)].cpu().numpy()
            for name, got, want in (('U', Ug, r['U'].cpu().numpy()), ('V', Vg, r['V'].cpu().numpy())):
                err = np.abs(got - want).max()
                # fp32 sums of a row's gradients are regrouped per rank: the more ranks, the more duplicates per row in the
                # global minibatch; CML's rank-weighted coefficients (~10) amplify that
                good = np.allclose(got, want, rtol=5e-5, atol=(2e-6 if kind == 'bpr' else 1e-5) * max(1, world // 2))
                print('%s %s max|diff| = %.3g %s' % (kind, name, err, 'ok' if good else 'MISMATCH'))
                ok &= bool(good)
        # ---- evaluation: item-sharded top-K + all-gather merge vs single GPU
        T, K = 64, 100
        users = np.arange(T)
        tra = lil_matrix((T, ni), dtype=np.float32)
        for t in range(T):
            tra[t, rng.choice(ni, 30, replace=False)] = 1
        # query embeddings live on their owner rank: all-gather the tile (here: users of rank 0's shard, replicated)
        q = local_m.engine.U[:T].clone() if rank == 0 else torch.empty(T, local_m.engine.ld, device=dev)
        dist.broadcast(q, 0)
        local_mask = lil_matrix((T, item_shard_rows(ni, world, rank)), dtype=np.float32)
        for t in range(T):
            cols = [c // world for c in tra.rows[t] if c % world == rank]
            if cols:
                local_mask[t, cols] = 1
        gi, gv = distributed_topk(local_m.engine, q, K, DeviceCSR.from_scipy(local_mask, dev), world, rank)
        if rank == 0:
            # reference: a single engine holding the gathered tables
            full = mk(nu_l, ni)
            full.load_state_dict(dict(U=local_m.state_dict()['U'], V=torch.from_numpy(Vg)))
            wi, wv = full.engine.topk(torch.arange(T), K, DeviceCSR.from_scipy(tra, dev), return_values=True)
            same = bool(torch.equal(wi, gi) and torch.equal(wv, gv))
            print('%s sharded top-%d == single GPU: %s' % (kind, K, same))
            ok &= same
        # the all-to-all form: every rank merges only its slice of the users; the slices together are the same lists
        lo, si, sv = distributed_topk(local_m.engine, q, K, DeviceCSR.from_scipy(local_mask, dev), world, rank, gather=False)
        same = bool(torch.equal(si, gi[lo:lo + si.shape[0]]) and torch.equal(sv, gv[lo:lo + si.shape[0]]))
        cover = torch.tensor([si.shape[0]], device=dev)
        dist.all_reduce(cover)
        flag = torch.tensor([int(same and int(cover.item()) == T)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print('%s sharded top-%d, sliced merge (all-to-all) == all-gather merge on every rank: %s' % (kind, K, bool(flag.item())))
            ok &= bool(flag.item())
        # the user-sharded form: the item table is all-gathered, every rank scores ITS slice of the users against the
        # whole catalogue (global item ids in the mask); the slices together are the single-GPU lists
        from collaborativefilteringusingtensorflow_b200.dist import user_sharded_topk
        c_u = (T + world - 1) // world
        ulo, uhi = min(T, rank * c_u), min(T, (rank + 1) * c_u)
        saved_U, saved_nu = local_m.engine.U, local_m.engine.n_users
        local_m.engine.U, local_m.engine.n_users = q.contiguous(), T       # (the query tile stands in for the rank's own users)
        try:
            if uhi > ulo:
                full_rows = lil_matrix((T, ni), dtype=np.float32)           # the mask is indexed by user id
                for t in range(ulo, uhi):
                    full_rows[t, tra.rows[t]] = 1
                ui, uv = user_sharded_topk(local_m.engine, torch.arange(ulo, uhi, dtype=torch.int32, device=dev), K,
                                           DeviceCSR.from_scipy(full_rows, dev), ni, world, rank)
                same = bool(torch.equal(ui, gi[ulo:uhi]) and torch.equal(uv, gv[ulo:uhi]))
            else:
                from collaborativefilteringusingtensorflow_b200.dist import gather_item_table
                gather_item_table(local_m.engine, ni, world, rank)          # (collective: every rank takes part)
                same = True
        finally:
            local_m.engine.U, local_m.engine.n_users = saved_U, saved_nu
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print('%s user-sharded top-%d (gathered item table) == all-gather merge on every rank: %s' % (kind, K, bool(flag.item())))
            ok &= bool(flag.item())
    # ---- the trainer's own step(): sampler-driven minibatches.  With the 'replicate' transport step() is a three-stream
    # pipeline (sampling of k + 1 and the all-gather of k beside the compute stream); it must give what the same minibatches
    # give one after the other on one stream (CF_REPLICA_OVERLAP=0), up to the order of the atomic adds
    if transport in ('replicate', 'auto'):
        from collaborativefilteringusingtensorflow_b200.samplers.sampler_ranking import Sampler
        nu_l, ni, d, B, W = 500, 1203, 128, 1024, 3
        rs = np.random.default_rng(100 + rank)
        R_local = lil_matrix((nu_l, ni), dtype=np.float32)
        cells = rs.choice(nu_l * ni, 9 * B + 17, replace=False)
        R_local[cells // ni, cells % ni] = 1
        results = []
        for overlap in ('1', '0', '0'):
            os.environ['CF_REPLICA_OVERLAP'] = overlap
            mm = CML(nu_l, item_shard_rows(ni, world, rank), n_factors=d, reg_cov=1.0, margin=1.0, verbose=False, seed=50 + rank, device=dev)
            sm = Sampler(R_local, n_neg=W, batch_size=B, seed=9 + rank)
            trs = DistributedTrainer(mm, sm, ni, world, rank, item_transport='replicate')
            losses = torch.cat([trs.step(4), trs.step(1), trs.step(3)])          # 8 minibatches, pipelined inside each call
            mm.engine.check_flags()
            results.append((mm.state_dict(), losses.clone()))
            trs.close()
        os.environ.pop('CF_REPLICA_OVERLAP')
        # Eight CML minibatches apart, two runs of the SAME schedule already differ by the order of their float atomics
        # (item-row gradients red.added into the dense table; rank-weighted coefficients ~10 with heavy cancellation).  That
        # noise is measured -- the one-stream form run twice -- and the pipelined run must be no farther from a one-stream
        # run than 10x that (+ a floor of 1e-5 relative; N = 2 on B200: ratios 0.3 .. 2.6): a stale or half-gathered item row would be
        # off by 1e-2 and more.
        (a_, la), (b_, lb), (c_, lc) = results
        good, report = True, []
        for k in ('U', 'V', 'accU', 'accV'):
            noise = float(((b_[k] - c_[k]).abs() / (c_[k].abs() + 1e-2)).max())
            diff = float(((a_[k] - b_[k]).abs() / (b_[k].abs() + 1e-2)).max())
            report.append('%s %.2e (noise %.2e)' % (k, diff, noise))
            good &= diff <= 10 * noise + 1e-5 and diff < 1e-2
        nl, dl = float(((lb - lc).abs() / lc.abs()).max()), float(((la - lb).abs() / lb.abs()).max())
        report.append('loss %.2e (noise %.2e)' % (dl, nl))
        good &= dl <= 10 * nl + 1e-8 and dl < 1e-5
        print('rank %d pipelined step() vs one stream, max relative distance: %s' % (rank, ', '.join(report)))
        flag = torch.tensor([int(good)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print('replicate: pipelined step() (sampling + all-gather beside the compute stream) == one stream on every rank: %s'
                  % bool(flag.item()))
            ok &= bool(flag.item())
    # ---- metrics: users sharded, sums all-reduced (every rank gets the global values)
    from collaborativefilteringusingtensorflow_b200.dist import DistributedALS, distributed_evaluate
    from collaborativefilteringusingtensorflow_b200.metrics.ranking import evaluateCV
    from collaborativefilteringusingtensorflow_b200 import WRMF
    rng = np.random.default_rng(11)
    Tm, nm, km = 101, 300, 10                                            # 101 users: uneven shards
    truth = [set(rng.choice(nm, int(rng.integers(1, 9)), replace=False).tolist()) for _ in range(Tm)]
    pred = [rng.choice(nm, 15, replace=False).tolist() for _ in range(Tm)]
    lo, hi = DistributedALS.row_range(Tm, world, rank)
    names = ['pre', 'recall', 'ndcg', 'map', 'mrr']
    got = distributed_evaluate(truth[lo:hi], pred[lo:hi], names, km)
    want = evaluateCV(truth, pred, names, km)
    same = bool(np.allclose(got, want, rtol=0, atol=1e-12))
    if not same:
        print('rank %d metrics %r vs %r' % (rank, got, want))
    flag = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('sharded metrics == evaluateCV on every rank: %s' % bool(flag.item()))
        ok &= bool(flag.item())
    # ---- WRMF ALS: rows range-sharded, Gram all-reduced, solved rows all-gathered, vs the single-GPU sweeps
    nu_a, ni_a, da = 301, 515, 64
    Ra = lil_matrix((nu_a, ni_a), dtype=np.float32)
    for u in range(nu_a):
        Ra[u, rng.choice(ni_a, int(rng.integers(1, 30)), replace=False)] = 1
    one = WRMF(nu_a, ni_a, weight=4.0, reg=0.3, n_factors=da, verbose=False, seed=2, solver='als', device=dev)
    many = WRMF(nu_a, ni_a, weight=4.0, reg=0.3, n_factors=da, verbose=False, seed=2, solver='als', device=dev)
    many.load_state_dict(one.state_dict())
    csr = DeviceCSR.from_scipy(Ra, dev)
    csr_t = csr.transpose()
    ulo, uhi = DistributedALS.row_range(nu_a, world, rank)
    ilo, ihi = DistributedALS.row_range(ni_a, world, rank)
    als = DistributedALS(many.engine, csr.select_rows(torch.arange(ulo, uhi, device=dev)),
                         csr_t.select_rows(torch.arange(ilo, ihi, device=dev)))
    for sweep in range(2):
        one.engine.als_half_sweep('users', csr)
        one.engine.als_half_sweep('items', csr_t)
        als.sweep()
    a, b = one.state_dict(), many.state_dict()
    good = all(bool(torch.allclose(a[k], b[k], rtol=5e-4, atol=1e-5)) for k in ('U', 'V'))
    if not good:
        print('rank %d ALS diffs %s' % (rank, {k: float((a[k] - b[k]).abs().max()) for k in ('U', 'V')}))
    flag = torch.tensor([int(good)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('sharded ALS sweeps == single GPU on every rank: %s (max|dU| %.3g)' % (bool(flag.item()), float((a['U'] - b['U']).abs().max())))
        ok &= bool(flag.item())
    # ---- GBPR: replicated tables, dense gradients all-reduced, vs ONE GPU on the global minibatch
    from collaborativefilteringusingtensorflow_b200 import GBPRMF
    from collaborativefilteringusingtensorflow_b200.dist import ReplicatedTrainer
    nu_g, ni_g, dg, Bg, Wg, Gg = 400, 300, 64, 1024, 5, 3
    mkg = lambda s: GBPRMF(nu_g, ni_g, rho=0.4, gsize=Gg, reg=0.01, n_factors=dg, verbose=False, seed=s, device=dev)
    rep = mkg(100 + rank)                                                # different init per rank: the trainer broadcasts rank 0's

    class NoSamplerG(object):
        batch_size = Bg
    trg = ReplicatedTrainer(rep, NoSamplerG())
    refg = mkg(100)
    rng = np.random.default_rng(13)
    for step in range(3):
        pairs = np.stack([rng.integers(0, nu_g, (world, Bg)), rng.integers(0, ni_g, (world, Bg))], 2).astype(np.int32)
        negs = rng.integers(0, ni_g, (world, Bg, Wg)).astype(np.int32)
        grp = rng.integers(0, nu_g, (world, Bg, Gg)).astype(np.int32)
        lg = trg.step_chunk(torch.from_numpy(pairs[rank]).to(dev), torch.from_numpy(negs[rank]).to(dev), group=torch.from_numpy(grp[rank]).to(dev))
        lr_ = refg.step(pairs.reshape(-1, 2), negs.reshape(-1, Wg), grp.reshape(-1, Gg))
    a, b = refg.state_dict(), rep.state_dict()
    # Is the distance between the replicated result and a single GPU just fp32 regrouping noise (atomics order on one GPU
    # vs per-rank atomics + NCCL reduction order)?  Measure it: the numpy oracle sums every row's gradient terms in FP64
    # and rounds once (oracle/steps.py: _segment_sum), i.e. it is the "fp64 all-reduce" of the same minibatches.  Every rank
    # also ran the SAME global minibatches on ONE GPU (refg): `world` independent samples of the single-GPU atomics noise.
    # The replicated result must be no farther from the fp64-summed oracle than the single-GPU runs are (factor 4: the
    # maximum of a handful of samples), in absolute and in relative terms, for parameters and Adagrad accumulators alike.
    from oracle import steps as osteps
    o = {k: v.cpu().numpy().copy() for k, v in mkg(100).state_dict().items()}
    rng = np.random.default_rng(13)
    for step in range(3):
        pairs = np.stack([rng.integers(0, nu_g, (world, Bg)), rng.integers(0, ni_g, (world, Bg))], 2).astype(np.int32)
        negs = rng.integers(0, ni_g, (world, Bg, Wg)).astype(np.int32)
        grp = rng.integers(0, nu_g, (world, Bg, Gg)).astype(np.int32)
        osteps.gbpr_step(o['U'], o['V'], o['b'], o['accU'], o['accV'], o['accb'], pairs.reshape(-1, 2), negs.reshape(-1, Wg),
                         grp.reshape(-1, Gg), 0.1, 0.01, 0.4)

    def dist_to_oracle(state):
        out = []
        for k in sorted(o):
            x, w = state[k].cpu().numpy().astype(np.float64), o[k].astype(np.float64)
            out += [np.abs(x - w).max(), (np.abs(x - w) / (np.abs(w) + 1e-3)).max()]
        return out
    e_single = torch.tensor(dist_to_oracle(a), dtype=torch.float64, device=dev)
    e_multi = torch.tensor(dist_to_oracle(b), dtype=torch.float64, device=dev)
    dist.all_reduce(e_single, op=dist.ReduceOp.MAX)       # farthest of the `world` single-GPU runs
    dist.all_reduce(e_multi, op=dist.ReduceOp.MAX)        # (the replicated tables are identical on every rank)
    floor = torch.tensor([1e-6, 1e-5] * len(o), dtype=torch.float64, device=dev)
    sane = torch.tensor([5e-2, 2e-3] * len(o), dtype=torch.float64, device=dev)   # and neither is far from the oracle at all
    good = bool((e_multi <= 4 * e_single + floor).all()) and bool((e_multi <= sane).all()) and bool((e_single <= sane).all()) \
        and abs(float(lg.item()) - lr_) < 2e-5 * abs(lr_)
    if rank == 0:
        names = [k + s for k in sorted(o) for s in (' abs', ' rel')]
        print('GBPR distance to the fp64-summed oracle, max over elements (single GPU: farthest of %d runs | replicated):' % world)
        for n_, x, y in zip(names, e_single.tolist(), e_multi.tolist()):
            print('  %-9s %.3g | %.3g' % (n_, x, y))
    flag = torch.tensor([int(good)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('replicated GBPR (all-reduced dense gradients) == single GPU on every rank: %s' % bool(flag.item()))
        ok &= bool(flag.item())
    if rank == 0:
        print('DIST_CHECK OK' if ok else 'DIST_CHECK FAILED')
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
