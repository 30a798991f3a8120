"""bench.py's N > 1 arm: weak scaling of the N=1 workload (same users, interactions and minibatch PER GPU; the item
catalogue keeps its global size and is row-sharded, like BASELINE.json configs[4] keeps 10 M items for 8 x 12.5 M users):
users range-sharded, items sharded by item % N, per-minibatch exchange of requested item rows and their gradients
(SURVEY.md 8e).  `--grow-catalogue` instead multiplies the catalogue by N (every rank owns n_items rows).
Every N > 1 line also carries `c5`: BASELINE.json configs[4] at this N (BPRMF, 12.5 M users and 625 M interactions per
GPU, 10 M items row-sharded: the sharded step and the item-sharded 10 M-item top-100)."""
import json
import sys
import time

import torch
import torch.distributed as dist


def _max_over_ranks(x, device):
    t = torch.tensor([float(x)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_training(B_, wl, args, rank, world, device, n_items_global, K, Wm, want_e2e=True):
    """Builds this rank's share of the workload and times K sharded minibatches (barrier + synchronize on both sides, CUDA
    events, max over ranks).  Returns (model, csr, trainer, dict of numbers)."""
    from collaborativefilteringusingtensorflow_b200.dist import DistributedTrainer, item_shard_rows
    B = args.batch
    csr = B_.synth_interactions(wl['n_users'], n_items_global, wl['nnz'], B_.SEED + rank, device)
    local_wl = dict(wl, n_items=item_shard_rows(n_items_global, world, rank))
    model = B_.make_model(local_wl, device, seed=B_.SEED + rank, optimizer=args.optimizer, update='sync')
    sampler = B_.make_sampler(wl, csr, B, B_.SEED + rank, device)
    tr = DistributedTrainer(model, sampler, n_items_global, world, rank, item_transport=args.item_transport)
    tr.step(Wm)
    model.engine.check_flags()
    warm = sampler.next_chunk(K)      # allocator blocks of the timed region's K-minibatch index buffer (see bench.py)
    del warm
    torch.cuda.synchronize()
    dist.barrier()
    l0, s0, b0, p0 = tr.launches, sampler.launches, tr.bytes_sent, tr.bytes_pulled
    q0 = int(tr.req_rows_dev.item()) if tr.req_rows_dev is not None else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    if args.phases:
        tr.step_events = []
    losses = tr.step(K)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.time()
    ms = _max_over_ranks(e0.elapsed_time(e1), device)
    launches = (tr.launches - l0) + (sampler.launches - s0)
    if args.phases and rank == 0:
        evs, tr.step_events = [e0] + tr.step_events, None
        print('per-minibatch ms: ' + ' '.join('%.2f' % evs[k].elapsed_time(evs[k + 1]) for k in range(len(evs) - 1)), file=sys.stderr)
    tr.step_events = None
    sent = (tr.bytes_sent - b0) / K
    remote = (world - 1) / world                                  # the share of the item rows that lives on other GPUs
    pulled = (tr.bytes_pulled - p0) / K * remote                  # pull transport: one row per occurrence, read inside k_step
    if tr.req_rows_dev is not None:
        # device-side exchange: the unique rows this rank requested (exact, counted on the device).  Each is fetched once
        # (fetch transport) and its gradient row is read once by its owner; by symmetry an owner reads as many remote
        # gradient rows as a requester has remote unique rows (items hash uniformly over the owners)
        uniq = (int(tr.req_rows_dev.item()) - q0) / K
        rowb = model.engine.ld * 4
        # (pull transport: rows in per occurrence -- counted above -- and gradient rows OUT per occurrence, the other direction)
        pulled += (tr.bytes_pulled - p0) / K * remote if tr._push else (1 if tr._pull else 2) * uniq * rowb * remote
    model.engine.check_flags()
    out = dict(ms=ms, t0=t0, t1=t1, launches=launches, losses=losses, sent=sent, pulled=pulled, e2e_ms=None, h2d=None)
    if want_e2e:
        # e2e: pinned host index buffers -> H2D (copy stream, double-buffered) -> on-device sampling launch -> sharded step ->
        # D2H loss, every step (bench.time_e2e, the N = 1 loop with the sharded step plugged in)
        ms2, h2d, _ = B_.time_e2e(None, sampler, B, K, device, step=lambda bufs: tr.step_chunk(bufs[0], bufs[1], B), pre=dist.barrier)
        dist.barrier()
        out['e2e_ms'] = _max_over_ranks(ms2, device)
        out['h2d'] = h2d
    # per-phase CUDA-event times (synchronises after every phase, so the phases do not overlap: their sum exceeds a step)
    tr.phase_ms = {}
    tr.step(5)
    ph = {k: v / 5 for k, v in tr.phase_ms.items()}
    tr.phase_ms = None
    keys = sorted(ph)
    vals = torch.tensor([ph[k] for k in keys], device=device, dtype=torch.float64)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    out['phases'] = dict(zip(keys, [float(v) for v in vals.tolist()]))
    return model, csr, tr, out


def sharded_topk(B_, model, csr, n_items_global, wl, Tq, rank, world, device, pk):
    """Item-sharded full-catalogue top-100 of rank 0's first Tq users, merged across ranks."""
    from collaborativefilteringusingtensorflow_b200.dist import distributed_topk, shard_mask_csr
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    eng = model.engine
    # the query users live on rank 0: broadcast their embeddings and their training rows (global item ids)
    q = eng.U[:Tq].clone() if rank == 0 else torch.empty(Tq, eng.ld, device=device)
    dist.broadcast(q, 0)
    sub = csr.select_rows(torch.arange(Tq, device=device)) if rank == 0 else None
    meta = torch.tensor([sub.nnz if rank == 0 else 0], device=device)
    dist.broadcast(meta, 0)
    nnz = int(meta.item())
    ind = sub.indices if rank == 0 else torch.empty(nnz, dtype=torch.int32, device=device)
    rws = sub.rows if rank == 0 else torch.empty(nnz, dtype=torch.int32, device=device)
    ptr = sub.indptr if rank == 0 else torch.empty(Tq + 1, dtype=torch.int64, device=device)
    for t in (ind, rws, ptr):
        dist.broadcast(t, 0)
    mask = shard_mask_csr(DeviceCSR(ptr, ind, rws, None, (Tq, n_items_global)), world, rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    distributed_topk(eng, q, 100, mask, world, rank, method='tensor', gather=False)           # warm-up: sizes every workspace
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    lo, gi, gv = distributed_topk(eng, q, 100, mask, world, rank, method='tensor', gather=False)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    tms = _max_over_ranks(e0.elapsed_time(e1), device)
    fl = 2.0 * n_items_global * wl['d'] * Tq
    tf = fl / (tms * 1e-3) / 1e12
    return dict(metric='users/s full-catalog top-100 (mask train items), items sharded over %d GPUs, local lists exchanged with one '
                       'all-to-all, every rank merges 1/N of the users' % world,
                value=Tq / (tms * 1e-3), unit='users/s', users=Tq, n_items=n_items_global, ms=tms, tflops_aggregate=tf,
                frac_of_tensor_peak_per_gpu=tf / world / pk['bf16'], fallback_rows_rank0=int(eng.tc_stats[0].item()))


def user_sharded_topk_bench(B_, model, csr, n_items_global, wl, Tq, rank, world, device, pk):
    """Full-catalogue top-100 of Tq users in total, Tq / N of them on every rank (each rank's OWN first users, with their own
    training rows): the item table is all-gathered inside the timed region, then every rank scores its users against the
    whole catalogue.  value = Tq / max-over-ranks time."""
    from collaborativefilteringusingtensorflow_b200.dist import user_sharded_topk
    eng = model.engine
    Tl = min((Tq + world - 1) // world, eng.n_users)
    users = torch.arange(Tl, dtype=torch.int32, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    user_sharded_topk(eng, users[:4096], 100, csr, n_items_global, world, rank, method='tensor', return_values=False)
    user_sharded_topk(eng, users, 100, csr, n_items_global, world, rank, method='tensor', return_values=False)   # sizes the workspace
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    user_sharded_topk(eng, users, 100, csr, n_items_global, world, rank, method='tensor', return_values=False)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    tms = _max_over_ranks(e0.elapsed_time(e1), device)
    fl = 2.0 * n_items_global * wl['d'] * Tl * world
    tf = fl / (tms * 1e-3) / 1e12
    return dict(metric='users/s full-catalog top-100 (mask train items), users sharded over %d GPUs: the row-sharded item table is '
                       'all-gathered (inside the timed region) and every rank scores its own users against the whole catalogue' % world,
                value=Tl * world / (tms * 1e-3), unit='users/s', users=Tl * world, n_items=n_items_global, ms=tms, tflops_aggregate=tf,
                frac_of_tensor_peak_per_gpu=tf / world / pk['bf16'], fallback_rows_rank0=int(eng.tc_stats[0].item()))


def step_line(B_, wl, args, world, K, Wm, r, n_items_global, tr, pk):
    B = args.batch
    ms = r['ms']
    units = world * B * wl['W'] * K
    bpp = B_.bytes_per_pair(wl['model'], wl['d'], wl['W'], wl['G'], args.optimizer)
    achieved = bpp * B / (ms / K * 1e-3) / 1e9            # per GPU, whole sharded step (exchange included)
    nv_bytes = r['sent'] + r['pulled']
    return dict(value=units / (ms * 1e-3), unit='triple updates/s', ms_per_step=ms / K, steps=K, gpu_launches=r['launches'],
                item_transport=('item table replicated on every GPU (users stay sharded): item-row gradients red.added into a dense table by '
                                'k_step, one NCCL reduce-scatter, cf_apply_dense on the owner\'s shard, one all-gather of the updated rows'
                                if getattr(tr, '_replicate', False) else
                                ('device-side exchange over peer memory: ' +
                                 ('item rows read per occurrence inside k_step and gradients red.added into the owners\' dense tables by the same kernel (pull + push)'
                                  if tr._push else 'item rows read per occurrence inside k_step (pull) + gradient rows read in place by the owners'
                                  if tr._pull else 'unique item rows gathered once by k_exchange_prepare (fetch) + gradient rows read in place by the owners'))
                                if tr.device_side else
                                'nccl (all-to-all of ids, unique rows and gradient rows)'),
                roofline=dict(bound='hbm', kernel='whole sharded step per GPU (plan + k_count + k_step + k_apply_staged + exchange + owner apply)',
                              achieved=achieved, peak=pk['hbm'], unit='GB/s', frac=achieved / pk['hbm'], traffic=None,
                              peak_source=pk['source']),
                nvlink=dict(bytes_per_step_per_gpu=nv_bytes, bytes_sent_nccl=r['sent'], bytes_read_peer=r['pulled'],
                            achieved_GBs=nv_bytes / (ms / K * 1e-3) / 1e9, peak_GBs_per_direction=770.0,
                            frac=nv_bytes / (ms / K * 1e-3) / 1e9 / 770.0,
                            note='item rows in + gradient rows out + ids, per GPU per minibatch, over the whole step time; '
                                 'measured peer copy 770 GB/s per direction'),
                phases_ms_per_step=r['phases'],
                loss_first_last=[float(r['losses'][0]), float(r['losses'][-1])])


def sharded_als(B_, cfg, rank, world, device, pk):
    """BASELINE.json configs[3] at this N (ALS): one USER half-sweep of WRMF weighted ALS with the users
    range-sharded -- every rank holds cfg['n_users'] users and cfg['nnz'] interactions of its own (weak scaling, as in the
    training line), both factor tables replicated.  Per half-sweep: partial tcgen05 Gram of the rank's slice of the item
    table -> all-reduce of the 128 x 128 Gram -> cf_als_solve_rows on the rank's users -> all-gather of the solved rows.
    Device-timed between barriers, max over ranks."""
    from collaborativefilteringusingtensorflow_b200 import WRMF
    from collaborativefilteringusingtensorflow_b200.dist import DistributedALS
    nu, ni, nnz, d = cfg['n_users'], cfg['n_items'], cfg['nnz'], cfg['d']
    csr = B_.synth_interactions(nu, ni, nnz, B_.SEED + 31 * rank, device)
    m = WRMF(world * nu, ni, weight=cfg['weight'], reg=cfg['reg'], n_factors=d, verbose=False, seed=1, solver='als', device=device)
    eng = m.engine
    eng.accU = eng.accV = None
    torch.cuda.empty_cache()
    als = DistributedALS(eng, csr, None)
    als.half_sweep('users')                                   # warm-up (sizes the workspace, opens the communicators)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    dist.barrier()
    torch.cuda.synchronize()
    e[0].record()
    als.half_sweep('users')
    e[1].record()
    torch.cuda.synchronize()
    dist.barrier()
    lo, hi = DistributedALS.row_range(world * nu, world, rank)
    e[2].record()
    eng.als_solve_rows(eng.U[lo:hi], eng.V, csr, als.G)       # the rank's solve alone (no Gram, no collectives)
    e[3].record()
    torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, solve_ms = float(t[0]), float(t[1])
    nz = csr.nnz
    out = dict(workload='configs[3] at %d GPUs: WRMF weighted ALS user half-sweep, %d users per GPU (%d in total, of 10M) x %d items, %d '
                        'interactions per GPU, d=%d, weight %.1f, reg %.1f; users range-sharded, tables replicated' % (
                            world, nu, world * nu, ni, nz, d, cfg['weight'], cfg['reg']),
               value=world * nu / (ms * 1e-3), unit='rows solved/s', ms_per_half_sweep=ms, solve_only_ms=solve_ms,
               exchange_ms=ms - solve_ms, nnz_per_gpu=nz,
               all_gather_bytes_per_gpu=float(world * nu * d * 4),
               note='half-sweep = partial Gram + all-reduce(64 KB) + local solve + all-gather of the solved rows (NCCL); '
                    'solve_only_ms is the rank\'s cf_als_solve_rows alone, max over ranks')
    del m, eng, csr, als
    torch.cuda.empty_cache()
    return out


def run_distributed(args, rank, world, device):
    import bench as B_
    wl = dict(B_.WORKLOADS[args.workload])
    if wl['model'] not in ('cml', 'bpr'):
        raise SystemExit('the sharded path supports the cml / bpr workloads')
    n_items_global = wl['n_items'] * (world if getattr(args, 'grow_catalogue', False) else 1)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    pk = B_.peaks()
    clk = B_.ClockSampler(device.index)
    if rank == 0:
        clk.start()
        time.sleep(0.3)
    model, csr, tr, r = sharded_training(B_, wl, args, rank, world, device, n_items_global, K, Wm)
    clocks = clk.stop(r['t0'], r['t1']) if rank == 0 else None
    topk = None
    if args.topk_users > 0:
        # sub-objects must not cost the line (the same sizes on every rank: a failure lands on every rank alike)
        try:
            Tq = min(args.topk_users, wl['n_users'])
            topk = user_sharded_topk_bench(B_, model, csr, n_items_global, wl, Tq, rank, world, device, pk)
            topk['item_sharded'] = sharded_topk(B_, model, csr, n_items_global, wl, Tq, rank, world, device, pk)
        except Exception as e:
            topk = dict(topk or {}, **B_.sub_error(e))
    line = step_line(B_, wl, args, world, K, Wm, r, n_items_global, tr, pk)
    tr.close()
    del model, csr, tr
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs[4] at this N: BPRMF, 12.5M users + 625M interactions per GPU, 10M items row-sharded
    c5 = None
    if args.workload == 'c2' and not args.no_c5:
        w5 = dict(B_.WORKLOADS['c5'])
        K5 = max(3, min(K, args.other_steps))
        m5 = csr5 = tr5 = None
        try:
            m5, csr5, tr5, r5 = sharded_training(B_, w5, args, rank, world, device, w5['n_items'], K5, 3, want_e2e=False)
            c5 = step_line(B_, w5, args, world, K5, 3, r5, w5['n_items'], tr5, pk)
            c5['workload'] = w5['desc'] + ' -- at %d GPUs: %d users, %d interactions in total' % (world, world * w5['n_users'], world * csr5.nnz)
            if args.topk_users > 0:
                c5['topk'] = user_sharded_topk_bench(B_, m5, csr5, w5['n_items'], w5, min(args.topk_users, w5['n_users']), rank, world, device, pk)
                c5['topk']['item_sharded'] = sharded_topk(B_, m5, csr5, w5['n_items'], w5, min(args.topk_users, w5['n_users']), rank, world, device, pk)
        except Exception as e:
            c5 = dict(c5 or {}, **B_.sub_error(e))
        try:
            if tr5 is not None:
                tr5.close()
        except Exception as e:
            c5 = dict(c5 or {}, close_error=B_.sub_error(e)['error'])
        m5 = csr5 = tr5 = None
        torch.cuda.empty_cache()
    als = None
    if args.workload == 'c2' and not args.no_other_configs:
        try:
            als = sharded_als(B_, B_.ALS_SLICE, rank, world, device, pk)
        except Exception as e:          # a sub-object must not cost the line (the same sizes on every rank: every rank lands here)
            als = dict(error='%s: %s' % (type(e).__name__, str(e)[:200]))
            torch.cuda.empty_cache()
    if rank != 0:
        return
    units = world * B * wl['W'] * K
    cfg = B_.same_config(wl, args)          # the same `config` object as the N = 1 line and the reference arm print
    sharding = ('users, interactions and minibatch PER GPU (weak scaling, batch_pairs per GPU = %d); %d items in total, row-sharded by item %% N '
                '(the engines, evaluation and state); users range-sharded; per minibatch: %s' % (
                    B, n_items_global, 'dense item gradients reduce-scattered over per-GPU replicas of the item table'
                    if line['item_transport'].startswith('item table replicated') else 'item rows and gradient rows exchanged'))
    out = dict(metric=B_.metric_name(wl['d']), value=line['value'], unit='triple updates/s', n_gpus=world, steps=K, warmup=Wm,
               ms_per_step=line['ms_per_step'], higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
               data='synthetic', config=cfg, sharding=sharding, gpu_launches=line['gpu_launches'], item_transport=line['item_transport'],
               e2e=dict(value=units / (r['e2e_ms'] * 1e-3), unit='triple updates/s', ms_per_step=r['e2e_ms'] / K,
                        h2d_bytes_per_step=r['h2d'], d2h_bytes_per_step=8),
               roofline=line['roofline'], nvlink=line['nvlink'], phases_ms_per_step=line['phases_ms_per_step'], topk=topk, c5=c5, c4_als=als,
               cpu_baseline=None, clocks=clocks, loss_first_last=line['loss_first_last'])
    print(json.dumps(out))
