"""bench.py's N > 1 arm: weak scaling of the N=1 workload (same users, interactions and minibatch PER GPU; the item
catalogue keeps its global size and is row-sharded, like BASELINE.json configs[4] keeps 10 M items for 8 x 12.5 M users):
users range-sharded, items sharded by item % N, per-minibatch NCCL all-to-all of requested item rows and their gradients
(SURVEY.md 8e).  `--grow-catalogue` instead multiplies the catalogue by N (every rank owns n_items rows)."""
import json
import sys
import time

import torch
import torch.distributed as dist


def run_distributed(args, rank, world, device):
    import bench as B_
    from collaborativefilteringusingtensorflow_b200.dist import DistributedTrainer, item_shard_rows
    wl = dict(B_.WORKLOADS[args.workload])
    if wl['model'] not in ('cml', 'bpr'):
        raise SystemExit('the sharded path supports the cml / bpr workloads')
    n_items_global = wl['n_items'] * (world if getattr(args, 'grow_catalogue', False) else 1)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    pk = B_.peaks()
    csr = B_.synth_interactions(wl['n_users'], n_items_global, wl['nnz'], 2026 + rank, device)
    local_wl = dict(wl, n_items=item_shard_rows(n_items_global, world, rank))
    model = B_.make_model(local_wl, device, seed=2026 + rank, optimizer=args.optimizer, update='sync')
    sampler = B_.make_sampler(wl, csr, B, 2026 + rank, device)
    tr = DistributedTrainer(model, sampler, n_items_global, world, rank, item_transport=args.item_transport)
    tr.step(Wm)
    model.engine.check_flags()
    warm = sampler.next_chunk(K)      # allocator blocks of the timed region's K-minibatch index buffer (see bench.py)
    del warm
    torch.cuda.synchronize()
    dist.barrier()

    clk = B_.ClockSampler(device.index)
    if rank == 0:
        clk.start()
        time.sleep(0.3)
    l0, s0, b0, p0 = tr.launches, sampler.launches, tr.bytes_sent, tr.bytes_pulled
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
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = (tr.launches - l0) + (sampler.launches - s0)
    if args.phases and rank == 0:
        evs, tr.step_events = [e0] + tr.step_events, None
        print('per-minibatch ms: ' + ' '.join('%.2f' % evs[k].elapsed_time(evs[k + 1]) for k in range(len(evs) - 1)), file=sys.stderr)
    sent = (tr.bytes_sent - b0) / K
    pulled = (tr.bytes_pulled - p0) / K * (world - 1) / world     # the share of the item rows that lives on other GPUs
    model.engine.check_flags()

    # e2e: host index buffers -> H2D -> sharded step -> D2H loss, every step
    host = [t.cpu().pin_memory() for t in sampler.next_chunk(K)]
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        dev = [t[k * B:(k + 1) * B].to(device, non_blocking=True) for t in host]
        _ = tr.step_chunk(dev[0], dev[1], B).cpu()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=device)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2 = float(ms2.item())
    # ---- evaluation: item-sharded full-catalogue top-100 of rank 0's users, merged with an all-gather
    topk = None
    if args.topk_users > 0:
        from collaborativefilteringusingtensorflow_b200.dist import distributed_topk, shard_mask_csr
        from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
        Tq = min(args.topk_users, wl['n_users'])
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
        distributed_topk(eng, q[:1024], 100, shard_mask_csr(DeviceCSR(ptr[:1025].clone(), ind[:int(ptr[1024])], rws[:int(ptr[1024])],
                                                                      None, (1024, n_items_global)), world, rank), world, rank, method='tensor')
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        gi, gv = distributed_topk(eng, q, 100, mask, world, rank, method='tensor')
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        tms = float(tms.item())
        fl = 2.0 * n_items_global * wl['d'] * Tq
        topk = dict(metric='users/s full-catalog top-100 (mask train items), items sharded over %d GPUs, all-gather merge' % world,
                    value=Tq / (tms * 1e-3), users=Tq, n_items=n_items_global, ms=tms, tflops_aggregate=fl / (tms * 1e-3) / 1e12,
                    fallback_rows_rank0=int(eng.tc_stats[0].item()))
    phases = None
    if args.phases:
        tr.phase_ms = {}
        tr.step(5)
        phases = {k: v / 5 for k, v in tr.phase_ms.items()}
        tr.phase_ms = None
    if rank != 0:
        return
    clocks = clk.stop(t0, t1)
    units = world * B * wl['W'] * K
    bpp = B_.bytes_per_pair(wl['model'], wl['d'], wl['W'], wl['G'], args.optimizer)
    achieved = bpp * B / (ms / K * 1e-3) / 1e9            # per GPU, whole sharded step (exchange included)
    out = dict(metric='triple updates/s (fused pairwise-ranking step incl. on-device sampling) @d=%d' % wl['d'],
               value=units / (ms * 1e-3), unit='triple updates/s', n_gpus=world, steps=K, warmup=Wm, ms_per_step=ms / K,
               higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=dict(workload=wl['desc'] + ' -- users, interactions and minibatch PER GPU; %d items in total, row-sharded by '
                           'item %% N; users range-sharded; NCCL all-to-all of item rows + gradients per minibatch'
                           % n_items_global,
                           batch_pairs_per_gpu=B, negatives=wl['W'], optimizer=args.optimizer, update='sync',
                           item_transport=('peer (fused NVLink reads in k_step)' if tr._pull else 'nccl (all-to-all of unique rows)'),
                           l2='inputs larger than L2 (random rows of GB-sized tables)'),
               gpu_launches=launches,
               e2e=dict(value=units / (ms2 * 1e-3), unit='triple updates/s', ms_per_step=ms2 / K,
                        h2d_bytes_per_step=sum(int(t[:B].numel()) * t.element_size() for t in host), d2h_bytes_per_step=8),
               roofline=dict(bound='hbm', kernel='whole sharded step per GPU (k_count + k_step + k_apply_staged + exchange + owner apply)',
                             achieved=achieved, peak=pk['hbm'], unit='GB/s', frac=achieved / pk['hbm'], traffic=None,
                             peak_source=pk['source']),
               nvlink=dict(bytes_pulled_per_step_per_gpu=pulled,
                           bytes_sent_per_step_per_gpu=sent, achieved_GBs=sent / (ms / K * 1e-3) / 1e9,
                           peak_GBs_per_direction=770.0, note='rows out + gradients back + ids; measured peer copy 770 GB/s/dir'),
               phases_ms_per_step=phases, topk=topk, cpu_baseline=None, clocks=clocks, loss_first_last=[float(losses[0]), float(losses[-1])])
    print(json.dumps(out))
