#!/usr/bin/env python
"""bench.py -- the hot path's headline number on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Metric (BASELINE.json): triple updates/s of the fused pairwise-ranking training step at d=128, incl. on-device
sampling; secondary: users/s of full-catalog masked top-100.  Workload at N=1 = BASELINE.json configs[1]:
CML, synthetic 1M users x 500k items, d=128, 100M interactions, W=5 negatives, hinge margin 1.0 + rank weight,
reg_cov 1.0, unit-norm clip, the reference's optimizer (TF1 Adagrad) and minibatch-synchronous semantics.
A "step" is one minibatch of B pairs (one counting kernel + one fused step kernel; the sampler kernel that
generates the K minibatches' indices on the device is inside the timed region too).

Prints ONE JSON line (see the contract in the task description): value/unit, e2e (host index buffers ->
H2D -> step -> D2H loss, every step), roofline of the dominant kernel (CUDA-event time of the fused step kernel),
cpu_baseline (the numpy oracle port of the reference's TF1 step, on this box's host cores), clocks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, n_users, n_items, nnz, d, W, G, hyper)
    'c2': dict(model='cml', n_users=1_000_000, n_items=500_000, nnz=100_000_000, d=128, W=5, G=0,
               hyper=dict(reg_cov=1.0, margin=1.0, use_rank_weight=True, clip_norm=1.0, lr=0.1),
               desc='configs[1]: CML synthetic 1M users x 500k items, d=128, 100M interactions, W=5, margin 1.0, '
                    'rank weight, reg_cov 1.0, unit-norm clip, TF1-Adagrad, minibatch-synchronous'),
    'c2-bpr': dict(model='bpr', n_users=1_000_000, n_items=500_000, nnz=100_000_000, d=128, W=1, G=0,
                   hyper=dict(reg=0.1, lr=0.1), desc='BPRMF on the configs[1] shape, d=128, W=1'),
    'c3': dict(model='gbpr', n_users=138_493, n_items=26_744, nnz=20_000_000, d=64, W=5, G=3,
               hyper=dict(reg=0.01, rho=0.4, lr=0.1), desc='configs[2]: GBPR ML-20M shape, G=3, d=64, W=5'),
    'c5': dict(model='bpr', n_users=12_500_000, n_items=10_000_000, nnz=625_000_000, d=128, W=1, G=0,
               hyper=dict(reg=0.1, lr=0.1),
               desc='configs[4] per GPU: BPRMF, 12.5M users and 625M interactions per GPU (100M users / 5B interactions at 8 GPUs), '
                    '10M items, d=128, W=1, reg 0.1, TF1-Adagrad, minibatch-synchronous'),
    'small': dict(model='cml', n_users=20_000, n_items=10_000, nnz=1_000_000, d=128, W=5, G=0,
                  hyper=dict(reg_cov=1.0, margin=1.0, use_rank_weight=True, clip_norm=1.0, lr=0.1),
                  desc='smoke-size CML'),
}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=float(j['hbm_gbs']), bf16=float(j['bf16_tflops']), bf16_sustained=float(j['bf16_tflops_sustained']),
                    sm_max=float(j.get('sm_max_mhz', 1965.0)), source='MEASURED_PEAKS.json (measured)')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max=1965.0, source='B200_PROFILING.md fallback')


def bytes_per_pair(model, d, W, G, optimizer='adagrad'):
    """rows touched R = 2 + W + G, row bytes 4d; adagrad: 4 * R * 4d (params + accumulators, r + w)."""
    R = 2 + (0 if model == 'wrmf' else W) + G
    return (4 if optimizer == 'adagrad' else 2) * R * 4 * d


# ---------------------------------------------------------------------------------------------- synthetic data
def synth_interactions(n_users, n_items, nnz, seed, device):
    """user degree ~ Zipf(1.0) truncated to [1, min(ni/2, 10*mean)], items ~ Zipf(0.8) popularity,
    no per-user duplicates, sorted CSR on the device.  Data generation only (torch ops), not part of the hot path."""
    import torch
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    mean = nnz / n_users
    cap = max(1.0, min(n_items / 2, 10 * mean))
    ranks = torch.arange(1, n_users + 1, device=device, dtype=torch.float64)
    lo, hi = 0.0, float(nnz) * 10
    for _ in range(60):        # scale c so that sum clamp(c / rank, 1, cap) = nnz (oversampled 3% for the dedupe)
        c = 0.5 * (lo + hi)
        tot = float(torch.clamp(c / ranks, 1.0, cap).sum())
        lo, hi = (c, hi) if tot < nnz * 1.03 else (lo, c)
    deg = torch.clamp(c / ranks, 1.0, cap).round().to(torch.int64)
    deg = deg[torch.randperm(n_users, device=device, generator=g)]
    users = torch.repeat_interleave(torch.arange(n_users, device=device, dtype=torch.int64), deg)
    pop = torch.arange(1, n_items + 1, device=device, dtype=torch.float64) ** -0.8
    cdf = torch.cumsum(pop / pop.sum(), 0)
    r = torch.rand(users.numel(), device=device, generator=g, dtype=torch.float64)
    items = torch.searchsorted(cdf, r).clamp_(max=n_items - 1)
    del r
    items = torch.randperm(n_items, device=device, generator=g)[items]      # popular items are not the low ids
    key = torch.unique(users * n_items + items)                               # sorted, per-user duplicates dropped
    del users, items
    if key.numel() > nnz:                                                     # thin uniformly down to nnz
        keep = torch.randperm(key.numel(), device=device, generator=g)[:nnz]
        key = key[torch.sort(keep).values]
    rows = (key // n_items).to(torch.int32)
    cols = (key % n_items).to(torch.int32)
    counts = torch.bincount(rows.to(torch.int64), minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return DeviceCSR(indptr, cols, rows, None, (n_users, n_items))


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_baseline(wl, B, csr_host, budget_s=20.0, max_steps=8, seed=2026, threads=None):
    """Times the CPU restatement of the reference's TF1 step + its sampler on this box's host cores: same tables, same
    B, same W; whole-table clip every step as cml.py:119-129.  BPR / CML use the torch port (oracle/steps_torch.py) on
    all host cores; the numpy restatement (oracle/steps.py, one core) serves GBPR."""
    from oracle import steps
    import torch
    use_torch = wl['model'] in ('cml', 'bpr')
    if use_torch:
        from oracle import steps_torch
        torch.set_num_threads(threads or os.cpu_count())
    rng = np.random.default_rng(seed)
    nu, ni, d, W, G = wl['n_users'], wl['n_items'], wl['d'], wl['W'], wl['G']
    indptr, indices, rows = csr_host
    U = (0.1 * rng.standard_normal((nu, d), dtype=np.float32))
    V = (0.1 * rng.standard_normal((ni, d), dtype=np.float32))
    aU, aV = np.full_like(U, 0.1), np.full_like(V, 0.1)
    b = ab = None
    if wl['model'] == 'gbpr':
        b, ab = (0.1 * rng.standard_normal(ni, dtype=np.float32)), np.full(ni, 0.1, np.float32)
    nnz = len(indices)
    allkeys = rows.astype(np.int64) * ni + indices           # the reference's dict user -> set(items), built in __init__
    done, t_total = 0, 0.0
    while done < max_steps and (done < 1 or t_total < budget_s):
        t0 = time.perf_counter()
        # sampler_ranking.py:22-37 restated: shuffled positives + rejection-sampled negatives
        p = rng.integers(0, nnz, B)
        pairs = np.stack([rows[p], indices[p]], 1)
        negs = rng.integers(0, ni, (B, W))
        for _ in range(64):                                  # re-draw while the negative is a positive (:35-36)
            key = (pairs[:, 0:1].astype(np.int64) * ni + negs).ravel()
            pos = np.minimum(np.searchsorted(allkeys, key), nnz - 1)
            bad = (allkeys[pos] == key).reshape(B, W)
            if not bad.any():
                break
            negs[bad] = rng.integers(0, ni, int(bad.sum()))
        h_ = wl['hyper']
        if use_torch and done == 0:
            tU, tV, taU, taV = (torch.from_numpy(x) for x in (U, V, aU, aV))   # share memory with the numpy tables
        if wl['model'] == 'cml':
            steps_torch.cml_step(tU, tV, taU, taV, torch.from_numpy(pairs), torch.from_numpy(negs), h_['lr'], h_['reg_cov'],
                                 h_['margin'], h_['use_rank_weight'], h_['clip_norm'])
        elif wl['model'] == 'bpr':
            steps_torch.bpr_step(tU, tV, taU, taV, torch.from_numpy(pairs), torch.from_numpy(negs), h_['lr'], h_['reg'])
        elif wl['model'] == 'gbpr':
            group = rng.integers(0, nu, (B, G))
            steps.gbpr_step(U, V, b, aU, aV, ab, pairs, negs, group, h_['lr'], h_['reg'], h_['rho'])
        t_total += time.perf_counter() - t0
        done += 1
    return done, t_total, (torch.get_num_threads() if use_torch else 1)


# ---------------------------------------------------------------------------------------------- main
def make_model(wl, device, seed=2026, **kw):
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF
    h = wl['hyper']
    common = dict(n_factors=wl['d'], verbose=False, seed=seed, device=device, lr=h['lr'], **kw)
    if wl['model'] == 'cml':
        return CML(wl['n_users'], wl['n_items'], reg_cov=h['reg_cov'], margin=h['margin'],
                   use_rank_weight=h['use_rank_weight'], clip_norm=h['clip_norm'], **common)
    if wl['model'] == 'bpr':
        return BPRMF(wl['n_users'], wl['n_items'], reg=h['reg'], **common)
    return GBPRMF(wl['n_users'], wl['n_items'], rho=h['rho'], gsize=wl['G'], reg=h['reg'], **common)


def make_sampler(wl, csr, B, seed, device):
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_gbpr, sampler_ranking
    if wl['model'] == 'gbpr':
        return sampler_gbpr.Sampler(csr, wl['G'], wl['W'], B, seed=seed, device=device)
    return sampler_ranking.Sampler(csr, wl['W'], B, seed=seed, device=device)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        from bench_dist import run_distributed          # multi-GPU path (row-sharded tables + all-to-all)
        return run_distributed(args, rank, world, device)

    wl = WORKLOADS[args.workload]
    B, K, Wm = args.batch, args.steps, args.warmup
    pk = peaks()
    t_setup = time.time()
    csr = synth_interactions(wl['n_users'], wl['n_items'], wl['nnz'], 2026, device)
    model = make_model(wl, device, optimizer=args.optimizer, update=args.update)
    eng = model.engine
    sampler = make_sampler(wl, csr, B, 2026, device)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    def run_steps(n, profile=None):
        chunk = sampler.next_chunk(n)                       # ONE sampler launch for the n minibatches
        return model._train_arrays(chunk, B) if profile is None else \
            eng.train_batches(chunk[0], chunk[1], chunk[2] if len(chunk) > 2 else None, batch_size=B, profile=profile)

    # ---- warm-up (also performs CML's one-time whole-table clip)
    run_steps(max(Wm, 3))
    eng.check_flags()
    sampler.check_flags()
    # the timed region samples K minibatches with ONE launch into one [K * B, 2 + W] index buffer: have the caching
    # allocator own blocks of that size already (a first-time cudaMalloc of ~3 GB inside the timed region cost up to
    # 1.3 ms per step in some runs)
    warm = sampler.next_chunk(K)
    del warm
    torch.cuda.synchronize()

    # ---- timed: exactly K steps, CUDA events on the launching stream
    clk = ClockSampler(local)
    clk.start()
    time.sleep(0.3)
    l0, s0 = eng.launches, sampler.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    losses = run_steps(K)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = (eng.launches - l0) + (sampler.launches - s0)
    clocks = clk.stop(t0, t1)
    eng.check_flags()
    units = B * wl['W'] * K
    value = units / (ms * 1e-3)

    # ---- per-kernel roofline: CUDA events around every kernel of K more steps
    prof = {}
    run_steps(K, prof)
    bpp = bytes_per_pair(wl['model'], wl['d'], wl['W'], wl['G'], args.optimizer)
    step_ms = prof['step_ms'] / prof['n_batches']
    apply_ms = prof.get('apply_ms', 0.0) / prof['n_batches']
    # The algorithmic bytes of a minibatch (4 * R * 4d per pair) are moved by the fused step kernel and, for
    # rows that occur more than once, by the staged-apply kernel that follows it: the roofline is quoted on their sum.
    achieved = bpp * B / ((step_ms + apply_ms) * 1e-3) / 1e9
    # DRAM bytes per launch (k_step + k_apply_staged) from the committed `ncu --set full` capture of this exact configuration
    traffic = None
    if args.workload == 'c2' and B == (1 << 20) and args.optimizer == 'adagrad' and args.update == 'sync':
        traffic = (6.952 + 3.395 + 1.177 + 1.058) * 1e9     # profiles/r1_ncu_full_c2_cml_B1M.txt
    roofline = dict(bound='hbm', kernel='cfstep::k_step<%s> + cfstep::k_apply_staged' % wl['model'], achieved=achieved,
                    peak=pk['hbm'], unit='GB/s', frac=achieved / pk['hbm'], traffic=traffic,
                    traffic_source='profiles/r1_ncu_full_c2_cml_B1M.txt (dram__bytes_read + write; below the algorithmic bytes '
                                   'because rows that occur more than once in a minibatch share their traffic)' if traffic else None,
                    peak_source=pk['source'],
                    k_step_only_GBs=bpp * B / (step_ms * 1e-3) / 1e9,
                    algorithmic_bytes_per_launch=bpp * B, kernel_ms_per_launch=step_ms,
                    count_kernel_ms_per_launch=prof['count_ms'] / prof['n_batches'],
                    apply_kernel_ms_per_launch=prof.get('apply_ms', 0.0) / prof['n_batches'],
                    kernel_share_of_step=(step_ms + apply_ms) / (ms / K))

    # ---- e2e: host (pinned) index buffers -> H2D -> step -> D2H loss, every step, through the public engine API
    host_chunk = [t.cpu().pin_memory() for t in sampler.next_chunk(K)]
    torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        dev = [t[k * B:(k + 1) * B].to(device, non_blocking=True) for t in host_chunk]
        loss_k = model._train_arrays(dev, B)
        _ = loss_k.cpu()                                     # D2H of the step's result
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    h2d = sum(int(t[:B].numel()) * t.element_size() for t in host_chunk)
    e2e = dict(value=units / (ms_e2e * 1e-3), unit='triple updates/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
               ms_per_step=ms_e2e / K)

    # ---- secondary metric: users/s of full-catalog masked top-100 (tcgen05/TMA candidate pass + exact fp64 re-rank)
    topk = None
    if args.topk_users > 0:
        def time_topk(engine, users, mask, reps=2):
            engine.topk(users[:1024], 100, mask, method='tensor')
            torch.cuda.synchronize()
            best = None
            for _ in range(reps):
                e0.record()
                engine.topk(users, 100, mask, method='tensor')
                e1.record()
                torch.cuda.synchronize()
                best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
            st = engine.tc_stats.cpu().numpy()
            return best, int(st[0]), float(st[1]) / max(1, len(users) - int(st[0]))

        Tq = args.topk_users
        users = torch.randperm(wl['n_users'], device=device)[:Tq].to(torch.int32)
        tk_ms, fb, cand = time_topk(eng, users, csr)
        flops = 2.0 * wl['n_items'] * wl['d'] * Tq
        topk = dict(metric='users/s full-catalog top-100 (mask train items), exact result via tensor-core candidate pass',
                    value=Tq / (tk_ms * 1e-3), users=Tq, n_items=wl['n_items'], ms=tk_ms,
                    kernels='k_prep x2 + k_topk_tc (tcgen05.mma bf16 + TMA) + k_rerank (fp64) + k_topk_exact (fallback rows)',
                    tflops=flops / (tk_ms * 1e-3) / 1e12, frac_of_bf16_peak=flops / (tk_ms * 1e-3) / 1e12 / pk['bf16'],
                    peak_tflops=pk['bf16'], fallback_rows=fb, candidates_per_row=cand)
        if args.topk_c5_items > 0:
            # configs[4]'s catalogue size on one GPU (item-sharded over P GPUs: x P): BPRMF scoring, 10M items, d=128
            from collaborativefilteringusingtensorflow_b200.engine import FactorEngine
            big = FactorEngine('bpr', Tq, args.topk_c5_items, 128, device, seed=7)
            uq = torch.arange(Tq, dtype=torch.int32, device=device)
            ms5, fb5, cand5 = time_topk(big, uq, None, reps=2)
            fl5 = 2.0 * args.topk_c5_items * 128 * Tq
            topk['c5_catalogue'] = dict(n_items=args.topk_c5_items, d=128, users=Tq, ms=ms5, value=Tq / (ms5 * 1e-3),
                                        tflops=fl5 / (ms5 * 1e-3) / 1e12, frac_of_bf16_peak=fl5 / (ms5 * 1e-3) / 1e12 / pk['bf16'],
                                        fallback_rows=fb5, candidates_per_row=cand5)
            del big
            torch.cuda.empty_cache()

    # ---- CPU baseline: the oracle port on this box's cores, bounded sample
    cpub = None
    if not args.no_cpu_baseline:
        csr_host = (csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), csr.rows.cpu().numpy())
        Bc = min(B, args.cpu_batch)
        n_cpu, t_cpu, cores = cpu_baseline(wl, Bc, csr_host, budget_s=args.cpu_budget)
        cpub = dict(value=n_cpu * Bc * wl['W'] / t_cpu, unit='triple updates/s', cores=cores, kind='port',
                    sample='%d minibatches of B=%d pairs x W=%d (torch-CPU oracle port of the TF1 step incl. '
                           'whole-table clip + numpy rejection sampler), %.1f s; host has %d cores'
                           % (n_cpu, Bc, wl['W'], t_cpu, os.cpu_count()))

    out = dict(metric='triple updates/s (fused pairwise-ranking step incl. on-device sampling) @d=%d' % wl['d'],
               value=value, unit='triple updates/s', n_gpus=1, steps=K, warmup=max(Wm, 3), ms_per_step=ms / K,
               higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=dict(workload=wl['desc'], batch_pairs=B, negatives=wl['W'], optimizer=args.optimizer,
                           update=args.update, nnz=csr.nnz, l2='inputs larger than L2 (tables+accumulators %.1f GB, '
                           'random rows)' % (4 * 4 * wl['d'] * (wl['n_users'] + wl['n_items']) / 1e9 / 2),
                           pairs_per_s=value / wl['W'], setup_s=setup_s),
               gpu_launches=launches, e2e=e2e, roofline=roofline, cpu_baseline=cpub, clocks=clocks, topk=topk,
               loss_first_last=[float(losses[0]), float(losses[-1])])
    print(json.dumps(out))


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path.  TensorFlow 1.x cannot be installed here (no
    network, no wheel), so this is the oracle port of the TF1 graph (oracle/steps_torch.py: torch CPU ops on every host
    core, tested equal to the numpy restatement oracle/steps.py) plus the numpy rejection sampler."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    B = min(args.batch, args.cpu_batch)
    # host-side synthetic CSR of the same shape (numpy; a bounded sample of users keeps generation short)
    rng = np.random.default_rng(2026)
    nu, ni = wl['n_users'], wl['n_items']
    deg = np.clip((wl['nnz'] / nu * rng.pareto(1.5, nu)).astype(np.int64), 1, min(ni // 2, 10 * wl['nnz'] // nu))
    deg = (deg * (wl['nnz'] / 4 / deg.sum())).astype(np.int64).clip(1)       # quarter-density sample: sampling cost is per pair
    indptr = np.zeros(nu + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    rows = np.repeat(np.arange(nu, dtype=np.int32), deg)
    indices = rng.integers(0, ni, len(rows)).astype(np.int32)
    order = np.lexsort((indices, rows))
    indices = indices[order]
    steps_done, t, cores = cpu_baseline(wl, B, (indptr, indices, rows), budget_s=max(20.0, args.cpu_budget),
                                        max_steps=max(1, args.steps))
    value = steps_done * B * wl['W'] / t
    cb = dict(value=value, unit='triple updates/s', cores=cores, kind='port',
              sample='%d minibatches of B=%d pairs x W=%d, torch-CPU oracle port of the TF1 step on all host cores (TensorFlow not installable)'
                     % (steps_done, B, wl['W']))
    print(json.dumps(dict(impl='reference', metric='triple updates/s (fused pairwise-ranking step incl. sampling) @d=%d' % wl['d'],
                          value=value, unit='triple updates/s', n_gpus=args.gpus, steps=steps_done, warmup=0,
                          ms_per_step=1e3 * t / steps_done, higher_is_better=True, scaling='weak', vs_baseline=None,
                          dtype='f32', data='synthetic', config=dict(workload=wl['desc'], batch_pairs=B, negatives=wl['W']),
                          cpu_baseline=cb, e2e=dict(value=value, unit='triple updates/s', h2d_bytes_per_step=0,
                                                    d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=1 << 20, help='pairs per minibatch = the models\' batch_size (SURVEY 8d: 2^20)')
    ap.add_argument('--cpu-batch', type=int, default=65536, help='minibatch of the bounded CPU-baseline sample')
    ap.add_argument('--optimizer', default='adagrad', choices=['adagrad', 'sgd'])
    ap.add_argument('--update', default='sync', choices=['sync', 'hogwild'])
    ap.add_argument('--topk-users', type=int, default=37888, help='query users of the top-K measurement (148 SMs x 256 rows)')
    ap.add_argument('--topk-c5-items', type=int, default=10_000_000, help='also time top-100 over a configs[4]-sized catalogue (0 = skip)')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--grow-catalogue', action='store_true', help='N > 1: n_items x N items in total instead of a fixed catalogue')
    ap.add_argument('--item-transport', default='auto', choices=['nccl', 'peer', 'auto'],
                    help='N > 1: how item rows reach the step (NCCL all-to-all of unique rows / NVLink peer reads in the kernel)')
    ap.add_argument('--phases', action='store_true', help='N > 1: also report per-phase times of the sharded step')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
