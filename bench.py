#!/usr/bin/env python
"""bench.py -- the hot path's headline number on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Metric (BASELINE.json): triple updates/s of the fused pairwise-ranking training step at d=128, incl. sampling;
secondary: users/s of full-catalog masked top-100.  Workload at N=1 = BASELINE.json configs[1]:
CML, synthetic 1M users x 500k items, d=128, 100M interactions, W=5 negatives, hinge margin 1.0 + rank weight,
reg_cov 1.0, unit-norm clip, the reference's optimizer (TF1 Adagrad) and minibatch-synchronous semantics.
A "step" is one minibatch of B pairs (one counting kernel + one fused step kernel + one staged-apply kernel; the sampler
kernel that generates the K minibatches' indices on the device is inside the timed region too).

Prints ONE JSON line (see the contract in the task description): value/unit, e2e (pinned host index buffers ->
H2D -> step -> D2H loss, every step), roofline of the dominant kernels (CUDA-event times), cpu_baseline (the oracle port
of the reference's TF1 step on this box's host cores: best-effort and reference-faithful settings), clocks, the top-K
object, and `other_configs`: BPR W=1 (the metric's namesake), configs[2] (GBPR) and a configs[3] (WRMF ALS) slice.
`--impl reference` times the CPU restatement alone (TensorFlow is not installable here) on the same metric and config.
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
# configs[3] slice: 1M of the 10M users, the whole 1M-item catalogue, 50 interactions per user (the configs[3] mean)
ALS_SLICE = dict(n_users=1_000_000, n_items=1_000_000, nnz=50_000_000, d=128, weight=2.0, reg=0.1)
ALS_SMALL = dict(n_users=20_000, n_items=20_000, nnz=1_000_000, d=128, weight=2.0, reg=0.1)
SEED = 2026


def metric_name(d):
    """ONE string for both arms (the driver divides the two arms' numbers only when `metric` is identical)."""
    return 'triple updates/s (fused pairwise-ranking step incl. sampling) @d=%d' % d


def same_config(wl, args):
    """The `config` object both arms print: static description of the workload only (run-dependent facts go elsewhere)."""
    return dict(workload=wl['desc'], batch_pairs=args.batch, negatives=wl['W'], optimizer=args.optimizer, update=args.update,
                generator='bench.synth_interactions(seed %d): user degree ~ Zipf(1.0) clipped to [1, min(ni/2, 10*mean)], items ~ '
                          'Zipf(0.8) popularity, no per-user duplicates (SURVEY 8d)' % SEED,
                l2='inputs larger than L2 (tables+accumulators %.1f GB, random rows)'
                   % (2 * 4 * wl['d'] * (wl['n_users'] + wl['n_items']) / 1e9))


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


def measured_traffic(workload, B, optimizer, update):
    """DRAM bytes per launch of the step kernels from a COMMITTED `ncu --set full` capture of this exact configuration
    (profiles/traffic.json, written by tools/ncu_summary.py --traffic; keyed by configuration) -- None when no capture of
    this configuration is on record."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(p):
        return None, None
    e = json.load(open(p)).get('%s|B=%d|%s|%s' % (workload, B, optimizer, update))
    return (float(e['bytes_per_launch']), e['source']) if e else (None, None)


# ---------------------------------------------------------------------------------------------- synthetic data
def synth_coo(n_users, n_items, nnz, seed, device):
    """user degree ~ Zipf(1.0) truncated to [1, min(ni/2, 10*mean)], items ~ Zipf(0.8) popularity,
    no per-user duplicates; returns the sorted keys user * n_items + item (torch int64, on `device` -- CUDA for the GPU
    arm, CPU for the reference arm: same code, same distribution).  Data generation only, not part of the hot path."""
    import torch
    mean = nnz / n_users
    cap = max(1.0, min(n_items / 2, 10 * mean))
    ranks = torch.arange(1, n_users + 1, device=device, dtype=torch.float64)
    pop = torch.arange(1, n_items + 1, device=device, dtype=torch.float64) ** -0.8
    cdf = torch.cumsum(pop / pop.sum(), 0)
    over = 1.03                # draws per kept interaction: per-user duplicates are dropped, so oversample (and retry
    for attempt in range(4):   # with a larger factor when a small, skewed catalogue loses more than that)
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        lo, hi = 0.0, float(nnz) * 20
        for _ in range(60):    # scale c so that sum clamp(c / rank, 1, cap) = over * nnz
            c = 0.5 * (lo + hi)
            tot = float(torch.clamp(c / ranks, 1.0, cap).sum())
            lo, hi = (c, hi) if tot < nnz * over else (lo, c)
        deg = torch.clamp(c / ranks, 1.0, cap).round().to(torch.int64)
        deg = deg[torch.randperm(n_users, device=device, generator=g)]
        users = torch.repeat_interleave(torch.arange(n_users, device=device, dtype=torch.int64), deg)
        r = torch.rand(users.numel(), device=device, generator=g, dtype=torch.float64)
        items = torch.searchsorted(cdf, r).clamp_(max=n_items - 1)
        del r
        items = torch.randperm(n_items, device=device, generator=g)[items]      # popular items are not the low ids
        key = torch.unique(users * n_items + items)                               # sorted, per-user duplicates dropped
        del users, items
        if key.numel() >= nnz or attempt == 3:
            break
        over *= 1.02 * nnz / key.numel()
    if key.numel() > nnz:                                                         # thin uniformly down to nnz
        keep = torch.randperm(key.numel(), device=device, generator=g)[:nnz]
        key = key[torch.sort(keep).values]
    return key


def synth_interactions(n_users, n_items, nnz, seed, device):
    """synth_coo as a sorted device CSR."""
    import torch
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    key = synth_coo(n_users, n_items, nnz, seed, device)
    rows = (key // n_items).to(torch.int32)
    cols = (key % n_items).to(torch.int32)
    counts = torch.bincount(rows.to(torch.int64), minlength=n_users)
    indptr = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return DeviceCSR(indptr, cols, rows, None, (n_users, n_items))


def synth_host_csr(n_users, n_items, nnz, seed):
    """The same generator on the host (reference arm): (indptr int64, indices int32, rows int32) numpy arrays."""
    key = synth_coo(n_users, n_items, nnz, seed, 'cpu').numpy()
    rows = (key // n_items).astype(np.int32)
    indices = (key % n_items).astype(np.int32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n_users), out=indptr[1:])
    return indptr, indices, rows


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), power_w=float(np.median(pw)) if pw else None)


# ---------------------------------------------------------------------------------------------- CPU baseline (oracle)
def _init_host_tables(wl, seed):
    rng = np.random.default_rng(seed)
    nu, ni, d = wl['n_users'], wl['n_items'], wl['d']
    U = (0.1 * rng.standard_normal((nu, d), dtype=np.float32))
    V = (0.1 * rng.standard_normal((ni, d), dtype=np.float32))
    return rng, U, V, np.full_like(U, 0.1), np.full_like(V, 0.1)


def cpu_baseline(wl, B, csr_host, budget_s=20.0, max_steps=8, min_steps=1, seed=SEED, threads=None):
    """BEST-EFFORT setting: the CPU restatement of the reference's TF1 step + a vectorised numpy rejection sampler on this
    box's host cores, large minibatch B, same tables, same W; whole-table clip every step as cml.py:119-129.  BPR / CML use
    the torch port (oracle/steps_torch.py) on all host cores; the numpy restatement (oracle/steps.py, one core) serves GBPR.
    Returns (steps done, seconds, threads)."""
    from oracle import steps
    import torch
    use_torch = wl['model'] in ('cml', 'bpr')
    if use_torch:
        from oracle import steps_torch
        torch.set_num_threads(threads or os.cpu_count())
    nu, ni, d, W, G = wl['n_users'], wl['n_items'], wl['d'], wl['W'], wl['G']
    indptr, indices, rows = csr_host
    rng, U, V, aU, aV = _init_host_tables(wl, seed)
    b = ab = None
    if wl['model'] == 'gbpr':
        b, ab = (0.1 * rng.standard_normal(ni, dtype=np.float32)), np.full(ni, 0.1, np.float32)
    nnz = len(indices)
    allkeys = rows.astype(np.int64) * ni + indices           # the reference's dict user -> set(items), built in __init__
    done, t_total = 0, 0.0
    while done < max_steps and (done < min_steps or t_total < budget_s):
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


def cpu_baseline_faithful(wl, csr_host, budget_s=6.0, max_steps=400, seed=SEED, sample_users=20_000):
    """REFERENCE-FAITHFUL setting: what testcml.py / testbprmf.py + sampler_ranking.py actually do --
    batch_size 100 (testbprmf.py:21-30), ONE sampler stream producing one batch at a time (sampler_ranking.py:22-37,
    restated by oracle.samplers.ranking_batches: shuffled positives, per-batch rejection of positives), a Python loop with
    one sess.run-equivalent per minibatch (bprmf.py:143-148), whole-table clip after every step (cml.py:119-129).  The
    tables have the workload's full size; the sampler walks the interactions of the first `sample_users` users (building
    the reference's dict / shuffling all 100M pairs is a one-off per-epoch cost, not a per-step one).  TF1's intra-op
    thread pool defaults to every core, so the torch step does too; the sampler is single-threaded like the reference's."""
    import torch
    from scipy.sparse import csr_matrix
    from oracle import samplers as osamp
    from oracle import steps_torch
    torch.set_num_threads(os.cpu_count())
    B = 100
    indptr, indices, _ = csr_host
    nu_s = min(sample_users, wl['n_users'])
    nz = int(indptr[nu_s])
    sub = csr_matrix((np.ones(nz, np.float32), indices[:nz], indptr[:nu_s + 1]), shape=(nu_s, wl['n_items']))
    gen = osamp.ranking_batches(sub, wl['W'], B, seed)
    _, U, V, aU, aV = _init_host_tables(wl, seed)
    tU, tV, taU, taV = (torch.from_numpy(x) for x in (U, V, aU, aV))
    h_ = wl['hyper']
    next(gen)                                                # the first call shuffles the epoch (not timed)
    done, t_total = 0, 0.0
    while done < max_steps and (done < 3 or t_total < budget_s):
        t0 = time.perf_counter()
        pairs, negs = next(gen)
        if wl['model'] == 'cml':
            steps_torch.cml_step(tU, tV, taU, taV, torch.from_numpy(pairs), torch.from_numpy(negs), h_['lr'], h_['reg_cov'],
                                 h_['margin'], h_['use_rank_weight'], h_['clip_norm'])
        else:
            steps_torch.bpr_step(tU, tV, taU, taV, torch.from_numpy(pairs), torch.from_numpy(negs), h_['lr'], h_['reg'])
        t_total += time.perf_counter() - t0
        done += 1
    return dict(value=done * B * wl['W'] / t_total, unit='triple updates/s', cores=torch.get_num_threads(), kind='port',
                setting='reference-faithful', batch_pairs=B, steps=done, ms_per_step=1e3 * t_total / done,
                sample='%d minibatches of B=100 pairs x W=%d (testbprmf.py:21-30), one sampler stream (oracle.samplers.'
                       'ranking_batches = sampler_ranking.py:22-37) over the first %d users, one step per Python-loop iteration, '
                       'torch-CPU port of the TF1 step incl. the whole-table clip on the full-size tables, %.1f s'
                       % (done, wl['W'], nu_s, t_total))


# ---------------------------------------------------------------------------------------------- main
def make_model(wl, device, seed=SEED, **kw):
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


def step_kernel_name(wl, update):
    """Which step kernel cf_train_steps launches for this workload (cf_step.cu: the specialised forms of cf_step_fast.cu serve
    one negative per pair and GBPR with 5 negatives and a group of 3 or 1, SYNC mode, rows of 36..128 floats)."""
    fast = update == 'sync' and 32 < wl['d'] <= 128 and os.environ.get('CF_STEP_GENERIC', '0') in ('', '0') and (
        (wl['model'] in ('bpr', 'cml') and wl['W'] == 1) or (wl['model'] == 'gbpr' and wl['W'] == 5 and wl['G'] in (1, 3)))
    return 'k_step_fast' if fast else 'k_step'


def time_training(wl, csr, B, K, Wm, device, optimizer, update, pk):
    """Warm-up + exactly K timed minibatches (CUDA events on the launching stream) + per-kernel event times of K more.
    Returns (model, sampler, dict)."""
    import torch
    model = make_model(wl, device, optimizer=optimizer, update=update, batch_size=B)
    eng = model.engine
    sampler = make_sampler(wl, csr, B, SEED, device)

    def run_steps(n, profile=None):
        if profile is None:
            return model._epoch(sampler, n)                 # the product's own epoch loop: one sampler launch per minibatch
                                                            # (at B = 2^20), issued on a side stream one minibatch ahead
        chunk = sampler.next_chunk(n)
        return eng.train_batches(chunk[0], chunk[1], chunk[2] if len(chunk) > 2 else None, batch_size=B, profile=profile)

    run_steps(max(Wm, 3))                                   # warm-up (also CML's one-time whole-table clip)
    eng.check_flags()
    sampler.check_flags()
    # the timed region samples K minibatches with ONE launch into one [K * B, 2 + W] index buffer: have the caching
    # allocator own blocks of that size already (a first-time cudaMalloc of GBs inside the timed region costs ms)
    warm = sampler.next_chunk(K)
    del warm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0, s0 = eng.launches, sampler.launches
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    losses = run_steps(K)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = (eng.launches - l0) + (sampler.launches - s0)
    eng.check_flags()
    prof = {}
    run_steps(K, prof)
    nb = prof['n_batches']
    step_ms, apply_ms, count_ms = prof['step_ms'] / nb, prof.get('apply_ms', 0.0) / nb, prof['count_ms'] / nb
    bpp = bytes_per_pair(wl['model'], wl['d'], wl['W'], wl['G'], optimizer)
    # The algorithmic bytes of a minibatch (4 * R * 4d per pair) are moved by the fused step kernel and, for
    # rows that occur more than once, by the staged-apply kernel that follows it: the roofline is quoted on their sum.
    achieved = bpp * B / ((step_ms + apply_ms) * 1e-3) / 1e9
    whole = bpp * B / (ms / K * 1e-3) / 1e9
    out = dict(ms=ms, t0=t0, t1=t1, launches=launches, losses=losses, units=B * wl['W'] * K,
               roofline=dict(bound='hbm', kernel='cfstep::%s<%s> + cfstep::k_apply_staged' % (step_kernel_name(wl, update), wl['model']), achieved=achieved,
                             peak=pk['hbm'], unit='GB/s', frac=achieved / pk['hbm'], traffic=None, peak_source=pk['source'],
                             algorithmic_bytes_per_launch=bpp * B, kernel_ms_per_launch=step_ms + apply_ms,
                             step_kernel_ms=step_ms, apply_kernel_ms=apply_ms, count_kernel_ms=count_ms,
                             whole_step_GBs=whole, whole_step_frac=whole / pk['hbm'],
                             kernel_share_of_step=(step_ms + apply_ms) / (ms / K)))
    return model, sampler, out


def time_e2e(model, sampler, B, K, device, step=None, pre=None):
    """The same K minibatches through the public API from HOST buffers: every step's index arrays come from pinned host
    memory (H2D on a copy stream, double-buffered so that the copy of minibatch k+1 runs under step k) and every step's
    loss is read back into pinned host memory (async D2H, one event wait at the end).  The metric includes sampling, so
    every step ALSO launches the on-device sampler for one minibatch (the batches that are stepped on are the host ones,
    like a reference `next_batch()` result fed through feed_dict)."""
    import torch
    host = [t.cpu().pin_memory() for t in sampler.next_chunk(K)]
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device=device)
    bufs = [[torch.empty_like(t[:B], device=device) for t in host] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(K, dtype=torch.float64).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step = step or (lambda bufs_: model._train_arrays(bufs_, B))
    torch.cuda.synchronize()
    if pre is not None:
        pre()                                                # N > 1: barrier, so that every rank starts together
        torch.cuda.synchronize()
    e0.record()
    copy.wait_event(e0)
    for k in range(K):
        s = k & 1
        with torch.cuda.stream(copy):
            if k >= 2:
                copy.wait_event(free[s])                     # step k-2 has consumed this buffer
            for dst, src in zip(bufs[s], host):
                dst.copy_(src[k * B:(k + 1) * B], non_blocking=True)
            ready[s].record(copy)
        main.wait_event(ready[s])
        sampled = sampler.next_chunk(1)                      # the sampling work of this step (metric: "incl. sampling")
        loss_k = step(bufs[s])
        loss_host[k:k + 1].copy_(loss_k, non_blocking=True)  # D2H of the step's result
        free[s].record(main)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    h2d = sum(int(t[:B].numel()) * t.element_size() for t in host)
    return ms, h2d, loss_host


def time_topk(engine, users, mask, K=100, warm=4096):
    """users/s of the tensor-core top-K over all `users` in ONE call (whole call: operand prep, tcgen05 candidate pass,
    exact re-rank, fallback rows); a short warm-up call first; plain time of the one timed call.  SM clocks and power are
    sampled during the timed call: a long tensor-core sweep runs into the board's power cap (sw_power_cap, SM clock well
    below max), which is why the fraction of the SUSTAINED tensor peak is reported next to the burst one."""
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.topk(users[:warm], K, mask, method='tensor')
    engine.topk(users, K, mask, method='tensor')             # sizes the workspace (first-time cudaMalloc) -- not timed
    torch.cuda.synchronize()
    clk = ClockSampler(engine.device.index or 0)
    clk.start()
    time.sleep(0.7)                                          # (nvidia-smi needs a few hundred ms before its first sample)
    t0 = time.time()
    e0.record()
    engine.topk(users, K, mask, method='tensor')
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = clk.stop(t0, t1)
    st = engine.tc_stats.cpu().numpy()
    return e0.elapsed_time(e1), int(st[0]), float(st[1]) / max(1, len(users) - int(st[0])), clocks


def topk_object(ms, T, n_items, d, fb, cand, pk, **extra):
    fl = 2.0 * n_items * d * T
    tf = fl / (ms * 1e-3) / 1e12
    return dict(value=T / (ms * 1e-3), unit='users/s', users=T, n_items=n_items, d=d, K=100, ms=ms, tflops=tf,
                frac_of_tensor_peak=tf / pk['bf16'], frac_of_sustained_tensor_peak=tf / pk['bf16_sustained'],
                peak_tflops=pk['bf16'], fallback_rows=fb, candidates_per_row=cand, **extra)


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
        from bench_dist import run_distributed          # multi-GPU path (row-sharded tables + exchange)
        return run_distributed(args, rank, world, device)

    wl = WORKLOADS[args.workload]
    B, K, Wm = args.batch, args.steps, args.warmup
    pk = peaks()
    t_setup = time.time()
    csr = synth_interactions(wl['n_users'], wl['n_items'], wl['nnz'], SEED, device)
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    clk = ClockSampler(local)
    clk.start()
    time.sleep(0.3)
    model, sampler, tr = time_training(wl, csr, B, K, Wm, device, args.optimizer, args.update, pk)
    clocks = clk.stop(tr['t0'], tr['t1'])
    eng = model.engine
    ms, units, losses = tr['ms'], tr['units'], tr['losses']
    value = units / (ms * 1e-3)
    roofline = tr['roofline']
    traffic, tsrc = measured_traffic(args.workload, B, args.optimizer, args.update)
    roofline['traffic'] = traffic
    roofline['traffic_source'] = tsrc

    # ---- e2e: pinned host index buffers -> H2D -> step -> D2H loss, every step, through the public engine API
    ms_e2e, h2d, _ = time_e2e(model, sampler, B, K, device)
    e2e = dict(value=units / (ms_e2e * 1e-3), unit='triple updates/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
               ms_per_step=ms_e2e / K, pipeline='per step: on-device sampling of one minibatch + H2D of the minibatch\'s index arrays '
                                                'from pinned host memory (copy stream, double-buffered: the copy of minibatch k+1 runs '
                                                'under step k) + the three step kernels + async D2H of the loss')

    # ---- secondary metric: users/s of full-catalog masked top-100 (tcgen05/TMA candidate pass + exact fp64 re-rank)
    topk = None
    if args.topk_users > 0:
        try:
            Tq = min(args.topk_users, wl['n_users'])
            g = torch.Generator(device=device)
            g.manual_seed(SEED)
            users = torch.randperm(wl['n_users'], device=device, generator=g)[:Tq].to(torch.int32)   # Philox-chosen sample
            tk_ms, fb, cand, tk_clk = time_topk(eng, users, csr)
            topk = topk_object(tk_ms, Tq, wl['n_items'], wl['d'], fb, cand, pk,
                               metric='users/s full-catalog top-100 (mask train items), exact result via tensor-core candidate pass',
                               kernels='k_prep x2 + k_topk_tc (tcgen05.mma fp16 operands, fp32 TMEM accumulators, TMA) + k_rerank (fp64) '
                                       '+ k_topk_exact (fallback rows)',
                               model='the %s model trained above' % wl['model'], clocks=tk_clk)
        except Exception as e:
            topk = sub_error(e)
        if args.topk_c5_items > 0 and 'error' not in topk:
            # configs[4]'s catalogue on one GPU (item-sharded over P GPUs: see the N > 1 lines): BPRMF scoring, 10M items
            try:
                from collaborativefilteringusingtensorflow_b200.engine import FactorEngine
                eng._tc_ws = None
                torch.cuda.empty_cache()
                big = FactorEngine('bpr', Tq, args.topk_c5_items, 128, device, seed=7)
                big.accU = big.accV = None                      # scoring only
                uq = torch.arange(Tq, dtype=torch.int32, device=device)
                ms5, fb5, cand5, clk5 = time_topk(big, uq, None)
                topk['c5_catalogue'] = topk_object(ms5, Tq, args.topk_c5_items, 128, fb5, cand5, pk, clocks=clk5)
                del big
            except Exception as e:
                big = None
                topk['c5_catalogue'] = sub_error(e)
            torch.cuda.empty_cache()

    # ---- the other configurations of BASELINE.json on this GPU (same engine, same kernels)
    other = {}
    eng._tc_ws = None
    if not args.no_other_configs:
        Ko = max(3, min(K, args.other_steps))
        # BPRMF, W = 1 -- the metric's namesake -- on the configs[1] shape (same interactions)
        wb = WORKLOADS['c2-bpr'] if args.workload == 'c2' else dict(wl, model='bpr', W=1, G=0, hyper=dict(reg=0.1, lr=0.1), desc='BPRMF on the same shape, W=1')
        try:
            mb, sb, tb = time_training(wb, csr, B, Ko, Wm, device, args.optimizer, args.update, pk)
            mse, h2db, _ = time_e2e(mb, sb, B, Ko, device)
            if args.workload == 'c2':
                tb['roofline']['traffic'], tsrc_b = measured_traffic('c2-bpr', B, args.optimizer, args.update)
                if tsrc_b:
                    tb['roofline']['traffic_source'] = tsrc_b
            other['bpr_w1'] = dict(workload=wb['desc'], value=tb['units'] / (tb['ms'] * 1e-3), unit='triple updates/s', steps=Ko,
                                   ms_per_step=tb['ms'] / Ko, batch_pairs=B, roofline=tb['roofline'],
                                   e2e=dict(value=tb['units'] / (mse * 1e-3), unit='triple updates/s', ms_per_step=mse / Ko,
                                            h2d_bytes_per_step=h2db, d2h_bytes_per_step=8),
                                   loss_first_last=[float(tb['losses'][0]), float(tb['losses'][-1])])
        except Exception as e:
            other['bpr_w1'] = sub_error(e)
        mb = sb = tb = None
        torch.cuda.empty_cache()
        if args.workload == 'c2':
            try:
                w3 = WORKLOADS['c3']
                csr3 = synth_interactions(w3['n_users'], w3['n_items'], w3['nnz'], SEED, device)
                m3, s3, t3 = time_training(w3, csr3, B, Ko, Wm, device, args.optimizer, args.update, pk)
                t3['roofline']['note'] = 'tables + accumulators are 42 MB x 2: L2-resident, the HBM fraction is not a DRAM claim'
                t3['roofline']['traffic'], tsrc_3 = measured_traffic('c3', B, args.optimizer, args.update)
                if tsrc_3:
                    t3['roofline']['traffic_source'] = tsrc_3
                other['c3_gbpr'] = dict(workload=w3['desc'], value=t3['units'] / (t3['ms'] * 1e-3), unit='triple updates/s (pairs x W)',
                                        steps=Ko, ms_per_step=t3['ms'] / Ko, batch_pairs=B, nnz=csr3.nnz, roofline=t3['roofline'],
                                        loss_first_last=[float(t3['losses'][0]), float(t3['losses'][-1])])
            except Exception as e:
                other['c3_gbpr'] = sub_error(e)
            m3 = s3 = t3 = csr3 = None
            torch.cuda.empty_cache()
        try:
            other['c4_als_slice'] = als_slice(ALS_SLICE if args.workload == 'c2' else ALS_SMALL, device, pk)
        except Exception as e:
            other['c4_als_slice'] = sub_error(e)

    # ---- CPU baseline: the oracle port on this box's cores, bounded samples (best-effort and reference-faithful)
    cpub = cpuf = None
    if not args.no_cpu_baseline:
        try:
            csr_host = (csr.indptr.cpu().numpy(), csr.indices.cpu().numpy(), csr.rows.cpu().numpy())
            Bc = min(B, args.cpu_batch)
            n_cpu, t_cpu, cores = cpu_baseline(wl, Bc, csr_host, budget_s=args.cpu_budget)
            cpub = dict(value=n_cpu * Bc * wl['W'] / t_cpu, unit='triple updates/s', cores=cores, kind='port', setting='best-effort',
                        batch_pairs=Bc,
                        sample='%d minibatches of B=%d pairs x W=%d (torch-CPU oracle port of the TF1 step incl. '
                               'whole-table clip + vectorised numpy rejection sampler), %.1f s; host has %d cores'
                               % (n_cpu, Bc, wl['W'], t_cpu, os.cpu_count()))
            if wl['model'] in ('cml', 'bpr'):
                cpuf = cpu_baseline_faithful(wl, csr_host, budget_s=args.faithful_budget)
        except Exception as e:
            cpub = cpub or sub_error(e)
            cpuf = cpuf or (sub_error(e) if cpub.get('error') is None else None)

    out = dict(metric=metric_name(wl['d']), value=value, unit='triple updates/s', n_gpus=1, steps=K, warmup=max(Wm, 3),
               ms_per_step=ms / K, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
               config=same_config(wl, args), details=dict(nnz=csr.nnz, pairs_per_s=value / wl['W'], setup_s=setup_s,
                                                          sampling='on-device (k_sample_ranking), inside the timed region'),
               gpu_launches=tr['launches'], e2e=e2e, roofline=roofline, cpu_baseline=cpub, cpu_baseline_faithful=cpuf,
               clocks=clocks, topk=topk, other_configs=other, loss_first_last=[float(losses[0]), float(losses[-1])])
    print(json.dumps(out))


def sub_error(e):
    """A secondary measurement (top-K, the other configurations, the CPU baseline) that fails is reported in its own sub-object
    and must not cost the headline line, whose own measurement has finished by then and is never guarded."""
    try:
        import torch
        torch.cuda.empty_cache()
    except Exception:
        pass
    print('bench.py: sub-measurement failed: %s: %s' % (type(e).__name__, e), file=sys.stderr)
    return dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))


def als_slice(cfg, device, pk):
    """One user half-sweep of WRMF weighted ALS on a configs[3]-shaped slice, with the three terms BASELINE.md section 3
    asks for: tensor work (per-row Gram accumulation 2 nnz d^2 + global Gram 2 n d^2), Cholesky/solve (n (d^3/3 + 2 d^2),
    fp32 FMA pipe) and the gather of the observed rows (nnz * 4d bytes)."""
    import torch
    from collaborativefilteringusingtensorflow_b200 import WRMF
    nu, ni, nnz, d = cfg['n_users'], cfg['n_items'], cfg['nnz'], cfg['d']
    csr = synth_interactions(nu, ni, nnz, SEED, device)
    m = WRMF(nu, ni, weight=cfg['weight'], reg=cfg['reg'], n_factors=d, verbose=False, seed=1, solver='als', device=device)
    eng = m.engine
    eng.accU = eng.accV = None                               # the ALS solver has no optimizer state
    torch.cuda.empty_cache()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.als_half_sweep('users', csr)                         # warm-up (also sizes the workspace)
    torch.cuda.synchronize()
    e0.record()
    eng.als_half_sweep('users', csr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    phases = getattr(eng, 'als_phase_ms', None)
    nz = csr.nnz
    t_flops = 2.0 * nz * d * d + 2.0 * ni * d * d
    c_flops = nu * (d ** 3 / 3.0 + 2.0 * d * d)
    g_bytes = nz * 4.0 * d
    fma_peak = 148 * 128 * 2 * pk['sm_max'] * 1e6 / 1e12     # fp32 FMA pipe: 128 lanes x 2 flops x SMs x max clock, TFLOP/s
    s = ms * 1e-3
    out = dict(workload='configs[3] slice: WRMF weighted ALS user half-sweep, %d users (of 10M) x %d items, %d interactions, d=%d, '
                        'weight %.1f, reg %.1f' % (nu, ni, nz, d, cfg['weight'], cfg['reg']),
               value=nu / s, unit='rows solved/s', ms_per_half_sweep=ms, nnz=nz,
               terms=dict(tensor=dict(flops=t_flops, tflops=t_flops / s / 1e12, frac_of_tensor_peak=t_flops / s / 1e12 / pk['bf16']),
                          cholesky=dict(flops=c_flops, tflops=c_flops / s / 1e12, frac_of_fp32_fma_peak=c_flops / s / 1e12 / fma_peak,
                                        fp32_fma_peak_tflops=fma_peak),
                          gather=dict(bytes=g_bytes, GBs=g_bytes / s / 1e9, frac_of_hbm_peak=g_bytes / s / 1e9 / pk['hbm'])),
               note='the three terms are the ALGORITHMIC work of the direct form (SURVEY 8d: per-row Gram 2 nnz d^2, d^3/3 Cholesky per row, '
                    'nnz * 4d gathered bytes), each over the WHOLE half-sweep time; the solver itself does less: rows with n <= 128 observed '
                    'columns solve an n x n system in the whitened basis (DESIGN 4.6), so these are equivalent rates, not pipe utilisation',
               phases_ms=phases)
    del m, eng, csr
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    """The reference arm: the reference's CPU implementation of the path.  TensorFlow 1.x cannot be installed here (no
    network, no wheel), so this is the oracle port of the TF1 graph (oracle/steps_torch.py: torch CPU ops on every host
    core, tested equal to the numpy restatement oracle/steps.py) plus the numpy rejection sampler -- on the SAME metric,
    workload shape and synthetic generator as the GPU arm (generated on the host here)."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    B = min(args.batch, args.cpu_batch)
    t0 = time.time()
    csr_host = synth_host_csr(wl['n_users'], wl['n_items'], wl['nnz'], SEED)
    setup_s = time.time() - t0
    Wm, K = max(0, args.warmup), max(1, args.steps)
    # bounded: each step is one minibatch of B pairs; W warm-up + K timed steps, capped so the run ends within minutes
    n_w, t_w, cores = cpu_baseline(wl, B, csr_host, budget_s=0.0, max_steps=max(1, min(Wm, 2)), min_steps=max(1, min(Wm, 2)))
    steps_done, t, cores = cpu_baseline(wl, B, csr_host, budget_s=max(20.0, args.cpu_budget), max_steps=min(K, 12), min_steps=1)
    value = steps_done * B * wl['W'] / t
    cb = dict(value=value, unit='triple updates/s', cores=cores, kind='port', setting='best-effort', batch_pairs=B,
              sample='%d minibatches of B=%d pairs x W=%d, torch-CPU oracle port of the TF1 step on all host cores + vectorised '
                     'numpy rejection sampler (TensorFlow not installable); host has %d cores' % (steps_done, B, wl['W'], os.cpu_count()))
    faithful = cpu_baseline_faithful(wl, csr_host, budget_s=args.faithful_budget) if wl['model'] in ('cml', 'bpr') else None
    print(json.dumps(dict(impl='reference', metric=metric_name(wl['d']),
                          value=value, unit='triple updates/s', n_gpus=args.gpus, steps=steps_done, warmup=n_w,
                          ms_per_step=1e3 * t / steps_done, higher_is_better=True, scaling='weak', vs_baseline=None,
                          dtype='f32', data='synthetic', config=same_config(wl, args),
                          same_config=dict(workload=True, generator=True, tables='full size', nnz=int(len(csr_host[1])),
                                           batch_pairs_used=B, batch_pairs_of_config=args.batch,
                                           note='the CPU arm runs the same step on minibatches of %d pairs (a bounded sample: a 2^20-pair '
                                                'minibatch takes ~1 min per step on the host); throughput per triple is what is compared' % B),
                          details=dict(setup_s=setup_s), cpu_baseline=cb, cpu_baseline_faithful=faithful,
                          e2e=dict(value=value, unit='triple updates/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


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
    ap.add_argument('--topk-users', type=int, default=1_000_000, help='query users of the top-K measurement (SURVEY 8d: a fixed 1M-user sample)')
    ap.add_argument('--topk-c5-items', type=int, default=10_000_000, help='also time top-100 over a configs[4]-sized catalogue (0 = skip)')
    ap.add_argument('--cpu-budget', type=float, default=15.0)
    ap.add_argument('--faithful-budget', type=float, default=6.0)
    ap.add_argument('--other-steps', type=int, default=30, help='timed minibatches of the other_configs sub-runs')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--no-c5', action='store_true', help='N > 1: skip the configs[4] sub-run')
    ap.add_argument('--grow-catalogue', action='store_true', help='N > 1: n_items x N items in total instead of a fixed catalogue')
    ap.add_argument('--item-transport', default='auto', choices=['nccl', 'peer', 'peer-push', 'fetch', 'replicate', 'auto'],
                    help='N > 1: how item rows reach the step (NCCL all-to-all of unique rows / NVLink peer reads in the kernel)')
    ap.add_argument('--phases', action='store_true', help='N > 1: also print the per-minibatch timeline to stderr')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
