#!/usr/bin/env python
"""Times the ranking sampler alone on the configs[1] interactions (one launch of n minibatches of 2^20 pairs, W negatives)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.WORKLOADS['c2']
dev = torch.device('cuda', 0)
csr = bench.synth_interactions(wl['n_users'], wl['n_items'], wl['nnz'], bench.SEED, dev)
for W in (5, 1):
    s = bench.make_sampler(dict(wl, W=W), csr, 1 << 20, bench.SEED, dev)
    s.next_chunk(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        s.next_chunk(4)
    e1.record()
    torch.cuda.synchronize()
    print('W=%d: %.3f ms per minibatch of 2^20 pairs' % (W, e0.elapsed_time(e1) / 20))
