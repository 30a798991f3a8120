#!/usr/bin/env python
"""Small single-process workload for `compute-sanitizer --tool memcheck|racecheck` (one tool per gpurun call): the exchange
kernels and the peer variants of the step kernel through a world-1 DistributedTrainer (the "peers" are this process's own
buffers: same kernels, same index arithmetic), the tensor top-K (tcgen05 / TMA / mbarrier ring), the neighbourhood, tuple
and ALS kernels.  Sizes are tiny: the sanitizer slows kernels down 50-100x."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, CPLR, PRIGP, WRMF, neighbors   # noqa: E402
from collaborativefilteringusingtensorflow_b200.dist import DistributedTrainer                       # noqa: E402
from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR                              # noqa: E402
from scipy.sparse import lil_matrix                                                                   # noqa: E402

rng = np.random.default_rng(0)
nu, ni, d, B, W = 300, 257, 128, 512, 3


class NoSampler(object):
    batch_size = B


for kind in ('bpr', 'cml'):
    for transport in ('fetch', 'peer', 'peer-push', 'nccl'):
        m = (BPRMF(nu, ni, n_factors=d, reg=0.05, verbose=False, seed=4) if kind == 'bpr' else
             CML(nu, ni, n_factors=d, reg_cov=1.0, margin=1.0, verbose=False, seed=4))
        tr = DistributedTrainer(m, NoSampler(), ni, 1, 0, item_transport=transport)
        for s in range(2):
            pairs = torch.from_numpy(np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B)], 1).astype(np.int32)).cuda()
            negs = torch.from_numpy(rng.integers(0, ni, (B, W)).astype(np.int32)).cuda()
            tr.step_chunk(pairs, negs, B)
        m.engine.check_flags()
        tr.close()
        print('exchange', kind, transport, 'ok', flush=True)

R = lil_matrix((nu, ni), dtype=np.float32)
for u in range(nu):
    R[u, rng.choice(ni, size=int(rng.integers(1, 70)), replace=False)] = 1
csr = DeviceCSR.from_scipy(R, 'cuda:0', with_values=True)
m = BPRMF(nu, 3000, n_factors=d, verbose=False, seed=1)
big = lil_matrix((nu, 3000), dtype=np.float32)
for u in range(nu):
    big[u, rng.choice(3000, size=20, replace=False)] = 1
users = torch.arange(nu, dtype=torch.int32).cuda()
ti = m.engine.topk(users, 100, DeviceCSR.from_scipy(big, 'cuda:0'), method='tensor')
ei = m.engine.topk(users, 100, DeviceCSR.from_scipy(big, 'cuda:0'), method='exact')
assert torch.equal(ti, ei)
print('tensor top-K ok', flush=True)
idx, sim = neighbors.cosine_topk(csr, 10)
sc = neighbors.neighbor_scores(csr, users, idx, sim, 'user')
neighbors.topk_dense(sc, 10, users, csr)
print('neighbours ok', flush=True)
p = PRIGP(nu, ni, n_factors=32, verbose=False, seed=1)
p.step(np.concatenate([rng.integers(0, nu, (64, 1)), rng.integers(0, ni, (64, 4))], 1).astype(np.int32))
c = CPLR(nu, ni, n_factors=32, verbose=False, seed=1)
c.step(np.concatenate([rng.integers(0, nu, (64, 1)), rng.integers(0, ni, (64, 3))], 1).astype(np.int32), rng.random((64, 2)).astype(np.float32))
print('tuple steps ok', flush=True)
w = WRMF(nu, ni, weight=3.0, reg=0.2, n_factors=64, verbose=False, seed=4, solver='als')
w.engine.als_half_sweep('users', csr)
w.engine.als_half_sweep('items', csr.transpose())
torch.cuda.synchronize()
print('ALS ok', flush=True)
print('SANITIZE TARGET DONE')
