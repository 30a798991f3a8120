#!/usr/bin/env python
"""Times WRMF-ALS half-sweeps on a synthetic interaction matrix (users x items x nnz, d)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench   # noqa: E402
from collaborativefilteringusingtensorflow_b200 import WRMF   # noqa: E402

if __name__ == '__main__':
    nu, ni, nnz, d = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1_000_000, 500_000, 100_000_000, 128)))
    dev = torch.device('cuda', 0)
    csr = bench.synth_interactions(nu, ni, nnz, 1, dev)
    csr_t = csr.transpose()
    m = WRMF(nu, ni, weight=2.0, reg=0.1, n_factors=d, verbose=False, seed=1, solver='als', device=dev)
    m.engine.als_half_sweep('users', csr)
    torch.cuda.synchronize()
    for side, c, n in (('users', csr, nu), ('items', csr_t, ni)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.engine.als_half_sweep(side, c)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fl = 2.0 * csr.nnz * d * d / 2 + n * (d ** 3 / 3.0 + 2 * d * d)
        print('ALS %s half-sweep: %d rows, nnz %d, d %d: %.2f ms  (%.0f rows/s, %.1f TFLOP/s fp32-equivalent)' % (side, n, csr.nnz, d, ms, n / (ms * 1e-3), fl / (ms * 1e-3) / 1e12))
