#!/usr/bin/env python
"""Times full-catalog masked top-K (tensor-core path vs exact path) on synthetic tables; prints users/s and TFLOP/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborativefilteringusingtensorflow_b200.engine import FactorEngine   # noqa: E402


def run(kind, T, N, d, K=100, method='tensor', reps=3):
    eng = FactorEngine(kind, T, N, d, 'cuda:0', seed=1, init_stddev=0.1)
    users = torch.arange(T, dtype=torch.int32, device='cuda:0')
    eng.topk(users[:512], K, None, method=method)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.topk(users, K, None, method=method)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fl = 2.0 * T * N * d
    st = eng.tc_stats.cpu().numpy() if method == 'tensor' else None
    if st is not None:
        import numpy as np
        print('      fallback rows %d / %d, candidates/row %.1f, max 2eps %.4g, bmax %.4g' % (st[0], T, st[1] / max(1, T - st[0]), np.int32(st[2]).view(np.float32), np.int32(st[3]).view(np.float32)))
    print('%-5s %-6s T=%7d N=%9d d=%3d K=%4d: %9.2f ms  %10.0f users/s  %7.1f TFLOP/s' % (kind, method, T, N, d, K, best, T / (best * 1e-3), fl / (best * 1e-3) / 1e12), flush=True)
    del eng
    torch.cuda.empty_cache()


if __name__ == '__main__':
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), reps=int(sys.argv[5]) if len(sys.argv) > 5 else 2,
            K=int(sys.argv[6]) if len(sys.argv) > 6 else 100)
        sys.exit(0)
    run('cml', 2048, 500_000, 128)
    run('cml', 37888, 500_000, 128)
    run('bpr', 37888, 500_000, 128)
    run('gbpr', 37888, 500_000, 64)
    run('bpr', 37888, 10_000_000, 128, reps=1)
