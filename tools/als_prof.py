import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
cfg = dict(n_users=200_000, n_items=200_000, nnz=10_000_000, d=128, weight=2.0, reg=0.1)
if len(sys.argv) > 1 and sys.argv[1] == 'big':
    cfg = bench.ALS_SLICE
print(bench.als_slice(cfg, torch.device('cuda', 0), bench.peaks())['ms_per_half_sweep'], 'ms per half-sweep of', cfg['n_users'], 'rows')
