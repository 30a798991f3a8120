import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.WORKLOADS['c2']
dev = torch.device('cuda', 0)
csr = bench.synth_interactions(wl['n_users'], wl['n_items'], wl['nnz'], 2026, dev)
model = bench.make_model(wl, dev)
sampler = bench.make_sampler(wl, csr, 1 << 20, 2026, dev)
for _ in range(2):
    model._train_arrays(sampler.next_chunk(50), 1 << 20)
eng = model.engine
T = 37888
users = torch.randperm(wl['n_users'], device=dev)[:T].to(torch.int32)
ti, tv = eng.topk(users, 100, csr, return_values=True, method='tensor')
st = eng.tc_stats.cpu().numpy()
print('fallback rows', st[0], 'cand/row', st[1] / max(1, T - st[0]), 'max 2eps', np.int32(st[2]).view(np.float32))
# recompute eps and find rows whose exact top-K is tightly packed
U = eng.U[users.long()]
un = U.norm(dim=1)
print('|u| quantiles', torch.quantile(un, torch.tensor([0., .1, .5, .9, 1.], device=dev)).tolist())
vn = eng.V.norm(dim=1)
print('|v| quantiles', torch.quantile(vn[:200000], torch.tensor([0., .1, .5, .9, 1.], device=dev)).tolist())
spread = (tv[:, 0] - tv[:, 99])
print('top1-top100 score spread quantiles', torch.quantile(spread.float(), torch.tensor([0., .01, .1, .5, .9], device=dev)).tolist())
# rows with tiny spread
tight = (spread < 0.02).sum().item()
print('rows with top-100 spread < 0.02:', tight)
deg = csr.row_lengths()[users.long()]
print('degree of users: mean %.1f, of tight rows: %.1f' % (deg.float().mean().item(), deg[spread < 0.02].float().mean().item() if tight else -1))
print('|u| of tight rows', un[spread < 0.02][:10].tolist())
