"""Size-independent properties at BASELINE.json's full configs[1] size (CML, 1M users x 500k items, d=128, ~100M
interactions, B = 2^20 pairs, W = 5), where the oracle itself is too slow to run:
sampler validity, workspace hygiene, untouched rows, unit-norm clip, accumulator monotonicity, and tensor-core top-K ==
exact top-K with masked, sorted results."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c2():
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    wl = bench.WORKLOADS['c2']
    dev = torch.device('cuda', 0)
    csr = bench.synth_interactions(wl['n_users'], wl['n_items'], wl['nnz'], 2026, dev)
    model = bench.make_model(wl, dev)
    sampler = bench.make_sampler(wl, csr, 1 << 20, 7, dev)
    return wl, csr, model, sampler


def _is_positive(csr, users, items):
    """Vectorised membership test on the device: is (u, i) a training pair?"""
    import torch
    key = users.to(torch.int64) * csr.shape[1] + items.to(torch.int64)
    allkeys = csr.rows.to(torch.int64) * csr.shape[1] + csr.indices.to(torch.int64)      # sorted (CSR order)
    pos = torch.searchsorted(allkeys, key).clamp_(max=allkeys.numel() - 1)
    return allkeys[pos] == key


def test_sampler_at_full_size(c2):
    import torch
    wl, csr, model, sampler = c2
    assert csr.nnz > 95_000_000 and csr.shape == (1_000_000, 500_000)
    pairs, negs = sampler.next_chunk(2)
    assert pairs.shape == (2 << 20, 2) and negs.shape == (2 << 20, 5)
    assert bool(_is_positive(csr, pairs[:, 0], pairs[:, 1]).all())                      # positives are training pairs
    u5 = pairs[:, 0:1].expand(-1, 5).reshape(-1)
    assert not bool(_is_positive(csr, u5, negs.reshape(-1)).any())                      # negatives never are
    assert int(negs.min()) >= 0 and int(negs.max()) < 500_000
    key = pairs[:, 0].to(torch.int64) * 500_000 + pairs[:, 1].to(torch.int64)
    assert torch.unique(key).numel() == key.numel()                                      # an epoch visits a pair at most once
    sampler.seek(0, 0)
    p2, n2 = sampler.next_chunk(2)
    assert torch.equal(p2, pairs) and torch.equal(n2, negs)                              # same seed, same stream
    assert abs(float(negs.float().mean()) - 249_999.5) < 500                             # uniform over the catalogue
    sampler.check_flags()


def test_step_properties_at_full_size(c2):
    import torch
    wl, csr, model, sampler = c2
    eng = model.engine
    B = 1 << 20
    sampler.seek(0, 0)
    model._train_arrays(sampler.next_chunk(1), B)                                        # first step (+ one-time full clip)
    U0, V0, aU0, aV0 = eng.U.clone(), eng.V.clone(), eng.accU.clone(), eng.accV.clone()
    pairs, negs = sampler.next_chunk(2)
    losses = model._train_arrays((pairs, negs), B)
    eng.check_flags()
    assert losses.shape == (2,) and bool(torch.isfinite(losses).all()) and float(losses.min()) > 0
    ws = eng._ws
    assert int(ws['metaU'].abs().sum()) == 0 and int(ws['metaV'].abs().sum()) == 0       # workspace returned to rest
    assert int((ws['slot_row'] != -1).sum()) == 0 and float(ws['staging'].abs().sum()) == 0.0
    touched_u = torch.zeros(wl['n_users'], dtype=torch.bool, device=eng.device)
    touched_u[pairs[:, 0].long()] = True
    touched_v = torch.zeros(wl['n_items'], dtype=torch.bool, device=eng.device)
    touched_v[pairs[:, 1].long()] = True
    touched_v[negs.reshape(-1).long()] = True
    assert torch.equal(eng.U[~touched_u], U0[~touched_u]) and torch.equal(eng.accU[~touched_u], aU0[~touched_u])
    assert torch.equal(eng.V[~touched_v], V0[~touched_v])
    assert bool((eng.accU >= aU0).all()) and bool((eng.accV >= aV0).all())               # Adagrad accumulators only grow
    assert bool((eng.accU[touched_u].sum(1) > aU0[touched_u].sum(1)).all())              # every touched row was applied once
    assert float(eng.U.norm(dim=1).max()) <= 1.0 + 1e-5 and float(eng.V.norm(dim=1).max()) <= 1.0 + 1e-5   # cml.py:119-122
    assert bool((eng.U[:, 128:] == 0).all()) if eng.ld > 128 else True


def test_tensor_topk_equals_exact_at_full_catalogue(c2):
    import torch
    wl, csr, model, sampler = c2
    eng = model.engine
    users = torch.randperm(wl['n_users'], device=eng.device)[:2048].to(torch.int32)
    ti, tv = eng.topk(users, 100, csr, return_values=True, method='tensor')
    ei, ev = eng.topk(users[:192], 100, csr, return_values=True, method='exact')
    assert torch.equal(ti[:192], ei) and torch.equal(tv[:192], ev)
    assert bool((tv[:, :-1] >= tv[:, 1:]).all())                                         # sorted by score
    assert int(ti.min()) >= 0 and int(ti.max()) < wl['n_items']
    u100 = users.view(-1, 1).expand(-1, 100).reshape(-1)
    assert not bool(_is_positive(csr, u100, ti.reshape(-1)).any())                       # training items are masked
    assert int(eng.tc_stats[0].item()) <= 20                                             # the tensor path did the work
