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


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: GBPR at the ML-20M shape (gbprmf.py + sampler_gbpr; 138 493 x 26 744, 20 M interactions,
# d = 64, G = 3, W = 5)
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def c3():
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    wl = bench.WORKLOADS['c3']
    dev = torch.device('cuda', 0)
    csr = bench.synth_interactions(wl['n_users'], wl['n_items'], wl['nnz'], 2026, dev)
    model = bench.make_model(wl, dev)
    sampler = bench.make_sampler(wl, csr, 1 << 20, 11, dev)
    return wl, csr, model, sampler


def test_gbpr_sampler_and_step_properties_at_configs2_size(c3):
    import torch
    wl, csr, model, sampler = c3
    eng = model.engine
    B, W, G = 1 << 20, wl['W'], wl['G']
    assert csr.shape == (138_493, 26_744) and csr.nnz > 19_000_000
    pairs, negs, group = sampler.next_chunk(2)
    assert pairs.shape == (2 * B, 2) and negs.shape == (2 * B, W) and group.shape == (2 * B, G)
    assert bool(_is_positive(csr, pairs[:, 0], pairs[:, 1]).all())
    uw = pairs[:, 0:1].expand(-1, W).reshape(-1)
    assert not bool(_is_positive(csr, uw, negs.reshape(-1)).any())                       # sampler_gbpr.py:35-36
    ig = pairs[:, 1:2].expand(-1, G).reshape(-1)
    assert bool(_is_positive(csr, group.reshape(-1), ig).all())                          # group members rated the item (sampler_gbpr.py:41)
    sampler.check_flags()
    st0 = {k: v.clone() for k, v in model.state_dict().items()}
    losses = model._train_arrays((pairs, negs, group), B)
    eng.check_flags()
    assert losses.shape == (2,) and bool(torch.isfinite(losses).all()) and float(losses.min()) > 0
    ws = eng._ws
    assert int(ws['metaU'].abs().sum()) == 0 and int(ws['metaV'].abs().sum()) == 0
    assert int((ws['slot_row'] != -1).sum()) == 0 and float(ws['staging'].abs().sum()) == 0.0
    st1 = model.state_dict()
    tu = torch.zeros(wl['n_users'], dtype=torch.bool, device=eng.device)
    tu[pairs[:, 0].long()] = True
    tu[group.reshape(-1).long()] = True
    tv = torch.zeros(wl['n_items'], dtype=torch.bool, device=eng.device)
    tv[pairs[:, 1].long()] = True
    tv[negs.reshape(-1).long()] = True
    for name, t in (('U', tu), ('accU', tu), ('V', tv), ('accV', tv), ('b', tv), ('accb', tv)):
        assert torch.equal(st1[name][~t], st0[name][~t]), name                           # untouched rows keep their bytes
    for name in ('accU', 'accV', 'accb'):
        assert bool((st1[name] >= st0[name]).all()), name                                # Adagrad accumulators only grow
    assert bool((st1['accU'][tu].sum(1) > st0['accU'][tu].sum(1)).all())
    assert bool(torch.isfinite(st1['U']).all()) and bool(torch.isfinite(st1['V']).all()) and bool(torch.isfinite(st1['b']).all())


def test_gbpr_step_equals_oracle_on_the_configs2_tables(c3):
    """One minibatch of 2^15 sampled pairs on the FULL-SIZE tables against the numpy restatement of gbprmf.py:58-106
    (the oracle finishes that in seconds; gathering only the touched rows keeps the comparison exact)."""
    import numpy as np
    import torch
    from oracle import steps
    wl, csr, model, sampler = c3
    B = 1 << 15
    pairs, negs, group = (t[:B].contiguous() for t in sampler.next_chunk(1))
    P = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    loss = float(model._train_arrays((pairs, negs, group), B)[0].item())
    model.engine.check_flags()
    h = wl['hyper']
    want = steps.gbpr_step(P['U'], P['V'], P['b'], P['accU'], P['accV'], P['accb'], pairs.cpu().numpy(), negs.cpu().numpy(),
                           group.cpu().numpy(), h['lr'], h['reg'], h['rho'])
    got = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    assert abs(loss - want) <= 1e-5 * abs(want)
    for k in ('U', 'V', 'b'):
        np.testing.assert_allclose(got[k], P[k], rtol=1e-5, atol=1e-6, err_msg=k)      # north_star: 1e-5 relative (fp32)
    for k in ('accU', 'accV', 'accb'):
        np.testing.assert_allclose(got[k], P[k], rtol=2e-5, atol=1e-6, err_msg=k)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4], one GPU's share at 8 GPUs: BPRMF, 12.5 M users x 10 M items, d = 128 (23 GB of tables)
# ------------------------------------------------------------------------------------------------------------------
def test_bpr_step_properties_on_a_configs4_share():
    import torch
    from collaborativefilteringusingtensorflow_b200 import BPRMF
    dev = torch.device('cuda', 0)
    nu, ni, d, B = 12_500_000, 10_000_000, 128, 1 << 20
    m = BPRMF(nu, ni, n_factors=d, reg=0.1, lr=0.1, batch_size=B, verbose=False, seed=3, device=dev)
    eng = m.engine
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    pairs = torch.stack([torch.randint(0, nu, (B,), device=dev, generator=g),
                         torch.randint(0, ni, (B,), device=dev, generator=g)], 1).to(torch.int32)
    pairs[:4096, 0] = pairs[0, 0]                                  # a hot user and a hot item: duplicated rows, summed once
    pairs[4096:8192, 1] = pairs[1, 1]
    negs = torch.randint(0, ni, (B, 1), device=dev, generator=g).to(torch.int32)
    tu = torch.zeros(nu, dtype=torch.bool, device=dev)
    tu[pairs[:, 0].long()] = True
    tv = torch.zeros(ni, dtype=torch.bool, device=dev)
    tv[pairs[:, 1].long()] = True
    tv[negs.reshape(-1).long()] = True
    # row checksums instead of 23 GB of clones
    cu0, cv0 = eng.U.double().sum(1), eng.V.double().sum(1)
    loss = m._train_arrays((pairs, negs), B)
    eng.check_flags()
    assert bool(torch.isfinite(loss).all()) and float(loss[0]) > 0
    cu1, cv1 = eng.U.double().sum(1), eng.V.double().sum(1)
    assert torch.equal(cu1[~tu], cu0[~tu]) and torch.equal(cv1[~tv], cv0[~tv])           # untouched rows keep their bytes
    assert bool((eng.accU[~tu] == 0.1).all()) and bool((eng.accV[~tv] == 0.1).all())
    assert bool((eng.accU[tu].sum(1) > 12.805).all()) and bool((eng.accV[tv].sum(1) > 12.805).all())   # every touched row applied (g^2 >= (reg * p)^2)
    assert bool((eng.accU >= 0.1).all()) and bool((eng.accV >= 0.1).all())
    ws = eng._ws
    assert int(ws['metaU'].abs().sum()) == 0 and int(ws['metaV'].abs().sum()) == 0       # workspace back at rest
    assert int((ws['slot_row'] != -1).sum()) == 0 and float(ws['staging'].abs().sum()) == 0.0
    # the hot user's row against its definition: all 4096 gradients at the pre-update rows, summed, ONE Adagrad apply
    # (bprmf.py:83-88); recomputed here in fp64 torch from the pre-update rows of a second, identically seeded model
    ref = BPRMF(nu, ni, n_factors=d, reg=0.1, lr=0.1, batch_size=B, verbose=False, seed=3, device=dev).engine
    u = int(pairs[0, 0])
    sel = (pairs[:, 0] == u).nonzero().reshape(-1)
    Uu = ref.U[u, :d].double()
    Vi, Vj = ref.V[pairs[sel, 1].long(), :d].double(), ref.V[negs[sel, 0].long(), :d].double()
    s = torch.sigmoid(Vi @ Uu - Vj @ Uu) - 1.0
    gsum = (s[:, None] * (Vi - Vj)).sum(0) + 0.1 * len(sel) * Uu
    acc = 0.1 + gsum * gsum
    want = Uu - 0.1 * gsum / acc.sqrt()
    assert torch.allclose(eng.U[u, :d].double(), want, rtol=1e-4, atol=1e-6)
    assert torch.allclose(eng.accU[u, :d].double(), acc, rtol=1e-4, atol=1e-6)
    del m, ref, eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------------
# Tensor top-K on the configs[4] catalogue (10 M items, d = 128): identical to the exact fp64 kernel
# ------------------------------------------------------------------------------------------------------------------
def test_tensor_topk_equals_exact_on_the_10m_catalogue():
    import torch
    from collaborativefilteringusingtensorflow_b200.engine import FactorEngine
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    dev = torch.device('cuda', 0)
    T, ni, d, K = 1024, 10_000_000, 128, 100
    eng = FactorEngine('bpr', T, ni, d, dev, seed=7)
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    # every query user has 64 training items to mask; plant each user's would-be top items among them so that the mask
    # matters: first find the unmasked top-8, then mask exactly those
    users = torch.arange(T, dtype=torch.int32, device=dev)
    top8 = eng.topk(users, 8, None, method='tensor')
    rnd = torch.randint(0, ni, (T, 56), device=dev, generator=g).to(torch.int32)
    cols = torch.cat([top8, rnd], 1)
    rows = torch.arange(T, device=dev, dtype=torch.int32).view(-1, 1).expand(-1, 64)
    key = torch.unique(rows.reshape(-1).long() * ni + cols.reshape(-1).long())
    mask = DeviceCSR.from_device_coo((key // ni).to(torch.int32), (key % ni).to(torch.int32), (T, ni))
    ti, tv = eng.topk(users, K, mask, return_values=True, method='tensor')
    assert int(eng.tc_stats[0].item()) == 0                                              # no row fell back to the exact kernel
    ei, ev = eng.topk(users[:256], K, mask, return_values=True, method='exact')
    assert torch.equal(ti[:256], ei) and torch.equal(tv[:256], ev)                       # bit-exact ids and fp64 scores
    assert bool((tv[:, :-1] >= tv[:, 1:]).all())
    assert not bool((ti[:, :, None] == top8[:, None, :]).any())                          # the planted training items are gone
    assert int(ti.min()) >= 0 and int(ti.max()) < ni
    del eng
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3] (WRMF ALS, 10 M x 1 M, d = 128, 500 M interactions): a slice with the same shape per row
# (1 M users of the 10 M, the whole 1 M-item catalogue, 50 interactions per user on average)
# ------------------------------------------------------------------------------------------------------------------
def _wals_objective(eng, csr, weight, reg):
    """sum_ui c_ui (r_ui - x_u.y_i)^2 + reg (|X|^2 + |Y|^2) over ALL pairs without the dense matrix:
    sum_all (x.y)^2 = <X^T X, Y^T Y>; observed pairs add weight (1 - s)^2 - s^2."""
    import torch
    d = eng.d
    X, Y = eng.U[:, :d], eng.V[:, :d]
    gx = torch.zeros(d, d, dtype=torch.float64, device=X.device)
    gy = torch.zeros(d, d, dtype=torch.float64, device=X.device)
    for t, gm in ((X, gx), (Y, gy)):
        for lo in range(0, t.shape[0], 1 << 18):
            blk = t[lo:lo + (1 << 18)].double()
            gm += blk.T @ blk
    tot = float((gx * gy).sum())
    obs = 0.0
    for lo in range(0, csr.nnz, 1 << 22):
        r, c = csr.rows[lo:lo + (1 << 22)].long(), csr.indices[lo:lo + (1 << 22)].long()
        s = (X[r].double() * Y[c].double()).sum(1)
        obs += float((weight * (1 - s) ** 2 - s * s).sum())
    return tot + obs + reg * float(torch.diagonal(gx).sum() + torch.diagonal(gy).sum())


def test_als_sweeps_do_not_increase_the_objective_on_a_configs3_slice():
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from collaborativefilteringusingtensorflow_b200 import WRMF
    dev = torch.device('cuda', 0)
    nu, ni, nnz, d, weight, reg = 1_000_000, 1_000_000, 50_000_000, 128, 2.0, 0.1
    csr = bench.synth_interactions(nu, ni, nnz, 2026, dev)
    m = WRMF(nu, ni, weight=weight, reg=reg, n_factors=d, verbose=False, seed=1, solver='als', device=dev)
    eng = m.engine
    csr_t = csr.transpose()
    obj = [_wals_objective(eng, csr, weight, reg)]
    for sweep in range(2):
        eng.als_half_sweep('users', csr)
        obj.append(_wals_objective(eng, csr, weight, reg))
        eng.als_half_sweep('items', csr_t)
        obj.append(_wals_objective(eng, csr, weight, reg))
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj, obj[1:])), obj                   # exact block minimisation: monotone
    assert obj[-1] < 0.9 * obj[0]
    assert bool(torch.isfinite(eng.U).all()) and bool(torch.isfinite(eng.V).all())
    # a row's normal equations hold: (G + (w-1) sum y y^T + reg I) x = w sum y  for a few users solved last sweep ... the
    # item half-sweep ran after them, so check ITEM rows (solved last) against a dense fp64 solve
    X, Y = eng.V[:, :d].double(), eng.U[:, :d].double()
    G = torch.zeros(d, d, dtype=torch.float64, device=dev)
    for lo in range(0, nu, 1 << 18):
        blk = Y[lo:lo + (1 << 18)]
        G += blk.T @ blk
    for i in (0, 12345, ni - 1):
        lo, hi = int(csr_t.indptr[i]), int(csr_t.indptr[i + 1])
        Yp = Y[csr_t.indices[lo:hi].long()]
        A = G + (weight - 1.0) * (Yp.T @ Yp) + reg * torch.eye(d, dtype=torch.float64, device=dev)
        want = torch.linalg.solve(A, weight * Yp.sum(0))
        assert torch.allclose(X[i], want, rtol=1e-2, atol=1e-3 * float(want.abs().max()) + 1e-7), i   # fp32 Cholesky vs fp64 solve
