"""PRIGP / CPLR: the gradient-only tuple step + dense applies against the autograd golden
(tests/golden/tuple_golden.npz) and the numpy oracle; the two tuple samplers against the invariants of
sampler_prigp.py:22-52 / sampler_uitj_ranking.py:22-38; the coefficient preprocessing against the oracle; a short ml-100k run."""
import json
import os

import numpy as np
import pytest

from oracle import neighbors as onb
from oracle import steps

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _model(name, nu, ni, d, h, seed=1):
    from collaborativefilteringusingtensorflow_b200 import CPLR, PRIGP
    if name.startswith('prigp'):
        return PRIGP(nu, ni, alpha=h['alpha'], reg=h['reg'], n_factors=d, lr=h['lr'], verbose=False, seed=seed)
    return CPLR(nu, ni, alpha=h['alpha'], beta=h['beta'], gamma=h['gamma'], reg=h['reg'], n_factors=d, lr=h['lr'], verbose=False, seed=seed)


@pytest.mark.parametrize('source', ['autograd', 'refgraph'])
@pytest.mark.parametrize('name', ['prigp', 'prigp_d20', 'cplr', 'cplr_d20'])
def test_tuple_steps_equal_the_autograd_golden(name, source):
    """source 'refgraph': the expected values come from the reference's own prigp.py / cplr_u.py graphs (TF1 stand-in)."""
    import refgraph_cases
    tg = refgraph_cases.golden('tuple', source)
    h = json.loads(str(tg[name + '/hyper']))
    init = {k: tg['%s/init/%s' % (name, k)] for k in ('U', 'V', 'b')}
    m = _model(name, init['U'].shape[0], init['V'].shape[0], init['U'].shape[1], h)
    m.load_state_dict(init)
    for s in range(2):
        t, c = tg['%s/batch%d/tuples' % (name, s)], tg['%s/batch%d/coefs' % (name, s)]
        loss = m.step(t) if name.startswith('prigp') else m.step(t, c)
        assert abs(loss - float(tg['%s/loss%d' % (name, s)])) <= 2e-5 * abs(loss)
        sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
        for k in ('U', 'V', 'b'):
            np.testing.assert_allclose(sd[k], tg['%s/step%d/%s' % (name, s, k)], rtol=1e-5, atol=1e-6, err_msg='%s %d %s' % (name, s, k))
            key = '%s/step%d/acc%s' % (name, s, k)
            if key in tg.files:
                np.testing.assert_allclose(sd['acc' + k], tg[key], rtol=2e-5, atol=1e-6, err_msg=key)
    if name.startswith('prigp'):
        assert np.array_equal(sd['b'], init['b']) and np.all(sd['accb'] == np.float32(0.1))     # prigp.py:134


@pytest.mark.parametrize('kind,nu,ni,d,B', [('prigp', 40, 30, 7, 64), ('prigp', 500, 800, 128, 1000), ('cplr', 33, 70, 50, 256), ('cplr', 600, 300, 200, 1000)])
def test_tuple_steps_equal_the_oracle(kind, nu, ni, d, B):
    rng = np.random.default_rng(nu + d)
    h = dict(lr=0.1, reg=0.05, alpha=3.0, beta=0.7, gamma=1.5)
    m = _model(kind, nu, ni, d, h, seed=3)
    P = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    for s in range(3):
        width = 5 if kind == 'prigp' else 4
        t = np.concatenate([rng.integers(0, nu, (B, 1)), rng.integers(0, ni, (B, width - 1))], axis=1).astype(np.int32)
        c = (rng.random((B, 2)) * 4).astype(np.float32)
        if kind == 'prigp':
            got = m.step(t)
            want = steps.prigp_step(P['U'], P['V'], P['b'], P['accU'], P['accV'], t, h['lr'], h['reg'], h['alpha'])
        else:
            got = m.step(t, c)
            want = steps.cplr_step(P['U'], P['V'], P['b'], P['accU'], P['accV'], P['accb'], t, c, h['lr'], h['reg'], h['alpha'], h['beta'], h['gamma'])
        assert abs(got - want) <= 2e-5 * abs(want)
        sd = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
        for k in P:
            scale = max(1.0, float(np.abs(P[k]).max()))
            np.testing.assert_allclose(sd[k], P[k], rtol=2e-5, atol=2e-6 * scale, err_msg='%s step %d %s' % (kind, s, k))


def test_out_of_range_tuple_ids_are_flagged():
    m = _model('prigp', 20, 20, 8, dict(lr=0.1, reg=0.01, alpha=1.0))
    t = np.zeros((4, 5), np.int32)
    t[2, 3] = 20
    before = {k: v.clone() for k, v in m.state_dict().items()}
    with pytest.raises(RuntimeError, match='out of range'):
        m.step(t)
    after = m.state_dict()
    assert all((before[k] - after[k]).abs().max() < 1.0 for k in before)


def _random_problem(rng, nu, ni, deg):
    from scipy.sparse import lil_matrix
    R = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        k = int(rng.integers(0, deg + 1))
        if k:
            R[u, rng.choice(ni, size=k, replace=False)] = 1
    return R


@pytest.mark.parametrize('weighted', [False, True])
def test_coefficient_matrix_equals_the_oracle(ml100k, weighted):
    """prigp.py:60-90 (neighbour counts) / cplr_u.py:64-97 (similarity-weighted sums) over the top-K most similar users."""
    from collaborativefilteringusingtensorflow_b200 import CPLR, PRIGP
    tra = ml100k['tra']
    m = (CPLR if weighted else PRIGP)(943, 1682, topK=5, n_factors=8, verbose=False, seed=1)
    got = m.coefficient_matrix(tra).cpu().numpy()
    idx, val = onb.topk_neighbors(onb.cosine_sim(tra.tocsr()), 5)
    dense = np.zeros((943, 943), np.float32)
    r, c = np.nonzero(idx >= 0)
    dense[r, idx[r, c]] = val[r, c]
    want = onb.coef_matrix(tra, dense, weighted)
    if weighted:
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9)       # float32 similarities summed in float64: same values,
    else:                                                                 # the oracle's matmul rounds differently in the last bits
        assert np.array_equal(got, want)


def test_prigp_sampler_invariants(ml100k):
    import torch
    from collaborativefilteringusingtensorflow_b200 import PRIGP
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_prigp import Sampler
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    tra = ml100k['tra']
    m = PRIGP(943, 1682, topK=5, n_factors=8, verbose=False, seed=1)
    coef = m._coef_to_csr(m.coefficient_matrix(tra))
    csr = DeviceCSR.from_scipy(tra, 'cuda:0')
    B = 1000
    s = Sampler(csr, coef, B, seed=5)
    n = s.batches_per_epoch
    assert n == int(tra.nnz / B)
    t = s.next_chunk(n)[0].cpu().numpy()
    s.check_flags()
    cd = np.zeros((943, 1682))
    cr, cc, cv = coef.rows.cpu().numpy(), coef.indices.cpu().numpy(), coef.values.cpu().numpy()
    cd[cr, cc] = cv
    pos = np.asarray(tra.todense()) > 0
    u, i, j, tt, k = t.T
    assert pos[u, i].all() and not pos[u, j].any()                                  # (u, i) positive, j a non-positive
    key = u.astype(np.int64) * 1682 + i
    assert len(np.unique(key)) == len(key)                                           # an epoch visits a pair at most once
    has = (cd[u] != 0).any(1)
    assert (tt[~has] == i[~has]).all() and (k[~has] == j[~has]).all()                # no coefficients: t = i, k = j (:36)
    assert (cd[u[has], tt[has]] != 0).all()                                          # t from the coefficient row
    inside = cd[u[has], k[has]] != 0
    assert (cd[u[has], tt[has]][inside] > cd[u[has], k[has]][inside]).all()          # k inside the row: a smaller coefficient (:48-49)
    assert 0.3 < inside.mean() < 0.9                                                 # Phi(nnz / n_items) of the varied rows, about a half
    s.seek(0, 0)
    assert np.array_equal(s.next_chunk(n)[0].cpu().numpy(), t)                       # same seed, same stream
    s2 = Sampler(csr, coef, B, seed=6)
    assert not np.array_equal(s2.next_chunk(2)[0].cpu().numpy(), t[:2 * B])
    b = Sampler(csr, coef, 100, seed=5).next_batch()
    assert b.dtype == np.int64 and b.shape == (100, 5)                               # the reference's `dtype=int` batch
    assert torch.cuda.is_available()


def test_uitj_sampler_invariants(ml100k):
    from collaborativefilteringusingtensorflow_b200 import CPLR
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_uitj_ranking import Sampler
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    tra = ml100k['tra']
    m = CPLR(943, 1682, topK=5, n_factors=8, verbose=False, seed=1)
    coef = m._coef_to_csr(m._normalise(m.coefficient_matrix(tra)))
    csr = DeviceCSR.from_scipy(tra, 'cuda:0')
    s = Sampler(csr, coef, 1000, seed=9)
    t, c = (x.cpu().numpy() for x in s.next_chunk(20))
    s.check_flags()
    cd = np.zeros((943, 1682), np.float32)
    cd[coef.rows.cpu().numpy(), coef.indices.cpu().numpy()] = coef.values.cpu().numpy()
    pos = np.asarray(tra.todense()) > 0
    u, i, tt, j = t.T
    assert pos[u, i].all()                                                           # i a positive of u
    assert (cd[u, tt] != 0).all() and not pos[u, tt].any()                           # t collaborative: a coefficient, not a positive
    assert not pos[u, j].any() and (cd[u, j] == 0).all()                             # j in neither set
    assert np.array_equal(c[:, 0], cd[u, i]) and np.array_equal(c[:, 1], cd[u, tt])  # coefs = (coef[u, i], coef[u, t])
    assert len(np.unique(u)) > 500                                                   # users drawn uniformly
    uitj, coefs = Sampler(csr, coef, 50, seed=9).next_batch()
    assert uitj.dtype == np.int64 and uitj.shape == (50, 4) and coefs.shape == (50, 2) and coefs.dtype == np.float64


@pytest.mark.parametrize('cls', ['PRIGP', 'CPLR'])
def test_training_on_ml100k_learns(ml100k, cls):
    import collaborativefilteringusingtensorflow_b200 as pkg
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    if cls == 'PRIGP':      # testprigp.py:21-31: topK 5, alpha 10, reg 0.1, 100 factors, batches of 1000
        m = pkg.PRIGP(943, 1682, 5, 10, 'cv', names, 10, .1, 100, 1000, 6, verbose=False, seed=1)
    else:
        m = pkg.CPLR(943, 1682, 5, 10, 'cv', names, 1., 1., 1., .05, 100, 1000, 6, verbose=False, seed=1)
    scores = m.train(1, ml100k['tra'], ml100k['tst'])
    assert scores[names.index('ndcg')] > 0.35, scores          # PopRank reaches 0.34 on this fold; BPRMF about 0.5
    m.close()

