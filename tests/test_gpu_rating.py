"""Rating path on the GPU: cf_predict_pairs and cf_rating_metrics against the oracle and the golden
vectors of the reference's metrics/rating.py; the MF model (mf.py) step by step against the oracle and, end to end on
ml-100k fold 1 with the reference driver's hyper-parameters (basic/testmf.py), against the oracle's recorded epochs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import rating as orc
from oracle import steps

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def golden():
    return json.load(open(os.path.join(GOLDEN, 'rating_golden.json')))


def test_rating_metrics_match_reference_golden(golden):
    import torch
    from collaborativefilteringusingtensorflow_b200.metrics import rating
    names = ['mae', 'mse', 'rmse', 'nope']
    for c in golden['cases']:
        yt, yp = np.array(c['ys_true']), np.array(c['ys_pred'])
        got = rating.evaluate(yt, yp, names)
        assert got[3] is None
        for k in range(3):
            assert got[k] == pytest.approx(c['scores'][names[k]], rel=1e-12)
        assert rating.mean_absolute_error(yt, yp) == pytest.approx(c['mae'], rel=1e-12)
        assert rating.mean_squared_error(yt, yp) == pytest.approx(c['mse'], rel=1e-12)
        assert rating.root_mean_squared_error(yt, yp) == pytest.approx(c['rmse'], rel=1e-12)
        # float32 predictions already on the device (what MF feeds) give the same numbers
        got32 = rating.evaluate(yt, torch.from_numpy(yp.astype(np.float32)).cuda(), names[:3])
        want32 = orc.evaluate(yt, yp.astype(np.float32), names[:3])
        assert got32 == pytest.approx(want32, rel=1e-12)
    with pytest.raises(ZeroDivisionError):
        rating.mean_absolute_error(np.zeros(0), np.zeros(0))
    with pytest.raises(ValueError):
        rating.evaluate(np.zeros(3), np.zeros(4), ['mae'])


def test_rating_metrics_clip_and_large_input():
    from collaborativefilteringusingtensorflow_b200.metrics import rating
    rng = np.random.default_rng(5)
    n = 1_000_003
    t = rng.integers(1, 6, n).astype(np.float64)
    p = (t + rng.normal(0, 2.0, n)).astype(np.float32)
    got = rating.evaluate(t, p, ['rmse', 'mae', 'mse'], clip=(1, 5))
    want = orc.evaluate(t, np.clip(p, np.float32(1), np.float32(5)), ['rmse', 'mae', 'mse'])
    assert got == pytest.approx(want, rel=1e-11)


@pytest.mark.parametrize('kind,d', [('mf', 100), ('mf', 7), ('bpr', 128), ('cml', 50), ('gbpr', 64)])
def test_predict_pairs_equals_dense_scores(kind, d):
    import torch
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF, MF
    nu, ni = 300, 400
    cls = dict(mf=MF, bpr=BPRMF, cml=CML, gbpr=GBPRMF)[kind]
    m = cls(nu, ni, n_factors=d, verbose=False, seed=2)
    rng = np.random.default_rng(d)
    pairs = np.stack([rng.integers(0, nu, 5000), rng.integers(0, ni, 5000)], 1).astype(np.int32)
    got = m.engine.predict_pairs(pairs).cpu().numpy()
    dense = m.engine.scores(torch.arange(nu, dtype=torch.int32)).cpu().numpy()          # fp64 [nu, ni]
    np.testing.assert_array_equal(got, dense[pairs[:, 0], pairs[:, 1]].astype(np.float32))
    if kind == 'mf':
        st = m.state_dict()
        want = orc.mf_predict(st['U'].cpu().numpy(), st['V'].cpu().numpy(), pairs)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-8)
    assert m.engine.predict_pairs(np.zeros((0, 2), np.int32)).numel() == 0
    bad = pairs[:4].copy()
    bad[2, 1] = ni
    out = m.engine.predict_pairs(bad)
    with pytest.raises(RuntimeError, match='out of range'):
        m.engine.check_flags()
    assert bool(torch.isnan(out[2]).item())


def test_mf_step_matches_oracle():
    from collaborativefilteringusingtensorflow_b200 import MF
    nu, ni, d, B = 943, 1682, 100, 1000
    m = MF(nu, ni, reg=0.1, n_factors=d, batch_size=B, verbose=False, seed=4)
    P = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    rng = np.random.default_rng(9)
    for s in range(3):
        uir = np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B), rng.integers(1, 6, B)], 1).astype(np.float64)
        loss = m.step(uir)
        ol = steps.wrmf_step(P['U'], P['V'], P['accU'], P['accV'], uir, 0.1, 0.1, 1.0)
        st = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
        for k in ('U', 'V', 'accU', 'accV'):
            np.testing.assert_allclose(st[k], P[k], rtol=5e-5, atol=2e-6, err_msg='mf step %d %s' % (s, k))
        assert abs(loss - ol) < 1e-5 * abs(ol)


def test_mf_trains_ml100k_like_the_oracle(golden, capsys):
    """basic/testmf.py on fold 1: reg .1, range (1, 5), 100 factors, batches of 1000 in file order.  Given the initial
    tables the run has no randomness, so every epoch's loss and test RMSE / MAE / MSE must follow the numpy oracle."""
    from scipy.sparse import coo_matrix
    from collaborativefilteringusingtensorflow_b200 import MF
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_rating import Sampler
    g = golden['mf_ml100k']
    z = np.load(os.path.join(GOLDEN, 'ml100k_fold1.npz'))
    tra = np.stack([z['tra_u'], z['tra_i'], z['tra_r']], 1).astype(np.float64)
    tst = np.stack([z['tst_u'], z['tst_i'], z['tst_r']], 1).astype(np.float64)
    nu, ni = 943, 1682
    trasR = coo_matrix((tra[:, 2].astype(np.float32), (tra[:, 0].astype(np.int64), tra[:, 1].astype(np.int64))), shape=(nu, ni)).tolil()
    init = np.random.default_rng(g['init_seed'])
    U0, V0 = steps.truncated_normal(init, (nu, g['n_factors'])), steps.truncated_normal(init, (ni, g['n_factors']))
    mf = MF(nu, ni, ['rmse', 'mae', 'mse'], tuple(g['range_of_ratings']), g['reg'], g['n_factors'], g['batch_size'],
            max_iter=len(g['epochs']), verbose=True, seed=1)
    mf.load_state_dict(dict(U=U0, V=V0, accU=np.full_like(U0, 0.1), accV=np.full_like(V0, 0.1)))
    sampler = Sampler(trasR=trasR, negRatio=.0, batch_size=g['batch_size'])
    scores = mf.train(1, tra, tst, sampler)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('fold=1 iter=')]
    assert len(lines) == len(g['epochs'])
    for line, e in zip(lines, g['epochs']):
        assert 'TraLoss=%.4f' % e['loss'] in line or abs(float(line.split('TraLoss=')[1].split()[0]) - e['loss']) < 2e-3 * e['loss']
        got = {kv.split('=')[0]: float(kv.split('=')[1]) for kv in line.split('Tst:')[1].split()}
        assert got['rmse'] == pytest.approx(e['rmse'], abs=2e-3) and got['mae'] == pytest.approx(e['mae'], abs=2e-3)
    last = g['epochs'][-1]
    assert scores == pytest.approx([last['rmse'], last['mae'], last['mse']], rel=2e-3)
    assert scores[0] < 1.0
    mf.close()


@pytest.mark.parametrize('source', ['autograd', 'refgraph'])
@pytest.mark.parametrize('name', ['svd', 'svd_d7'])
def test_svd_step_golden(name, source):
    """SVD (svd.py:52-80) on the GPU against the torch-autograd golden and against the reference's own svd.py run through
    its train() on the TF1 stand-in ('refgraph'): tables, kernel matrix, accumulators, loss."""
    from collaborativefilteringusingtensorflow_b200 import SVD
    import refgraph_cases
    z = refgraph_cases.golden('svd', source)
    U0 = z[name + '/init/U']
    nu, d = U0.shape
    ni = z[name + '/init/V'].shape[0]
    m = SVD(nu, ni, reg=0.05, n_factors=d, verbose=False, seed=3)
    m.load_state_dict(dict(U=U0, V=z[name + '/init/V'], K=z[name + '/init/K'], accU=np.full_like(U0, 0.1),
                           accV=np.full((ni, d), 0.1, np.float32), accK=np.full((d, d), 0.1, np.float32)))
    for s in range(2):
        loss = m.step(z['%s/batch%d' % (name, s)])
        assert loss == pytest.approx(float(z['%s/loss%d' % (name, s)]), rel=1e-5)
        st = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
        for k in ('U', 'V', 'K'):
            np.testing.assert_allclose(st[k], z['%s/step%d/%s' % (name, s, k)], rtol=1e-5, atol=1e-6, err_msg='%s step %d %s' % (name, s, k))
            np.testing.assert_allclose(st['acc' + k], z['%s/step%d/acc%s' % (name, s, k)], rtol=5e-5, atol=1e-6, err_msg='%s step %d acc%s' % (name, s, k))
    assert float(m._gU.abs().max()) == 0.0 and float(m._gV.abs().max()) == 0.0 and float(m._gK.abs().max()) == 0.0


def test_svd_predict_and_multi_batch_vs_oracle():
    import torch
    from collaborativefilteringusingtensorflow_b200 import SVD
    nu, ni, d, B = 300, 200, 32, 100
    m = SVD(nu, ni, reg=0.1, n_factors=d, batch_size=B, verbose=False, seed=6)
    P = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    rng = np.random.default_rng(12)
    pairs = np.stack([rng.integers(0, nu, 3000), rng.integers(0, ni, 3000)], 1).astype(np.int32)
    np.testing.assert_allclose(m.predict_pairs(pairs), orc.svd_predict(P['U'], P['V'], P['K'], pairs), rtol=1e-6, atol=1e-8)
    # three minibatches in one call (the epoch loop's chunk) == three oracle steps
    uir = np.stack([rng.integers(0, nu, 3 * B), rng.integers(0, ni, 3 * B), rng.integers(1, 6, 3 * B)], 1).astype(np.float64)
    losses = m._train_arrays((uir,), B).cpu().numpy()
    m.engine.check_flags()
    for k in range(3):
        ol = steps.svd_step(P['U'], P['V'], P['K'], P['accU'], P['accV'], P['accK'], uir[k * B:(k + 1) * B], 0.1, 0.1)
        assert losses[k] == pytest.approx(ol, rel=1e-5)
    st = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    for k in ('U', 'V', 'K', 'accU', 'accV', 'accK'):
        np.testing.assert_allclose(st[k], P[k], rtol=5e-5, atol=2e-6, err_msg=k)
    bad = pairs[:4].copy()
    bad[1, 0] = nu
    with pytest.raises(RuntimeError, match='out of range'):
        m.predict_pairs(bad)
    with pytest.raises(ValueError):
        SVD(10, 10, n_factors=129)


def test_svd_trains_ml100k_like_the_oracle(capsys):
    """basic/testsvd.py on fold 1 (32 factors, batches of 100, reg .1): deterministic given the initial tables."""
    from scipy.sparse import coo_matrix
    from collaborativefilteringusingtensorflow_b200 import SVD
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_rating import Sampler
    g = json.load(open(os.path.join(GOLDEN, 'svd_ml100k_golden.json')))
    z = np.load(os.path.join(GOLDEN, 'ml100k_fold1.npz'))
    tra = np.stack([z['tra_u'], z['tra_i'], z['tra_r']], 1).astype(np.float64)
    tst = np.stack([z['tst_u'], z['tst_i'], z['tst_r']], 1).astype(np.float64)
    nu, ni, k = 943, 1682, g['n_factors']
    trasR = coo_matrix((tra[:, 2].astype(np.float32), (tra[:, 0].astype(np.int64), tra[:, 1].astype(np.int64))), shape=(nu, ni)).tolil()
    init = np.random.default_rng(g['init_seed'])
    U0, V0, K0 = steps.truncated_normal(init, (nu, k)), steps.truncated_normal(init, (ni, k)), steps.truncated_normal(init, (k, k))
    m = SVD(nu, ni, ['rmse', 'mae', 'mse'], tuple(g['range_of_ratings']), g['reg'], k, g['batch_size'],
            max_iter=len(g['epochs']), verbose=True, seed=1)
    m.load_state_dict(dict(U=U0, V=V0, K=K0, accU=np.full_like(U0, 0.1), accV=np.full_like(V0, 0.1), accK=np.full_like(K0, 0.1)))
    scores = m.train(1, tra, tst, Sampler(trasR=trasR, negRatio=.0, batch_size=g['batch_size']))
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('fold=1 iter=')]
    assert len(lines) == len(g['epochs'])
    for line, e in zip(lines, g['epochs']):
        got = {kv.split('=')[0]: float(kv.split('=')[1]) for kv in line.split('Tst:')[1].split()}
        assert got['rmse'] == pytest.approx(e['rmse'], abs=3e-3) and got['mae'] == pytest.approx(e['mae'], abs=3e-3)
        assert float(line.split('TraLoss=')[1].split()[0]) == pytest.approx(e['loss'], rel=3e-3)
    last = g['epochs'][-1]
    assert scores == pytest.approx([last['rmse'], last['mae'], last['mse']], rel=3e-3)
    m.close()
