"""The rating-path oracle (oracle/rating.py) against the golden vectors captured from the reference's own
src/metrics/rating.py (tests/golden/rating_golden.json, made by oracle/gen_golden.py rating)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import rating as orc


@pytest.fixture(scope='module')
def golden():
    return json.load(open(os.path.join(GOLDEN, 'rating_golden.json')))


def test_metrics_match_the_reference_outputs(golden):
    names = ['mae', 'mse', 'rmse', 'nope']
    assert len(golden['cases']) >= 5
    for c in golden['cases']:
        yt, yp = np.array(c['ys_true']), np.array(c['ys_pred'])
        got = orc.evaluate(yt, yp, names)
        assert got[3] is None and c['scores']['nope'] is None
        for k in range(3):
            assert got[k] == pytest.approx(c['scores'][names[k]], rel=1e-14, abs=0)
        assert orc.mean_absolute_error(yt, yp) == pytest.approx(c['mae'], rel=1e-14)
        assert orc.mean_squared_error(yt, yp) == pytest.approx(c['mse'], rel=1e-14)
        assert orc.root_mean_squared_error(yt, yp) == pytest.approx(c['rmse'], rel=1e-14)


def test_toy_input_of_the_reference_main(golden):
    c = golden['cases'][0]                       # rating.py:33
    assert c['ys_true'] == [2.5, 1.5, 0] and c['ys_pred'] == [1, 2, 1]
    assert c['mae'] == pytest.approx(1.0) and c['mse'] == pytest.approx(3.5 / 3) and c['rmse'] == pytest.approx(np.sqrt(3.5 / 3))


def test_empty_input_divides_by_zero_like_the_reference():
    with pytest.raises(ZeroDivisionError):
        orc.mean_absolute_error(np.zeros(0), np.zeros(0))


def test_mf_step_is_wrmf_with_unit_weight_and_descends(golden):
    from oracle import steps
    rng = np.random.default_rng(3)
    nu, ni, d, B = 50, 40, 8, 64
    U, V = steps.truncated_normal(rng, (nu, d)), steps.truncated_normal(rng, (ni, d))
    accU, accV = np.full_like(U, 0.1), np.full_like(V, 0.1)
    tra = np.stack([rng.integers(0, nu, 4 * B), rng.integers(0, ni, 4 * B), rng.integers(1, 6, 4 * B)], 1).astype(np.float64)
    hist = orc.mf_train(U, V, accU, accV, tra, tra[:100], ['rmse'], (1, 5), 0.02, B, 6)
    assert hist[-1][0] < hist[0][0] and hist[-1][1][0] < hist[0][1][0]
    ep = golden['mf_ml100k']['epochs']
    assert [round(e['rmse'], 4) for e in ep] == sorted([round(e['rmse'], 4) for e in ep], reverse=True) and ep[-1]['rmse'] < 1.0


@pytest.mark.parametrize('source', ['autograd', 'refgraph'])
def test_svd_step_matches_autograd_golden(source):
    """oracle.steps.svd_step (hand-derived gradients, svd.py:52-80) against the torch-autograd restatement of the same
    graph recorded in tests/golden/svd_golden.npz, and against the reference's own svd.py run through its train() on the
    TF1 stand-in (svd_refgraph_golden.npz)."""
    from oracle import steps
    import refgraph_cases
    z = refgraph_cases.golden('svd', source)
    for name in ('svd', 'svd_d7'):
        P = {k: z['%s/init/%s' % (name, k)].copy() for k in ('U', 'V', 'K')}
        A = {k: np.full_like(v, 0.1) for k, v in P.items()}
        for s in range(2):
            loss = steps.svd_step(P['U'], P['V'], P['K'], A['U'], A['V'], A['K'], z['%s/batch%d' % (name, s)], 0.1, 0.05)
            assert loss == pytest.approx(float(z['%s/loss%d' % (name, s)]), rel=1e-5)
            for k in ('U', 'V', 'K'):
                np.testing.assert_allclose(P[k], z['%s/step%d/%s' % (name, s, k)], rtol=1e-5, atol=1e-6)
                np.testing.assert_allclose(A[k], z['%s/step%d/acc%s' % (name, s, k)], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name', ['svd', 'svd_d7'])
def test_oracle_svd_evaluation_matches_the_reference_train_loop(name):
    """rmse / mae the reference's OWN svd.py train() returned (its clipped predictions, its metrics/rating.py; run on the TF1
    stand-in, oracle/gen_refgraph_golden.py) against the oracle's predictions + metrics on the tables it ended with."""
    import refgraph_cases
    from oracle import rating as orc
    z = refgraph_cases.golden('svd', 'refgraph')
    ev = json.loads(str(z[name + '/eval']))
    tst = np.asarray(ev['tst'])
    pred = orc.svd_predict(z[name + '/step1/U'], z[name + '/step1/V'], z[name + '/step1/K'], tst[:, :2], ev['range_of_ratings'])
    np.testing.assert_allclose(orc.evaluate(tst[:, 2], pred, ev['metrics']), ev['scores'], rtol=1e-6)


@pytest.mark.parametrize('which', ['mf', 'svd'])
def test_oracle_ml100k_runs_equal_the_reference_driver_runs(which):
    """tests/golden/rating_e2e_refgraph_golden.json: the worker() bodies of basic/testmf.py / testsvd.py from the reference's
    OWN modules (loader, sampler_rating with negRatio 0 = deterministic file-order minibatches, MF.train() / SVD.train() on
    the TF1 stand-in; oracle/gen_refgraph_golden.py rating-e2e), started from the oracle runs' initial tables.  The numpy
    oracle's epochs (rating_golden.json: mf_ml100k, svd_ml100k_golden.json -- the ones the GPU tests follow epoch by epoch)
    must be the reference's: mean loss and rmse / mae / mse of every epoch to 1e-5 relative (measured: 7.5e-9 / 5.5e-6)."""
    ref = json.load(open(os.path.join(GOLDEN, 'rating_e2e_refgraph_golden.json')))[which]
    ours = (json.load(open(os.path.join(GOLDEN, 'rating_golden.json')))['mf_ml100k'] if which == 'mf'
            else json.load(open(os.path.join(GOLDEN, 'svd_ml100k_golden.json'))))
    assert len(ref['epochs']) == len(ours['epochs']) and ref['n_factors'] == ours['n_factors'] and ref['batch_size'] == ours['batch_size']
    for a, b in zip(ref['epochs'], ours['epochs']):
        for k in ('loss', 'rmse', 'mae', 'mse'):
            assert b[k] == pytest.approx(a[k], rel=1e-5), (which, k, a, b)
