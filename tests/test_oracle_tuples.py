"""The numpy restatement of the PRIGP / CPLR minibatch steps (oracle/steps.py: hand-derived gradients + TF1 Adagrad on the
summed row gradients) against an independent torch-autograd restatement of the same TF graphs (tests/golden/
tuple_golden.npz, oracle/gen_golden.py tuples).  Also against the reference's own prigp.py / cplr_u.py graphs run on the TF-1.x stand-in (tuple_refgraph_golden.npz).  TensorFlow itself is not installable."""
import json
import os

import numpy as np
import pytest

from oracle import steps

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module', params=['autograd', 'refgraph'])
def tg(request):
    """'autograd': the torch-autograd restatement; 'refgraph': the reference's own prigp.py / cplr_u.py graphs run on the TF1
    stand-in (oracle/gen_refgraph_golden.py)."""
    import refgraph_cases
    return refgraph_cases.golden('tuple', request.param)


@pytest.mark.parametrize('name', ['prigp', 'prigp_d20', 'cplr', 'cplr_d20'])
def test_hand_derived_steps_equal_autograd(tg, name):
    h = json.loads(str(tg[name + '/hyper']))
    P = {k: tg['%s/init/%s' % (name, k)].copy() for k in ('U', 'V', 'b')}
    acc = {k: np.full_like(P[k], 0.1) for k in P}
    for s in range(2):
        t, c = tg['%s/batch%d/tuples' % (name, s)], tg['%s/batch%d/coefs' % (name, s)]
        if name.startswith('prigp'):
            loss = steps.prigp_step(P['U'], P['V'], P['b'], acc['U'], acc['V'], t, h['lr'], h['reg'], h['alpha'])
        else:
            loss = steps.cplr_step(P['U'], P['V'], P['b'], acc['U'], acc['V'], acc['b'], t, c, h['lr'], h['reg'], h['alpha'], h['beta'], h['gamma'])
        assert abs(loss - float(tg['%s/loss%d' % (name, s)])) <= 2e-5 * abs(loss)
        for k in P:
            np.testing.assert_allclose(P[k], tg['%s/step%d/%s' % (name, s, k)], rtol=1e-5, atol=1e-6, err_msg='%s step %d %s' % (name, s, k))
            key = '%s/step%d/acc%s' % (name, s, k)
            if key in tg.files:
                np.testing.assert_allclose(acc[k], tg[key], rtol=2e-5, atol=1e-6, err_msg=key)
    if name.startswith('prigp'):
        assert np.array_equal(P['b'], tg[name + '/init/b'])          # prigp.py:134: the bias is not in the optimizer's var_list


@pytest.mark.parametrize('name,weighted', [('prigp', False), ('cplr', True)])
def test_preprocessing_equals_the_reference_methods(ml100k, name, weighted):
    """tests/golden/coef_refgraph_golden.npz: ``__calsim__`` / ``__topk__`` / ``__calcoef__`` of the reference's OWN PRIGP / CPLR
    classes (prigp.py:64-90, cplr_u.py:66-97; instances built on the TF1 stand-in, oracle/gen_refgraph_golden.py coef) on ml-100k
    fold 1.  The oracle's similarities are bit-exact, its kept neighbours are the reference's wherever the cut does not fall
    inside a tie (np.argsort's order is undefined there; the kept similarity VALUES agree on every row), and the coefficient
    matrix built from the reference's neighbours is the reference's."""
    from oracle import neighbors as onb
    g = np.load(os.path.join(GOLDEN, 'coef_refgraph_golden.npz'))
    K, rows = int(g['topK']), g['rows']
    tra = ml100k['tra']
    sim = onb.cosine_sim(tra.tocsr())
    sim = np.asarray(sim.todense()) if hasattr(sim, 'todense') else np.asarray(sim)
    np.fill_diagonal(sim, 0)                                                  # prigp.py:71
    assert np.array_equal(sim[rows].astype(np.float32), g[name + '/sim_rows'])
    assert float(sim.astype(np.float64).sum()) == float(g[name + '/sim_checksum'])
    assert float((sim.astype(np.float64) ** 2).sum()) == float(g[name + '/sim_sq_checksum'])
    idx, val = onb.topk_neighbors(onb.cosine_sim(tra.tocsr()), K)
    ref_idx, ref_val, tie = g[name + '/nbr_idx'], g[name + '/nbr_sim'], g[name + '/tie_at_cut']
    ours_sorted = -np.sort(-np.where(idx >= 0, val, 0).astype(np.float32), axis=1)
    assert np.array_equal(ours_sorted, -np.sort(-ref_val.astype(np.float32), axis=1))          # same kept values on every row
    for u in np.nonzero(~tie)[0]:
        assert set(idx[u][idx[u] >= 0].tolist()) == set(ref_idx[u][ref_val[u] != 0].tolist()), u
    assert int(tie.sum()) == 66
    dense = np.zeros((tra.shape[0], tra.shape[0]), np.float32)
    r, c = np.nonzero(ref_val != 0)
    dense[r, ref_idx[r, c]] = ref_val[r, c]
    coef = onb.coef_matrix(tra, dense, weighted)
    assert int((coef != 0).sum()) == int(g[name + '/coef_nnz'])
    if weighted:
        np.testing.assert_allclose(coef[rows], g[name + '/coef_rows'], rtol=1e-12, atol=0)
        np.testing.assert_allclose(coef.sum(1), g[name + '/coef_row_sums'], rtol=1e-12)
        np.testing.assert_allclose(coef.sum(0), g[name + '/coef_col_sums'], rtol=1e-12)
    else:
        assert np.array_equal(coef[rows], g[name + '/coef_rows'])
        assert np.array_equal(coef.sum(1), g[name + '/coef_row_sums']) and np.array_equal(coef.sum(0), g[name + '/coef_col_sums'])
