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


def test_tuple_samplers_follow_the_reference_samplers(ml100k):
    """oracle.samplers.prigp_batches / uitj_batches against the contracts of sampler_prigp.py:22-52 and
    sampler_uitj_ranking.py:22-40 (shapes, dtypes, set memberships, the coefficient order of the collaborative pair, one pass
    over the shuffled positives per epoch, Phi(nnz / n_items) as the rate of the inside-the-row branch)."""
    from scipy.sparse import csr_matrix
    from oracle import samplers, train_tuples
    tra = ml100k['tra']
    ni = tra.shape[1]
    pos = np.asarray(tra.todense()) > 0
    coef = train_tuples.coefficients(tra, 5, False)
    B, nb = 1000, int(tra.nnz / 1000)
    gen = samplers.prigp_batches(tra, csr_matrix(coef), B, seed=4)
    epoch = np.concatenate([next(gen) for _ in range(nb)])
    assert epoch.dtype == np.int64 and epoch.shape == (nb * B, 5)
    u, i, j, t, k = epoch.T
    assert pos[u, i].all() and not pos[u, j].any()
    assert len(set(zip(u.tolist(), i.tolist()))) == nb * B                      # an epoch draws every positive at most once
    has = (coef != 0).sum(1)[u] > 0
    assert np.array_equal(t[~has], i[~has]) and np.array_equal(k[~has], j[~has])   # :36: no coefficients -> (t, k) = (i, j)
    assert (coef[u[has], t[has]] != 0).all()
    inside = has & (coef[u, k] != 0)
    assert (coef[u[inside], t[inside]] > coef[u[inside], k[inside]]).all()       # :44-48: other value, larger one first
    distinct = np.array([len(set(row[row != 0])) for row in coef])
    can = has & (distinct[u] > 1)
    from math import erf, sqrt
    want = np.mean([0.5 * (1 + erf(x / sqrt(2))) for x in ((coef != 0).sum(1)[u[can]] / float(ni))])
    assert abs(inside[can].mean() - want) < 0.01, (inside[can].mean(), want)     # :43 draws a NORMAL against nnz / n_items
    # CPLR: rows normalised by their mean, (u, i, t, j) with t a coefficient column outside the positives
    coefw = train_tuples.coefficients(tra, 200, True)
    nz = coefw != 0
    np.testing.assert_allclose(coefw.sum(1)[nz.any(1)] / nz.sum(1)[nz.any(1)], 1.0, rtol=1e-12)     # cplr_u.py:193-196
    uitj, c = next(samplers.uitj_batches(tra, csr_matrix(coefw), 5000, seed=4))
    assert uitj.dtype == np.int64 and uitj.shape == (5000, 4) and c.dtype == np.float64 and c.shape == (5000, 2)
    u, i, t, j = uitj.T
    assert pos[u, i].all() and (nz[u, t] & ~pos[u, t]).all() and not (pos[u, j] | nz[u, j]).any()
    assert np.array_equal(c[:, 0], coefw[u, i]) and np.array_equal(c[:, 1], coefw[u, t])
    assert len(np.unique(u)) > 800                                               # users drawn uniformly, not by their degree


@pytest.mark.parametrize('which', ['prigp', 'cplr'])
def test_oracle_training_follows_the_reference_driver_runs(ml100k, which):
    """tests/golden/e2e_prigp_refgraph_golden.json / e2e_cplr_refgraph_golden.json: the worker() bodies of pl/testprigp.py and
    pl/testcplr_u.py run from the reference's own modules (their preprocessing, their sampler threads, their train()) on the
    TF-1.x stand-in.  The oracle's end-to-end restatement (oracle/train_tuples.py: its own preprocessing, samplers and steps;
    another seed, the reference is unseeded) follows the same trajectory: after 5 and 10 epochs the mean training loss within
    3 % and pre / recall / ndcg @100 within 0.02 (the full 50 epochs, tools/oracle_tuple_trajectories.py ->
    profiles/r5_oracle_tuple_trajectories.log: last-epoch loss within 0.8 %, ndcg within 0.007)."""
    from oracle import train_tuples
    gold = json.load(open(os.path.join(GOLDEN, 'e2e_%s_refgraph_golden.json' % which)))
    assert ml100k['tra'].nnz == gold['nnz']
    hist = train_tuples.run(which, ml100k['tra'], ml100k['tst'], gold['hyper'], seed=3, epochs=10, eval_epochs={5, 10})
    ref = {x['epoch']: x for x in gold['history']}
    assert [x['epoch'] for x in hist] == [5, 10]
    for x in hist:
        r = ref[x['epoch']]
        assert abs(x['TraLoss'] - r['TraLoss']) < 0.03 * r['TraLoss'], (x, r)
        for k in ('pre', 'recall', 'ndcg'):
            assert abs(x[k] - r[k]) < 0.02, (k, x, r)


def test_tuple_samplers_match_the_reference_samplers_run_live(ml100k):
    """tests/golden/tuple_sampler_golden.json: the reference's own sampler_prigp.Sampler / sampler_uitj_ranking.Sampler (numpy
    + their producer threads) run live on ml-100k fold 1 with the drivers' coefficient matrices (oracle/gen_golden.py
    tuple-samplers): one PRIGP epoch, 442 CPLR batches.  The oracle's restatements give the same dtypes, shapes and
    invariants and the same distribution: branch rates, mean coefficients of the drawn items, mean negatives, the users'
    degree profile -- within a few standard errors of samples that size."""
    from scipy.sparse import csr_matrix
    from oracle import samplers, train_tuples
    gold = json.load(open(os.path.join(GOLDEN, 'tuple_sampler_golden.json')))
    tra = ml100k['tra']
    coef = train_tuples.coefficients(tra, 5, False)
    gen = samplers.prigp_batches(tra, csr_matrix(coef), 1000, seed=21)
    got, want = samplers.tuple_sampler_stats(tra, coef, [next(gen) for _ in range(int(tra.nnz / 1000))], 'prigp'), gold['prigp']
    for k, v in want.items():
        if not isinstance(v, float):
            assert got[k] == v, (k, got[k], v)
    for k, tol in (('frac_rows_with_coef', 1e-12), ('inside_rate', 0.01), ('mean_coef_t_inside', 0.05), ('mean_coef_k_inside', 0.03),
                   ('mean_coef_t_outside', 0.03), ('frac_t_is_positive', 0.01), ('mean_j', 0.01), ('mean_k_outside', 0.01)):
        assert abs(got[k] - want[k]) <= tol, (k, got[k], want[k])
    coefw = train_tuples.coefficients(tra, 200, True)
    gen = samplers.uitj_batches(tra, csr_matrix(coefw), 100, seed=21)
    got, want = samplers.tuple_sampler_stats(tra, coefw, [next(gen) for _ in range(442)], 'cplr'), gold['cplr']
    for k, v in want.items():
        if not isinstance(v, float):
            assert got[k] == v, (k, got[k], v)
    for k, tol in (('mean_user_degree', 1.0), ('mean_coef_i', 0.1), ('mean_coef_t', 0.02), ('mean_j', 0.01)):
        assert abs(got[k] - want[k]) <= tol, (k, got[k], want[k])
