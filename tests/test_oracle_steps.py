"""The numpy oracle (hand-derived gradients) vs the independent torch-autograd restatement of the TF graphs
(tests/golden/step_golden.npz, made by oracle/gen_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from oracle import steps

RTOL = 1e-5   # north_star: 'within 1e-5 relative (fp32)'; atol 1e-6 covers elements near zero (|x| ~ 0.1 scale)


def _load(g, name, k):
    return {p: g['%s/init/%s' % (name, p)].copy() for p in k}


def _close(a, b, rtol=RTOL):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-6)


@pytest.mark.parametrize('name', ['bpr', 'bpr_w3'])
def test_bpr_oracle_matches_autograd(step_golden, name):
    g = step_golden
    h = json.loads(str(g[name + '/hyper']))
    P = _load(g, name, 'UV')
    aU, aV = np.full_like(P['U'], 0.1), np.full_like(P['V'], 0.1)
    for s in range(2):
        pairs, negs = g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)]
        loss = steps.bpr_step(P['U'], P['V'], aU, aV, pairs, negs, h['lr'], h['reg'])
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * abs(loss)
        _close(P['U'], g['%s/step%d/U' % (name, s)])
        _close(P['V'], g['%s/step%d/V' % (name, s)])
        _close(aU, g['%s/step%d/accU' % (name, s)])
        _close(aV, g['%s/step%d/accV' % (name, s)])


@pytest.mark.parametrize('name', ['cml', 'cml_norank_noreg'])
def test_cml_oracle_matches_autograd(step_golden, name):
    g = step_golden
    h = json.loads(str(g[name + '/hyper']))
    P = _load(g, name, 'UV')
    aU, aV = np.full_like(P['U'], 0.1), np.full_like(P['V'], 0.1)
    for s in range(2):
        pairs, negs = g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)]
        f = steps.cml_forward(P['U'], P['V'], pairs, negs, h['margin'], h['use_rank_weight'], P['V'].shape[0])
        assert f['kink'] > 1e-5, 'golden inputs sit on a relu/indicator kink'
        loss = steps.cml_step(P['U'], P['V'], aU, aV, pairs, negs, h['lr'], h['reg_cov'], h['margin'],
                              h['use_rank_weight'], h['clip_norm'])
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * max(1.0, abs(loss))
        _close(P['U'], g['%s/step%d/U' % (name, s)])
        _close(P['V'], g['%s/step%d/V' % (name, s)])
        _close(aU, g['%s/step%d/accU' % (name, s)])
        _close(aV, g['%s/step%d/accV' % (name, s)])


def test_cml_touched_row_clip_equals_whole_table_clip_after_first_step(step_golden):
    """once every row has norm <= clip, clipping only the touched rows == the reference's whole-table clip."""
    g = step_golden
    name = 'cml'
    h = json.loads(str(g[name + '/hyper']))
    A = _load(g, name, 'UV')
    B = {k: v.copy() for k, v in A.items()}
    accA = [np.full_like(A['U'], 0.1), np.full_like(A['V'], 0.1)]
    accB = [x.copy() for x in accA]
    for s in range(2):
        pairs, negs = g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)]
        steps.cml_step(A['U'], A['V'], accA[0], accA[1], pairs, negs, h['lr'], h['reg_cov'], h['margin'], True, 1.0)
        steps.cml_step(B['U'], B['V'], accB[0], accB[1], pairs, negs, h['lr'], h['reg_cov'], h['margin'], True, 1.0,
                       clip_whole_table=(s == 0))
        _close(A['U'], B['U'], 1e-6)
        _close(A['V'], B['V'], 1e-6)


@pytest.mark.parametrize('name', ['gbpr', 'gbpr_g1'])
def test_gbpr_oracle_matches_autograd(step_golden, name):
    g = step_golden
    h = json.loads(str(g[name + '/hyper']))
    P = _load(g, name, 'UVb')
    acc = {k: np.full_like(v, 0.1) for k, v in P.items()}
    for s in range(2):
        pairs, negs, group = (g['%s/batch%d/%d' % (name, s, k)] for k in range(3))
        loss = steps.gbpr_step(P['U'], P['V'], P['b'], acc['U'], acc['V'], acc['b'], pairs, negs, group,
                               h['lr'], h['reg'], h['rho'])
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * abs(loss)
        for k in 'UVb':
            _close(P[k], g['%s/step%d/%s' % (name, s, k)])
            _close(acc[k], g['%s/step%d/acc%s' % (name, s, k)])


def test_wrmf_oracle_matches_autograd(step_golden):
    g = step_golden
    name = 'wrmf'
    h = json.loads(str(g[name + '/hyper']))
    P = _load(g, name, 'UV')
    aU, aV = np.full_like(P['U'], 0.1), np.full_like(P['V'], 0.1)
    for s in range(2):
        ui, r = g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)]
        uir = np.concatenate([ui.astype(np.float64), r[:, None].astype(np.float64)], axis=1)
        loss = steps.wrmf_step(P['U'], P['V'], aU, aV, uir, h['lr'], h['reg'], h['weight'])
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * abs(loss)
        _close(P['U'], g['%s/step%d/U' % (name, s)])
        _close(P['V'], g['%s/step%d/V' % (name, s)])
        _close(aU, g['%s/step%d/accU' % (name, s)])


def test_truncated_normal_bounds():
    x = steps.truncated_normal(np.random.default_rng(0), (1000, 8), 0.0, 0.1)
    assert x.dtype == np.float32 and np.abs(x).max() <= 0.2 + 1e-7 and 0.07 < x.std() < 0.1


def test_torch_multithreaded_port_matches_numpy_oracle():
    """oracle/steps_torch.py (bench.py's all-cores CPU baseline) == oracle/steps.py."""
    import torch
    from oracle import steps_torch
    rng = np.random.default_rng(3)
    nu, ni, d, B, W = 200, 300, 32, 256, 4
    U0 = (0.05 * rng.standard_normal((nu, d))).astype(np.float32)
    V0 = (0.05 * rng.standard_normal((ni, d))).astype(np.float32)
    pairs = np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B)], 1)
    negs = rng.integers(0, ni, (B, W))
    for kind in ('bpr', 'cml'):
        a = [U0.copy(), V0.copy(), np.full_like(U0, 0.1), np.full_like(V0, 0.1)]
        b = [torch.from_numpy(x.copy()) for x in a]
        if kind == 'bpr':
            steps.bpr_step(a[0], a[1], a[2], a[3], pairs, negs, 0.1, 0.05)
            steps_torch.bpr_step(b[0], b[1], b[2], b[3], torch.from_numpy(pairs), torch.from_numpy(negs), 0.1, 0.05)
        else:
            steps.cml_step(a[0], a[1], a[2], a[3], pairs, negs, 0.1, 1.0, 1.0, True, 1.0)
            steps_torch.cml_step(b[0], b[1], b[2], b[3], torch.from_numpy(pairs), torch.from_numpy(negs), 0.1, 1.0, 1.0, True, 1.0)
        for x, y in zip(a, b):
            np.testing.assert_allclose(y.numpy(), x, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize('name', ['bpr', 'bpr_w3', 'cml', 'cml_norank_noreg', 'gbpr', 'gbpr_g1', 'wrmf'])
def test_oracle_scoring_and_metrics_match_the_reference_train_loop(name):
    """The metric values the reference's OWN train() returned after its end-of-epoch evaluation (its __recommend = top_k +
    Python filter, its metrics/ranking.py; run on the TF1 stand-in, oracle/gen_refgraph_golden.py) against the oracle's
    scoring + masked top-N + metrics on the tables the reference ended with."""
    from oracle import ranking, scoring
    import refgraph_cases as R
    c = R.case(*R.load(), name)
    ev, f = c['ev'], c['final']
    kind = dict(cml=scoring.NEG_SQDIST, gbpr=scoring.DOT_BIAS).get(R.KIND[name], scoring.DOT)
    users = sorted(set(r for r, _ in ev['tst']))
    truth = [set(c['tst'].rows[u]) for u in users]
    seen = [set(c['tra'].rows[u]) for u in users]
    topn = 1000 if R.KIND[name] == 'cml' else ev['topN']                      # cml.py:203-211: the tail re-scores at 1000
    s = scoring.scores_f64(f['U'][users], f['V'], kind, f.get('b'))
    top = scoring.topn_masked(s, seen, min(topn, s.shape[1]))
    got = ranking.evaluateCV(truth, [[int(x) for x in r if x >= 0] for r in top], ev['metrics'], topn)
    np.testing.assert_allclose(got, ev['scores'], rtol=0, atol=1e-12)


def test_oracle_bprmf_ml100k_run_follows_the_reference_driver_run():
    """tests/golden/e2e_golden.json is the numpy oracle trained on ml-100k fold 1 with testbprmf.py's hyper-parameters;
    e2e_refgraph_golden.json is the reference driver's own worker() body (its loader, sampler thread and BPRMF.train()) on the
    TF-1.x stand-in.  Two unseeded / differently seeded runs of the same procedure: NDCG@10 after 10 and 20 epochs within 0.03,
    precision and recall within 0.02 (the reference's run-to-run noise at topN = 10 is ~0.01-0.02)."""
    import json
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    ours = json.load(open(os.path.join(golden, 'e2e_golden.json')))
    ref = json.load(open(os.path.join(golden, 'e2e_refgraph_golden.json')))
    for k in ('n_factors', 'batch_size', 'n_neg', 'reg', 'lr', 'topN'):
        assert ours['hyper'][k] == ref['hyper'][k]
    by_epoch = {x['epoch']: x for x in ref['history']}
    for x in ours['history']:
        r = by_epoch[x['epoch']]
        assert abs(x['ndcg'] - r['ndcg']) < 0.03 and abs(x['pre'] - r['pre']) < 0.02 and abs(x['recall'] - r['recall']) < 0.02, (x, r)
