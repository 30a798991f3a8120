"""Late additions of round 2, collected after every other GPU file because they were written when the round's GPU minutes were
spent and have not run on a B200 yet (their oracle-side twins run on the host: tests/test_oracle_tuples.py,
tests/test_oracle_scoring_samplers.py):
* PRIGP / CPLR on ml-100k for the reference drivers' 50 epochs against the trajectories of the reference's own driver bodies
  (tests/golden/e2e_prigp_refgraph_golden.json, e2e_cplr_refgraph_golden.json);
* the device samplers' distributions against the reference's sampler modules run live (tuple_sampler_golden.json,
  pair_sampler_stats_golden.json)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('cls', ['PRIGP', 'CPLR'])
def test_ml100k_50_epochs_follow_the_reference_driver_runs(ml100k, cls, capsys):
    """tests/golden/e2e_prigp_refgraph_golden.json / e2e_cplr_refgraph_golden.json: the worker() bodies of pl/testprigp.py:21-45
    (topK 5, alpha 10, reg .1, 100 factors, batches of 1000) and pl/testcplr_u.py:21-47 (topK 200, alpha = beta = gamma = 1,
    reg .1, batches of 100) run from the reference's own modules -- its preprocessing inside the classes, its sampler threads,
    its train() for 50 epochs -- on the TF-1.x stand-in (oracle/gen_refgraph_golden.py e2e-prigp | e2e-cplr).  The product gets
    the same positional constructor call and the same train(fold, trasR, tstsR) call and must follow the trajectory: NDCG@100
    and recall@100 within 0.02 at epochs 20 and 50, the last epoch's mean training loss within 5 % (the reference's samplers
    are unseeded; the CPU oracle's own end-to-end run lands within 0.8 % / 0.007 of it:
    profiles/r5_oracle_tuple_trajectories.log, tests/test_oracle_tuples.py)."""
    import re
    import collaborativefilteringusingtensorflow_b200 as pkg
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    gold = json.load(open(os.path.join(GOLDEN, 'e2e_%s_refgraph_golden.json' % cls.lower())))
    h = gold['hyper']
    tra, tst = ml100k['tra'], ml100k['tst']
    assert tra.nnz == gold['nnz']
    if cls == 'PRIGP':                                                                      # testprigp.py:41
        m = pkg.PRIGP(943, 1682, h['topK'], h['topN'], 'cv', names, h['alpha'], h['reg'], h['n_factors'], h['batch_size'], seed=13)
    else:                                                                                   # testcplr_u.py:43
        m = pkg.CPLR(943, 1682, h['topK'], h['topN'], 'cv', names, h['alpha'], h['beta'], h['gamma'], h['reg'], h['n_factors'],
                     h['batch_size'], seed=13)
    scores = m.train(1, tra, tst)
    out = capsys.readouterr().out
    rows = re.findall(r'iter=\s*(\d+):\s+TraLoss=([0-9.]+).*recall=([0-9.]+).*ndcg=([0-9.]+)', out)
    assert len(rows) == h['max_iter'] == 50
    ours = {int(e): (float(l), float(r), float(n)) for e, l, r, n in rows}
    ref = {x['epoch']: (x['TraLoss'], x['recall'], x['ndcg']) for x in gold['history']}
    for ep in (20, 50):
        assert abs(ours[ep][1] - ref[ep][1]) < 0.02 and abs(ours[ep][2] - ref[ep][2]) < 0.02, (ep, ours[ep], ref[ep])
    assert abs(ours[50][0] - ref[50][0]) < 0.05 * ref[50][0], (ours[50], ref[50])
    got, want = dict(zip(names, scores)), dict(zip(names, gold['final_scores']))
    assert abs(got['pre'] - want['pre']) < 0.02 and abs(got['mrr'] - want['mrr']) < 0.05, (got, want)
    m.close()


def test_device_tuple_samplers_match_the_reference_samplers_run_live(ml100k):
    """tests/golden/tuple_sampler_golden.json (the reference's sampler_prigp.Sampler / sampler_uitj_ranking.Sampler run live,
    oracle/gen_golden.py tuple-samplers): the device samplers (cf_sample_tuples), built with the reference's constructor
    calls and read through next_batch(), give the same dtypes, shapes, invariants and distribution statistics; tolerances
    twice those the oracle's samplers meet on the host (tests/test_oracle_tuples.py)."""
    import numpy as np
    from scipy.sparse import csr_matrix
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_prigp, sampler_uitj_ranking
    from oracle import samplers as chk
    from oracle import train_tuples
    gold = json.load(open(os.path.join(GOLDEN, 'tuple_sampler_golden.json')))
    tra = ml100k['tra']
    coef = train_tuples.coefficients(tra, 5, False)                                   # testprigp.py: topK 5, neighbour counts
    s = sampler_prigp.Sampler(tra, csr_matrix(coef), 1000, seed=21)
    got, want = chk.tuple_sampler_stats(tra, coef, [s.next_batch() for _ in range(int(tra.nnz / 1000))], 'prigp'), gold['prigp']
    for k, v in want.items():
        if not isinstance(v, float):
            assert got[k] == v, (k, got[k], v)
    for k, tol in (('frac_rows_with_coef', 1e-12), ('inside_rate', 0.02), ('mean_coef_t_inside', 0.1), ('mean_coef_k_inside', 0.06),
                   ('mean_coef_t_outside', 0.06), ('frac_t_is_positive', 0.02), ('mean_j', 0.02), ('mean_k_outside', 0.02)):
        assert abs(got[k] - want[k]) <= tol, (k, got[k], want[k])
    coefw = train_tuples.coefficients(tra, 200, True)                                 # testcplr_u.py: topK 200, row-normalised sums
    coefw32 = coefw.astype(np.float32).astype(np.float64)                             # the device CSR holds float32 coefficients
    s = sampler_uitj_ranking.Sampler(tra, csr_matrix(coefw), 100, seed=21)
    got, want = chk.tuple_sampler_stats(tra, coefw32, [s.next_batch() for _ in range(442)], 'cplr'), gold['cplr']
    for k, v in want.items():
        if not isinstance(v, float):
            assert got[k] == v, (k, got[k], v)
    for k, tol in (('mean_user_degree', 2.0), ('mean_coef_i', 0.2), ('mean_coef_t', 0.04), ('mean_j', 0.02)):
        assert abs(got[k] - want[k]) <= tol, (k, got[k], want[k])


def test_device_pair_samplers_match_the_reference_samplers_run_live(ml100k):
    """tests/golden/pair_sampler_stats_golden.json (sampler_ranking / sampler_gbpr / sampler_rating of the reference run live,
    oracle/gen_golden.py sampler-stats): the device samplers of the hot path (cf_sample_ranking, cf_sample_rating), built with
    the reference's constructor calls and read through next_batch(), draw from the same distributions -- one epoch each
    (W = 5, B = 100; G = 3), 200 rating batches at negRatio 1; tolerances 1.5 x those the oracle's samplers meet on the host
    (oracle.samplers.PAIR_STATS_TOL: >= 4 standard errors)."""
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_gbpr, sampler_ranking, sampler_rating
    from oracle import samplers as chk
    gold = json.load(open(os.path.join(GOLDEN, 'pair_sampler_stats_golden.json')))
    tra = ml100k['tra']
    nb = int(tra.nnz / 100)
    s = sampler_ranking.Sampler(trasR=tra, n_neg=5, batch_size=100, seed=31)
    assert chk.compare_pair_stats(chk.pair_sampler_stats(tra, 'ranking', [s.next_batch() for _ in range(nb)]), gold['ranking'], 1.5) is None
    s = sampler_gbpr.Sampler(tra, 3, 5, 100, seed=32)
    assert chk.compare_pair_stats(chk.pair_sampler_stats(tra, 'gbpr', [s.next_batch() for _ in range(nb)]), gold['gbpr'], 1.5) is None
    s = sampler_rating.Sampler(tra, 1, 100, seed=34)
    assert chk.compare_pair_stats(chk.pair_sampler_stats(tra, 'rating', [s.next_batch() for _ in range(200)]), gold['rating'], 1.5) is None
