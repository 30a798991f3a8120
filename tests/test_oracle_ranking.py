"""oracle.ranking vs the reference's own ranking.py: golden vectors (always) + live import (when /root/reference exists)."""
import os
import sys

import numpy as np
import pytest

from oracle import ranking as orc

NAMES = ['pre', 'recall', 'ndcg', 'map', 'mrr']


def test_cv_golden(ranking_golden):
    for c in ranking_golden['cv']:
        yt = [set(x) for x in c['yss_true']]
        got = orc.evaluateCV(yt, c['yss_pred'], NAMES, c['k'])
        for n, v in zip(NAMES, got):
            assert abs(v - c['cv'][n]) < 1e-12, (n, v, c['cv'][n])


def test_toy_vectors_from_survey_appendix_c():
    yt, yp = [{4, 2}, {3, 1}, {1}], [[3, 1, 2], [1, 2], [2, 3, 1]]
    got = orc.evaluateCV(yt, yp, NAMES, 3)
    want = [0.3333333333333333, 0.6666666666666666, 0.6666666666666666, 0.3333333333333333, 0.5555555555555555]
    assert np.allclose(got, want, atol=1e-15)
    yt, yp = [{0, 1, 3, 4, 5, 8, 10, 12, 16, 18}], [list(range(20))]
    for k, want in ((5, [0.8, 0.4, 0.355, 1.0, 0.9558295932317544]), (10, [0.6, 0.6, 0.505, 1.0, 0.9397911964740514]),
                    (20, [0.5, 1.0, 0.7357475805927819, 1.0, 0.9064434192688274])):
        got = orc.evaluateCV(yt, yp, ['pre', 'recall', 'map', 'mrr', 'ndcg'], k)
        assert np.allclose(got, want, atol=1e-12)


def test_loov_golden(ranking_golden):
    for c in ranking_golden['loov']:
        got = orc.evaluateLOOV(c['ys_true'], c['yss_pred'], ['hr', 'arhr'], c['k'])
        assert abs(got[0] - c['loov']['hr']) < 1e-12 and abs(got[1] - c['loov']['arhr']) < 1e-12


def test_unknown_metric_and_errors(ranking_golden):
    assert orc.evaluateCV([{1}], [[1]], ['auc', 'pre'], 1) == ranking_golden['unknown_metric'] == [None, 1.0]
    assert orc.evaluateLOOV([1], [[1]], ['pre'], 1) == [None]
    for bad in (([{1}], [[1], [2]], 3), ([], [], 3), ([{1}], [[1]], 0)):
        with pytest.raises(ValueError):
            orc.precision_k_score(*bad)
        with pytest.raises(ValueError):
            orc.hr_k_score([1] * len(bad[0]), bad[1], bad[2])


@pytest.mark.skipif(not os.path.isdir('/root/reference/src/metrics'), reason='reference not mounted')
def test_live_against_reference():
    sys.path.insert(0, '/root/reference/src/metrics')
    import ranking as ref
    rng = np.random.default_rng(5)
    for _ in range(20):
        n_users, n_items, k = int(rng.integers(1, 30)), int(rng.integers(5, 80)), int(rng.integers(1, 25))
        yt = [set(rng.choice(n_items, size=int(rng.integers(1, 5)), replace=False).tolist()) for _ in range(n_users)]
        yp = [rng.permutation(n_items)[:int(rng.integers(1, k + 5))].tolist() for _ in range(n_users)]
        a, b = orc.evaluateCV(yt, yp, NAMES, k), ref.evaluateCV(yt, yp, NAMES, k)
        assert np.allclose(a, b, atol=1e-12)
        ys = rng.integers(0, n_items, n_users).tolist()
        assert np.allclose(orc.evaluateLOOV(ys, yp, ['hr', 'arhr'], k), ref.evaluateLOOV(ys, yp, ['hr', 'arhr'], k))


def test_oracle_pipeline_reproduces_the_reference_poprank_run(ml100k):
    """The reference's own PopRank (basic/models/pop.py, numpy only) ran in the build container (oracle/gen_golden.py pop)
    on ml-100k fold 1.  The oracle's masked top-N (oracle.scoring: score desc, id asc) over popularity scores and the
    oracle's metrics must reproduce its recommended lists and its metric values: an end-to-end pin of scoring + top-N +
    metrics against reference OUTPUT, not just against a restatement."""
    import json
    from conftest import GOLDEN
    from oracle import scoring
    g = json.load(open(os.path.join(GOLDEN, 'pop_golden.json')))
    tra, tst = ml100k['tra'], ml100k['tst']
    pop = np.asarray((tra != 0).sum(0)).reshape(-1).astype(np.float64)
    ref = g['top10']
    users = ref['test_users']
    scores = np.tile(pop, (len(users), 1))
    lists = scoring.topn_masked(scores, [set(tra.rows[u]) for u in users], 10)
    assert [list(map(int, l)) for l in lists] == ref['lists']
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    got = orc.evaluateCV([set(tst.rows[u]) for u in users], ref['lists'], names, 10)
    for n, v in zip(names, got):
        assert abs(v - ref['scores'][n]) < 1e-12, (n, v, ref['scores'][n])
    loov = orc.evaluateLOOV([tst.rows[u][0] for u in users], ref['lists'], ['hr', 'arhr'], 10)
    assert loov == pytest.approx([g['loov10']['scores']['hr'], g['loov10']['scores']['arhr']], rel=1e-12)
