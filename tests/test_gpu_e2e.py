"""End to end on ml-100k fold 1 with the reference driver's hyper-parameters (testbprmf.py:21-30): NDCG@10 (the
reference's own definition) within +-0.02 of the CPU oracle run (tests/golden/e2e_golden.json),
and the other three models reach their BASELINE.md ballparks."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
NAMES = ['pre', 'recall', 'map', 'mrr', 'ndcg']


def test_bprmf_ml100k_ndcg_matches_oracle_run(ml100k, capsys):
    from collaborativefilteringusingtensorflow_b200.models.pl.models.bprmf import BPRMF
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_ranking import Sampler
    gold = json.load(open(os.path.join(GOLDEN, 'e2e_golden.json')))
    h = gold['hyper']
    tra, tst = ml100k['tra'], ml100k['tst']
    sampler = Sampler(trasR=tra, n_neg=h['n_neg'], batch_size=h['batch_size'], seed=2026)
    m = BPRMF(943, 1682, h['topN'], 'cv', NAMES, h['reg'], h['n_factors'], h['batch_size'], 20, seed=2026)
    scores = m.train(1, tra, tst, sampler)
    out = capsys.readouterr().out
    assert out.count('cv_fold=1 iter=') == 20 and 'TraLoss=' in out and 'Tst@10:pre=' in out
    want = gold['history'][-1]
    got = dict(zip(NAMES, scores))
    assert abs(got['ndcg'] - want['ndcg']) < 0.02, (got, want)
    assert abs(got['pre'] - want['pre']) < 0.02 and abs(got['mrr'] - want['mrr']) < 0.04
    m.close()


def test_bprmf_ml100k_50_epochs_matches_the_reference_driver_run(ml100k, capsys):
    """tests/golden/e2e_refgraph_golden.json: the body of the reference driver's worker() (pl/testbprmf.py:32-52) run from the
    reference's OWN modules -- IOUtil.loadSparseR, Util.matBinarize, sampler_ranking.Sampler with its producer thread,
    bprmf.BPRMF.train() for its default 50 epochs -- on the TF1 stand-in (oracle/gen_refgraph_golden.py e2e).  The product
    gets the same constructor call and must follow the same trajectory: NDCG@10 within +-0.02 at epochs 20 and 50
    (unseeded sampler, run-to-run noise ~ +-0.01), the mean training loss of the last epoch within 3 %."""
    import re
    from collaborativefilteringusingtensorflow_b200.models.pl.models.bprmf import BPRMF
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_ranking import Sampler
    gold = json.load(open(os.path.join(GOLDEN, 'e2e_refgraph_golden.json')))
    h = gold['hyper']
    tra, tst = ml100k['tra'], ml100k['tst']
    assert tra.nnz == gold['nnz']                                       # the reference's loader + binarisation saw the same matrix
    sampler = Sampler(trasR=tra, n_neg=h['n_neg'], batch_size=h['batch_size'], seed=7)
    m = BPRMF(943, 1682, h['topN'], 'cv', NAMES, h['reg'], h['n_factors'], h['batch_size'], seed=7)   # testbprmf.py:46
    scores = m.train(1, tra, tst, sampler)
    out = capsys.readouterr().out
    rows = re.findall(r'iter=\s*(\d+):\s+TraLoss=([0-9.]+).*ndcg=([0-9.]+)', out)
    assert len(rows) == h['max_iter'] == 50
    ours = {int(e): (float(l), float(n)) for e, l, n in rows}
    ref = {x['epoch']: (x['TraLoss'], x['ndcg']) for x in gold['history']}
    for ep in (20, 50):
        assert abs(ours[ep][1] - ref[ep][1]) < 0.02, (ep, ours[ep], ref[ep])
    assert abs(ours[50][0] - ref[50][0]) < 0.03 * ref[50][0], (ours[50], ref[50])
    assert abs(ours[1][0] - ref[1][0]) < 0.05 * ref[1][0], (ours[1], ref[1])
    got = dict(zip(NAMES, scores))
    want = dict(zip(NAMES, gold['final_scores']))
    assert abs(got['pre'] - want['pre']) < 0.02 and abs(got['recall'] - want['recall']) < 0.02 and abs(got['mrr'] - want['mrr']) < 0.04, (got, want)
    m.close()


def test_cml_ml100k_50_epochs_matches_the_reference_driver_run(ml100k, capsys):
    """tests/golden/e2e_cml_refgraph_golden.json: pl/testcml.py's worker() body from the reference's own modules (loader,
    binarisation, sampler thread, cml.CML.train() for 50 epochs + its topN = 5..1000 tail) on the TF1 stand-in.  The product,
    built with the same constructor call, must follow the trajectory: NDCG@10 within +-0.02 at epochs 20 and 50, the last
    epoch's mean loss within 5 %, and the tail (recommend once at 1000, score the prefixes: cml.py:203-211) within +-0.02
    on ndcg / recall at every topN."""
    import re
    from collaborativefilteringusingtensorflow_b200 import CML
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_ranking import Sampler
    gold = json.load(open(os.path.join(GOLDEN, 'e2e_cml_refgraph_golden.json')))
    h = gold['hyper']
    tra, tst = ml100k['tra'], ml100k['tst']
    sampler = Sampler(tra, n_neg=h['n_neg'], batch_size=h['batch_size'], seed=11)
    m = CML(943, 1682, h['topN'], 'cv', NAMES, h['reg_cov'], h['margin'], h['use_rank_weight'], h['clip_norm'], h['n_factors'],
            h['batch_size'], seed=11)                                                                  # testcml.py:49
    scores = m.train(1, tra, tst, sampler)
    out = capsys.readouterr().out
    rows = re.findall(r'iter=\s*(\d+):\s+TraLoss=([0-9.]+).*ndcg=([0-9.]+)', out)
    assert len(rows) == h['max_iter'] == 50
    ours = {int(e): (float(l), float(n)) for e, l, n in rows}
    ref = {x['epoch']: (x['TraLoss'], x['ndcg']) for x in gold['history']}
    for ep in (20, 50):
        assert abs(ours[ep][1] - ref[ep][1]) < 0.02, (ep, ours[ep], ref[ep])
    assert abs(ours[50][0] - ref[50][0]) < 0.05 * ref[50][0], (ours[50], ref[50])
    tail = re.findall(r'fold=1:\s+Tst@(\d+):(.*)', out)
    assert [int(t) for t, _ in tail] == [x['topN'] for x in gold['tail']] == [5, 10, 20, 50, 100, 200, 500, 1000]
    for (t, txt), want in zip(tail, gold['tail']):
        got = {kv.split('=')[0]: float(kv.split('=')[1]) for kv in txt.split()}
        assert abs(got['ndcg'] - want['ndcg']) < 0.02 and abs(got['recall'] - want['recall']) < 0.02, (t, got, want)
    assert abs(scores[NAMES.index('ndcg')] - gold['final_scores'][NAMES.index('ndcg')]) < 0.02
    m.close()


def test_gbpr_and_wrmf_ml100k_follow_the_reference_driver_runs(ml100k, capsys):
    """e2e_gbpr_refgraph_golden.json / e2e_wrmf_refgraph_golden.json: the worker() bodies of pl/testgbprmf.py (its douban set
    is absent: ml-100k) and basic/testwrmf.py from the reference's own modules (sampler_gbpr / sampler_rating threads, the
    models' own train() for their default 30 / 50 epochs) on the TF1 stand-in.  Same constructor calls here; the last epoch's
    ndcg / recall / pre within +-0.02 (mrr +-0.04), its mean training loss within 5 %."""
    import re
    from collaborativefilteringusingtensorflow_b200 import GBPRMF, WRMF
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_gbpr, sampler_rating
    tra, tst = ml100k['tra'], ml100k['tst']
    for kind in ('gbpr', 'wrmf'):
        gold = json.load(open(os.path.join(GOLDEN, 'e2e_%s_refgraph_golden.json' % kind)))
        h = gold['hyper']
        if kind == 'gbpr':
            m = GBPRMF(943, 1682, h['topN'], h['rho'], h['gsize'], 'cv', NAMES, h['reg'], h['n_factors'], h['batch_size'], seed=5)  # testgbprmf.py:48
            sampler = sampler_gbpr.Sampler(tra, h['gsize'], h['n_neg'], h['batch_size'], seed=5)
        else:
            m = WRMF(943, 1682, h['topN'], 'cv', NAMES, h['weight'], h['reg'], h['n_factors'], h['batch_size'], seed=5)             # testwrmf.py:43
            sampler = sampler_rating.Sampler(tra, h['negRatio'], h['batch_size'], seed=5)
        scores = m.train(1, tra, tst, sampler)
        out = capsys.readouterr().out
        rows = re.findall(r'iter=\s*(\d+):\s+TraLoss=([0-9.]+)', out)
        assert len(rows) == h['max_iter'] == len(gold['history'])
        last = gold['history'][-1]
        assert abs(float(rows[-1][1]) - last['TraLoss']) < 0.05 * last['TraLoss'], (kind, rows[-1], last)
        got, want = dict(zip(NAMES, scores)), dict(zip(NAMES, gold['final_scores']))
        for k, tol in (('ndcg', 0.02), ('recall', 0.02), ('pre', 0.02), ('mrr', 0.04)):
            assert abs(got[k] - want[k]) < tol, (kind, k, got, want)
        m.close()


def test_cml_gbpr_wrmf_train_on_ml100k(ml100k):
    from collaborativefilteringusingtensorflow_b200 import CML, GBPRMF, WRMF
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_gbpr, sampler_ranking, sampler_rating
    tra, tst = ml100k['tra'], ml100k['tst']
    cml = CML(943, 1682, 10, 'cv', NAMES, 1., 1., True, 1.0, 50, 50, 10, verbose=False, seed=1)      # testcml.py:22-34
    s = cml.train(1, tra, tst, sampler_ranking.Sampler(tra, n_neg=5, batch_size=50, seed=1))
    assert s[NAMES.index('ndcg')] > 0.40, s          # 0.53 after 50 epochs
    g = GBPRMF(943, 1682, 100, .4, 1, 'cv', NAMES, .01, 100, 100, 8, verbose=False, seed=1)           # testgbprmf.py:23-32
    s = g.train(1, tra, tst, sampler_gbpr.Sampler(tra, 1, 5, 100, seed=1))
    assert s[NAMES.index('recall')] > 0.5, s         # recall@100 0.665 after 30 epochs
    w = WRMF(943, 1682, 10, 'cv', NAMES, 2., .1, 100, 100, 10, verbose=False, seed=1)                 # testwrmf.py:22-30
    s = w.train(1, tra, tst, sampler_rating.Sampler(tra, 1, 100, seed=1))
    assert s[NAMES.index('ndcg')] > 0.40, s          # 0.51 after 50 epochs


def test_reference_style_sampler_object_is_accepted(ml100k):
    """train() also takes any object with the reference's next_batch() -> numpy arrays (here the CPU oracle's
    restatement of sampler_ranking), uploading batch by batch."""
    from collaborativefilteringusingtensorflow_b200 import BPRMF
    from oracle import samplers

    class RefLike(object):
        def __init__(self, gen):
            self.gen = gen

        def next_batch(self):
            return next(self.gen)

    tra, tst = ml100k['tra'], ml100k['tst']
    m = BPRMF(943, 1682, 10, 'cv', NAMES, .1, 32, 100, 2, verbose=False, seed=1)
    s = m.train(1, tra, tst, RefLike(samplers.ranking_batches(tra, 1, 100, seed=3)))
    assert len(s) == 5 and all(np.isfinite(s))


def test_loov_split(ml100k):
    from collaborativefilteringusingtensorflow_b200 import BPRMF
    from collaborativefilteringusingtensorflow_b200.samplers.sampler_ranking import Sampler
    tra, tst = ml100k['tra'], ml100k['tst']
    m = BPRMF(943, 1682, 10, 'loov', ['hr', 'arhr'], .1, 32, 100, 2, verbose=False, seed=1)
    s = m.train(1, tra, tst, Sampler(tra, 1, 100, seed=1))
    assert 0 <= s[1] <= s[0] <= 919


def test_driver_module_runs_a_fold_from_files(ml100k, tmp_path, capsys):
    """collaborativefilteringusingtensorflow_b200.drivers mirrors the reference's test*.py drivers (file layout, printed lines)."""
    from collaborativefilteringusingtensorflow_b200 import drivers
    from collaborativefilteringusingtensorflow_b200.utils import IOUtil
    d = tmp_path / 'ml-100k'
    d.mkdir()
    for part in ('tra', 'tst'):
        u, i, r = ml100k[part + '_raw']
        IOUtil.saveTriads(list(zip(u.tolist(), i.tolist(), r.astype(float).tolist())), str(d / ('ratings__1_%s.txt' % part)))
    res = drivers.run('bprmf', str(d) + '/', 943, 1682, folds=1, max_iter=3, seed=5)
    out = capsys.readouterr().out
    assert 'ml-100k@1: (943, 1682) 44243 46.92' in out and 'ave: pre,recall,map,mrr,ndcg@10=' in out and 'std:' in out
    assert res.shape == (1, 5) and res[0, 4] > 0.3


def test_mf_driver_runs_a_fold_from_files(ml100k, tmp_path, capsys):
    """The `mf` driver (basic/testmf.py): raw ratings, rating sampler without negatives, RMSE / MAE / MSE lines."""
    from collaborativefilteringusingtensorflow_b200 import drivers
    from collaborativefilteringusingtensorflow_b200.utils import IOUtil
    d = tmp_path / 'ml-100k'
    d.mkdir()
    for part in ('tra', 'tst'):
        u, i, r = ml100k[part + '_raw']
        IOUtil.saveTriads(list(zip(u.tolist(), i.tolist(), r.astype(float).tolist())), str(d / ('ratings__1_%s.txt' % part)))
    res_svd = drivers.run('svd', str(d) + '/', 943, 1682, folds=1, max_iter=2, seed=5)
    assert res_svd.shape == (1, 3) and 0.9 < res_svd[0, 0] < 1.3
    capsys.readouterr()
    res = drivers.run('mf', str(d) + '/', 943, 1682, folds=1, max_iter=4, seed=5)
    out = capsys.readouterr().out
    assert 'ml-100k: (943, 1682) 80000 84.84' in out and 'fold=0: rmse,mae,mse =' in out and 'ave=[' in out and 'std=[0.0000' in out
    assert 'fold=1 iter= 4:' in out and '\tTst:rmse=' in out
    assert res.shape == (1, 3) and 0.9 < res[0, 0] < 1.2 and abs(res[0, 2] - res[0, 0] ** 2) < 1e-9


def test_poprank_equals_the_reference_run(ml100k):
    """PopRank (basic/models/pop.py) is numpy-only, so the reference itself ran in the build container
    (oracle/gen_golden.py pop): its recommended lists (stable sort: ties -> lower item id) and its metric values on ml-100k
    fold 1 must be reproduced exactly by the masked top-N kernel + cf_rank_metrics."""
    import json
    import os
    from conftest import GOLDEN
    from collaborativefilteringusingtensorflow_b200 import PopRank
    g = json.load(open(os.path.join(GOLDEN, 'pop_golden.json')))
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    for topN in (10, 100):
        ref = g['top%d' % topN]
        m = PopRank(943, 1682, topN, 'cv', names)
        scores = m.train(1, ml100k['tra'], ml100k['tst'])
        for name, s in zip(names, scores):
            assert s == pytest.approx(ref['scores'][name], rel=1e-12, abs=1e-15), (topN, name)
        users = ref['test_users'][:len(ref['lists'])]
        assert m.recommend(users, topN, ml100k['tra']) == ref['lists']
    loov = PopRank(943, 1682, 10, 'loov', ['hr', 'arhr']).train(1, ml100k['tra'], ml100k['tst'])
    assert loov == pytest.approx([g['loov10']['scores']['hr'], g['loov10']['scores']['arhr']], rel=1e-12)
