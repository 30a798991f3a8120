#!/usr/bin/env python
"""Prints how close the ml-100k runs of tests/test_gpu_e2e.py (same constructor calls, same seeds) come to the reference
driver trajectories in tests/golden/e2e*_refgraph_golden.json: the margins of those tests' +-0.02 / 3-5 % thresholds (PRIGP /
CPLR: tests/test_gpu_zz_tuple_trajectories.py)."""
import contextlib
import io
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
NAMES = ['pre', 'recall', 'map', 'mrr', 'ndcg']


def ml100k():
    from scipy.sparse import coo_matrix
    z = np.load(os.path.join(GOLDEN, 'ml100k_fold1.npz'))
    out = {}
    for part in ('tra', 'tst'):
        u, i, r = z[part + '_u'].astype(np.int64), z[part + '_i'].astype(np.int64), z[part + '_r']
        keep = r > 3
        out[part] = coo_matrix((np.ones(int(keep.sum()), dtype=np.float32), (u[keep], i[keep])), shape=(943, 1682)).tolil()
    return out


def main():
    import collaborativefilteringusingtensorflow_b200 as pkg
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_gbpr, sampler_ranking, sampler_rating
    d = ml100k()
    tra, tst = d['tra'], d['tst']
    for kind, fname in (('bpr', 'e2e_refgraph_golden.json'), ('cml', 'e2e_cml_refgraph_golden.json'),
                        ('gbpr', 'e2e_gbpr_refgraph_golden.json'), ('wrmf', 'e2e_wrmf_refgraph_golden.json'),
                        ('prigp', 'e2e_prigp_refgraph_golden.json'), ('cplr', 'e2e_cplr_refgraph_golden.json')):
        gold = json.load(open(os.path.join(GOLDEN, fname)))
        h = gold['hyper']
        if kind == 'bpr':
            m = pkg.BPRMF(943, 1682, h['topN'], 'cv', NAMES, h['reg'], h['n_factors'], h['batch_size'], seed=7)
            s = sampler_ranking.Sampler(trasR=tra, n_neg=h['n_neg'], batch_size=h['batch_size'], seed=7)
        elif kind == 'cml':
            m = pkg.CML(943, 1682, h['topN'], 'cv', NAMES, h['reg_cov'], h['margin'], h['use_rank_weight'], h['clip_norm'],
                        h['n_factors'], h['batch_size'], seed=11)
            s = sampler_ranking.Sampler(tra, n_neg=h['n_neg'], batch_size=h['batch_size'], seed=11)
        elif kind == 'gbpr':
            m = pkg.GBPRMF(943, 1682, h['topN'], h['rho'], h['gsize'], 'cv', NAMES, h['reg'], h['n_factors'], h['batch_size'], seed=5)
            s = sampler_gbpr.Sampler(tra, h['gsize'], h['n_neg'], h['batch_size'], seed=5)
        elif kind == 'prigp':     # the constructor calls and seeds of tests/test_gpu_zz_tuple_trajectories.py; train() builds the sampler
            m = pkg.PRIGP(943, 1682, h['topK'], h['topN'], 'cv', NAMES, h['alpha'], h['reg'], h['n_factors'], h['batch_size'], seed=13)
            s = None
        elif kind == 'cplr':
            m = pkg.CPLR(943, 1682, h['topK'], h['topN'], 'cv', NAMES, h['alpha'], h['beta'], h['gamma'], h['reg'], h['n_factors'],
                         h['batch_size'], seed=13)
            s = None
        else:
            m = pkg.WRMF(943, 1682, h['topN'], 'cv', NAMES, h['weight'], h['reg'], h['n_factors'], h['batch_size'], seed=5)
            s = sampler_rating.Sampler(tra, h['negRatio'], h['batch_size'], seed=5)
        with contextlib.redirect_stdout(io.StringIO()) as log:
            scores = m.train(1, tra, tst, s) if s is not None else m.train(1, tra, tst)
        rows = re.findall(r'iter=\s*(\d+):\s+TraLoss=([0-9.]+).*ndcg=([0-9.]+)', log.getvalue())
        ours = {int(e): (float(l), float(n)) for e, l, n in rows}
        ref = {x['epoch']: (x['TraLoss'], x['ndcg']) for x in gold['history']}
        last = max(ref)
        print('%-5s epochs %d: ndcg ours/ref @20 %.4f/%.4f @%d %.4f/%.4f | loss@%d %.4f/%.4f (%.2f %%) | final %s vs %s' % (
            kind, last, ours[20][1], ref[20][1], last, ours[last][1], ref[last][1], last, ours[last][0], ref[last][0],
            100 * abs(ours[last][0] - ref[last][0]) / ref[last][0], ['%.4f' % x for x in scores], ['%.4f' % x for x in gold['final_scores']]))
        sys.stdout.flush()
        m.close()


if __name__ == '__main__':
    main()
