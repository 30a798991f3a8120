"""Shared by the 'refgraph' tests: the cases of tests/golden/step_refgraph_golden.npz (the reference's own model files run
through their train() on the TF1 stand-in, oracle/gen_refgraph_golden.py) and the stand-in for the reference's sampler."""
import json
import os

import numpy as np
from scipy.sparse import lil_matrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['bpr', 'bpr_w3', 'cml', 'cml_norank_noreg', 'gbpr', 'gbpr_g1', 'wrmf']
KIND = dict(bpr='bpr', bpr_w3='bpr', cml='cml', cml_norank_noreg='cml', gbpr='gbpr', gbpr_g1='gbpr', wrmf='wrmf')


class Overlay(object):
    """The *_refgraph_golden.npz files hold only RESULTS (tables, accumulators, losses, eval); the inputs (initial tables,
    minibatches, hyper-parameters) are those of the autograd golden file of the same name."""

    def __init__(self, base, top):
        self.base, self.top = base, top
        self.files = sorted(set(base.files) | set(top.files))

    def __getitem__(self, k):
        return self.top[k] if k in self.top.files else self.base[k]


SOURCES = ['autograd', 'refgraph']


def golden(kind, source):
    """kind: 'step' | 'tuple' | 'svd'; source 'autograd' = the torch-autograd restatement (oracle/gen_golden.py),
    'refgraph' = the reference's own model files run on the TF1 stand-in (oracle/gen_refgraph_golden.py)."""
    base = np.load(os.path.join(GOLDEN, kind + '_golden.npz'), allow_pickle=False)
    if source == 'autograd':
        return base
    return Overlay(base, np.load(os.path.join(GOLDEN, kind + '_refgraph_golden.npz'), allow_pickle=False))


def load():
    return (np.load(os.path.join(GOLDEN, 'step_golden.npz'), allow_pickle=False),
            np.load(os.path.join(GOLDEN, 'step_refgraph_golden.npz'), allow_pickle=False))


def case(base, ref, name):
    """-> dict(init, batches, hyper, ev, tra, tst, final): everything the reference's train() saw and returned."""
    keys = [k for k in ('U', 'V', 'b') if '%s/init/%s' % (name, k) in base.files]
    init = {k: base['%s/init/%s' % (name, k)].copy() for k in keys}
    nb = sum(1 for f in base.files if f.startswith('%s/batch0/' % name))
    batches = [[base['%s/batch%d/%d' % (name, s, i)] for i in range(nb)] for s in range(2)]
    ev = json.loads(str(ref[name + '/eval']))
    nu, ni = init['U'].shape[0], init['V'].shape[0]
    tra, tst = lil_matrix((nu, ni), dtype=np.float32), lil_matrix((nu, ni), dtype=np.float32)
    for r, c in ev['tra']:
        tra[r, c] = 1
    for r, c in ev['tst']:
        tst[r, c] = 1
    final = {k: ref['%s/step1/%s' % (name, k)] for k in keys}
    final.update({'acc' + k: ref['%s/step1/acc%s' % (name, k)] for k in keys})
    return dict(init=init, batches=batches, hyper=json.loads(str(base[name + '/hyper'])), ev=ev, tra=tra, tst=tst,
                final=final, losses=[float(ref['%s/loss%d' % (name, s)]) for s in range(2)])


class RecordedSampler(object):
    """The reference's samplers hand numpy minibatches out of next_batch(); this one hands out the recorded ones, in the
    formats of sampler_ranking.py (pairs, negatives), sampler_gbpr.py (pairs, negatives, group) and sampler_rating.py
    ([B, 3] user, item, rating)."""

    def __init__(self, kind, batches):
        self.kind, self.batches, self.k = kind, batches, 0

    def next_batch(self):
        b = self.batches[self.k]
        self.k += 1
        if self.kind == 'wrmf':
            return np.concatenate([b[0].astype(np.float64), b[1].astype(np.float64)[:, None]], 1)
        return tuple(x.astype(np.int32) for x in b)
