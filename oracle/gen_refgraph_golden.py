#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Runs the reference's OWN model files -- imported unmodified from /root/reference -- on the
torch-backed TensorFlow-1.x stand-in of oracle/tf1_shim (TensorFlow has no wheel in this image) and writes
tests/golden/step_refgraph_golden.npz: for every case of tests/golden/step_golden.npz (same initial tables, same
minibatches, same hyper-parameters) the tables, Adagrad accumulators and losses after every step of the reference's
``train()`` loop, and the metric values its end-of-epoch evaluation returned (its own ``__recommend`` = top_k + the Python
filter, its own metrics/ranking.py).  A fake sampler feeds the recorded minibatches through ``next_batch()``.

    python oracle/gen_refgraph_golden.py        (in the build container; /root/reference does not travel to the GPU box)
    python oracle/gen_refgraph_golden.py e2e | e2e-cml | e2e-gbpr | e2e-wrmf     the drivers' worker() bodies on ml-100k (minutes each)
    python oracle/gen_refgraph_golden.py coef                                     PRIGP / CPLR preprocessing on ml-100k
    python oracle/gen_refgraph_golden.py rating-e2e                               testmf.py / testsvd.py bodies on ml-100k
    python oracle/gen_refgraph_golden.py signatures                               constructor / train() / sampler signatures

It also prints how far these results are from the torch-autograd RESTATEMENT that generated step_golden.npz: both must
agree to float32 rounding, which pins the restatement (and with it oracle/steps.py) to the reference's graph code."""
import contextlib
import importlib
import io
import json
import os
import sys

import numpy as np
from scipy.sparse import lil_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/src'
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf1_shim'))
sys.path.insert(1, ROOT)
for p in ('models/pl/models', 'models/basic/models', 'metrics', 'samplers'):
    sys.path.insert(1, os.path.join(REF, p))
import tensorflow as tf          # the stand-in  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
VAR = dict(U='user_embed', V='item_embed', b='item_bias')


class FakeSampler(object):
    """next_batch() hands out the recorded minibatches in turn and snapshots the model state before each one."""

    def __init__(self, batches, pack):
        self.batches, self.pack, self.k, self.snaps = batches, pack, 0, []

    @staticmethod
    def snapshot():
        st = {v.name: v.value.detach().numpy().copy() for v in tf.VARIABLES}
        opt = tf.OPTIMIZERS[-1]                               # the optimizer of train()'s own train_op
        for v, a in opt.accum.items():
            st['acc/' + v.name] = (a.numpy().copy() if a is not None else np.full(v.shape, 0.1, np.float32))
        return st

    def next_batch(self):
        if self.k > 0:
            self.snaps.append(self.snapshot())
        b = self.batches[self.k]
        self.k += 1
        return self.pack(b)


def run_case(g, name, module, cls, kwargs, pack, names):
    tf.reset_default_graph()
    mod = importlib.reload(importlib.import_module(module))
    keys = [k for k in ('U', 'V', 'b') if '%s/init/%s' % (name, k) in g.files]
    init = {k: g['%s/init/%s' % (name, k)] for k in keys}
    nu, ni = init['U'].shape[0], init['V'].shape[0]
    steps = 0
    while '%s/batch%d/0' % (name, steps) in g.files:
        steps += 1
    nb = sum(1 for f in g.files if f.startswith('%s/batch0/' % name))
    batches = [[g['%s/batch%d/%d' % (name, s, i)] for i in range(nb)] for s in range(steps)]
    B = len(batches[0][0])
    rng = np.random.default_rng(sum(ord(ch) for ch in name))      # (hash() of a str changes from process to process)
    cells = rng.choice(nu * ni, steps * B + 40, replace=False)
    tra, tst = lil_matrix((nu, ni), dtype=np.float32), lil_matrix((nu, ni), dtype=np.float32)
    for c in cells[:steps * B]:                                # int(nnz / batch_size) = the recorded steps (bprmf.py:134)
        tra[c // ni, c % ni] = 1
    for c in cells[steps * B:]:
        tst[c // ni, c % ni] = 1
    topn = 5
    model = getattr(mod, cls)(nu, ni, topN=topn, split_method='cv', eval_metrics=names, n_factors=init['U'].shape[1],
                              batch_size=B, max_iter=1, **kwargs)
    for k in keys:
        tf.INIT_OVERRIDE[VAR[k]] = init[k]
    sampler = FakeSampler(batches, pack)
    losses = []
    real_run = tf.Session.run

    def logging_run(self, fetches, feed_dict=None):
        out = real_run(self, fetches, feed_dict)
        if isinstance(fetches, tuple) and not hasattr(fetches, 'indices') and len(fetches) == 2:   # train_op = (optimize, loss)
            losses.append(float(out[1]))
        return out
    tf.Session.run = logging_run
    try:
        with contextlib.redirect_stdout(io.StringIO()) as log:
            scores = model.train(1, tra.tocsr(), tst.tocsr(), sampler)
    finally:
        tf.Session.run = real_run
    sampler.snaps.append(sampler.snapshot())
    assert sampler.k == steps and len(losses) == steps, (sampler.k, len(losses))
    out = {}
    worst = 0.0
    for s in range(steps):
        out['%s/loss%d' % (name, s)] = np.float64(losses[s])
        worst = max(worst, abs(losses[s] - float(g['%s/loss%d' % (name, s)])) / abs(losses[s]))
        for k in keys:
            for pre, src in (('', VAR[k]), ('acc', 'acc/' + VAR[k])):
                a = sampler.snaps[s][src]
                out['%s/step%d/%s%s' % (name, s, pre, k)] = a
                want = g['%s/step%d/%s%s' % (name, s, pre, k)]
                worst = max(worst, float(np.max(np.abs(a - want) / (np.abs(want) + 1e-3))))
    out[name + '/eval'] = np.array(json.dumps(dict(
        topN=topn, metrics=names, scores=[float(x) for x in scores],
        tra=[[int(r), int(c)] for r, c in zip(*tra.nonzero())], tst=[[int(r), int(c)] for r, c in zip(*tst.nonzero())])))
    print('%-18s %s.%s.train(): %d steps, losses %s, eval %s; max rel. distance to the autograd restatement %.2e'
          % (name, module, cls, steps, ['%.4f' % x for x in losses], ['%.4f' % x for x in scores], worst))
    print('   reference printed:', log.getvalue().strip().splitlines()[-1][:150])
    return out, worst


def drive_graph(model, cls, feeds, n_steps):
    """For the models whose train() needs the whole preprocessing pipeline (PRIGP / CPLR: user similarities, their own
    sampler classes): the reference's graph (its ``__optimize__`` / ``__loss`` properties, its placeholders) driven by
    the five lines its train() wraps around it (prigp.py:197-212)."""
    train_op = (model.__optimize__, getattr(model, '_%s__loss' % cls))        # "must before the initializer"
    sess = tf.Session(config=tf.ConfigProto())
    sess.run(tf.global_variables_initializer())
    losses, snaps = [], []
    for s in range(n_steps):
        _, loss = sess.run(train_op, feeds(s))
        losses.append(float(loss))
        snaps.append(FakeSampler.snapshot())
    return losses, snaps


def tuple_cases():
    g = np.load(os.path.join(OUT, 'tuple_golden.npz'))
    out, worst = {}, 0.0
    for name in ('prigp', 'prigp_d20', 'cplr', 'cplr_d20'):
        tf.reset_default_graph()
        h = json.loads(str(g[name + '/hyper']))
        init = {k: g['%s/init/%s' % (name, k)] for k in ('U', 'V', 'b')}
        nu, ni, d = init['U'].shape[0], init['V'].shape[0], init['U'].shape[1]
        B = g[name + '/batch0/tuples'].shape[0]
        if name.startswith('prigp'):
            mod, cls = importlib.reload(importlib.import_module('prigp')), 'PRIGP'
            m = mod.PRIGP(nu, ni, alpha=h['alpha'], reg=h['reg'], n_factors=d, batch_size=B, lr=h['lr'])
            feeds = lambda s: {m._PRIGP__uijtk_placeholder: g['%s/batch%d/tuples' % (name, s)]}
        else:
            mod, cls = importlib.reload(importlib.import_module('cplr_u')), 'CPLR'
            m = mod.CPLR(nu, ni, alpha=h['alpha'], beta=h['beta'], gamma=h['gamma'], reg=h['reg'], n_factors=d, batch_size=B, lr=h['lr'])
            feeds = lambda s: {m._CPLR__uitj_placeholder: g['%s/batch%d/tuples' % (name, s)],
                               m._CPLR__coefs_placeholder: g['%s/batch%d/coefs' % (name, s)]}
        for k in init:
            tf.INIT_OVERRIDE[VAR[k]] = init[k]
        losses, snaps = drive_graph(m, cls, feeds, 2)
        w = 0.0
        for s in range(2):
            out['%s/loss%d' % (name, s)] = np.float64(losses[s])
            w = max(w, abs(losses[s] - float(g['%s/loss%d' % (name, s)])) / abs(losses[s]))
            for k in init:
                for pre, src in (('', VAR[k]), ('acc', 'acc/' + VAR[k])):
                    key = '%s/step%d/%s%s' % (name, s, pre, k)
                    if src not in snaps[s]:
                        assert key not in g.files, key            # PRIGP leaves item_bias out of the optimizer (prigp.py:145)
                        continue
                    out[key] = snaps[s][src]
                    w = max(w, float(np.max(np.abs(out[key] - g[key]) / (np.abs(g[key]) + 1e-3))))
        print('%-18s %s graph: losses %s; max rel. distance to the autograd restatement %.2e' % (name, cls, ['%.4f' % x for x in losses], w))
        worst = max(worst, w)
    np.savez_compressed(os.path.join(OUT, 'tuple_refgraph_golden.npz'), **out)
    print('tuple_refgraph_golden.npz: %d arrays; worst relative distance to tuple_golden.npz %.2e' % (len(out), worst))


def svd_cases():
    """svd.py through its own train(): recorded (user, item, rating) minibatches, its own clipped-prediction evaluation
    with the reference's metrics/rating.py."""
    g = np.load(os.path.join(OUT, 'svd_golden.npz'))
    out, worst = {}, 0.0
    for name in ('svd', 'svd_d7'):
        tf.reset_default_graph()
        mod = importlib.reload(importlib.import_module('svd'))
        init = {k: g['%s/init/%s' % (name, k)] for k in ('U', 'V', 'K')}
        nu, ni, d = init['U'].shape[0], init['V'].shape[0], init['U'].shape[1]
        batches = [[g['%s/batch%d' % (name, s)]] for s in range(2)]
        B = len(batches[0][0])
        rng = np.random.default_rng(5)
        tst = np.stack([rng.integers(0, nu, 80), rng.integers(0, ni, 80), rng.integers(1, 11, 80) / 2.0], 1)
        m = mod.SVD(nu, ni, eval_metrics=['rmse', 'mae'], range_of_ratings=(0.5, 5), reg=0.05, n_factors=d, batch_size=B, max_iter=1, lr=0.1)
        for k, v in dict(U='user_embed', V='item_embed', K='kernel').items():
            tf.INIT_OVERRIDE[v] = init[k]
        sampler = FakeSampler(batches, lambda b: b[0])
        losses = []
        real_run = tf.Session.run

        def logging_run(self, fetches, feed_dict=None):
            o = real_run(self, fetches, feed_dict)
            if isinstance(fetches, tuple) and len(fetches) == 2:
                losses.append(float(o[1]))
            return o
        tf.Session.run = logging_run
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                scores = m.train(1, np.zeros((2 * B, 3)), tst, sampler)
        finally:
            tf.Session.run = real_run
        sampler.snaps.append(sampler.snapshot())
        w = 0.0
        for s in range(2):
            out['%s/loss%d' % (name, s)] = np.float64(losses[s])
            w = max(w, abs(losses[s] - float(g['%s/loss%d' % (name, s)])) / abs(losses[s]))
            for k, v in dict(U='user_embed', V='item_embed', K='kernel').items():
                for pre, src in (('', v), ('acc', 'acc/' + v)):
                    key = '%s/step%d/%s%s' % (name, s, pre, k)
                    out[key] = sampler.snaps[s][src]
                    w = max(w, float(np.max(np.abs(out[key] - g[key]) / (np.abs(g[key]) + 1e-3))))
        out[name + '/eval'] = np.array(json.dumps(dict(metrics=['rmse', 'mae'], scores=[float(x) for x in scores], tst=tst.tolist(),
                                                       range_of_ratings=[0.5, 5])))
        print('%-18s svd.SVD.train(): losses %s, eval %s; max rel. distance to the autograd restatement %.2e'
              % (name, ['%.4f' % x for x in losses], ['%.5f' % x for x in scores], w))
        worst = max(worst, w)
    np.savez_compressed(os.path.join(OUT, 'svd_refgraph_golden.npz'), **out)
    print('svd_refgraph_golden.npz: %d arrays; worst relative distance to svd_golden.npz %.2e' % (len(out), worst))


def e2e_reference(which='bpr', max_iter=50):
    """The body of the reference drivers' worker() (pl/testbprmf.py:32-52, pl/testcml.py:35-57) on ml-100k fold 1 with their
    own hyper-parameters (max_iter is the models' default 50), built from the reference's own modules: utils/IOUtil.
    loadSparseR, utils/Util.matBinarize, samplers/sampler_ranking.Sampler (its producer thread, np.random seeded here) and
    models/bprmf.BPRMF.train() / models/cml.CML.train() on the TF1 stand-in.  Records every line train() printed (TraLoss +
    the five metrics per epoch; CML: also its tail at topN = 5 .. 1000) -> tests/golden/e2e_refgraph_golden.json /
    e2e_cml_refgraph_golden.json.  (CML's tail asks top_k for max|train set| + 1000 of 1682 items; the stand-in returns all
    1682 where TensorFlow would refuse -- the driver was written for larger catalogues.)"""
    import re
    import time
    sys.path.insert(1, os.path.join(REF, 'utils'))
    from IOUtil import loadSparseR
    from Util import matBinarize
    from sampler_ranking import Sampler
    tf.reset_default_graph()
    tf.set_random_seed(2026)
    np.random.seed(2026)
    dataset_dir = '/root/reference/data/movielens/ml-100k/'
    n_users, n_items, fold = 943, 1682, 0
    topN, split_method, eval_metrics = 10, 'cv', ['pre', 'recall', 'map', 'mrr', 'ndcg']
    trasR = lil_matrix(matBinarize(loadSparseR(n_users, n_items, dataset_dir + 'ratings__%d_tra.txt' % (fold + 1)), 3))
    tstsR = lil_matrix(matBinarize(loadSparseR(n_users, n_items, dataset_dir + 'ratings__%d_tst.txt' % (fold + 1)), 3))
    if which == 'bpr':
        reg, n_factors, batch_size, negSample = .1, 100, 100, 1                                   # testbprmf.py:21-30
        mod = importlib.reload(importlib.import_module('bprmf'))
        sampler = Sampler(trasR=trasR, n_neg=negSample, batch_size=batch_size)
        m = mod.BPRMF(n_users, n_items, topN, split_method, eval_metrics, reg, n_factors, batch_size, max_iter)
        hyper = dict(n_factors=n_factors, batch_size=batch_size, n_neg=negSample, reg=reg, lr=0.1, topN=topN, max_iter=max_iter)
        fname = 'e2e_refgraph_golden.json'
    elif which == 'gbpr':
        gsize, rho, reg, topN, n_factors, batch_size, negSample = 1, .4, .01, 100, 100, 100, 5       # testgbprmf.py:23-32 (its douban
        from sampler_gbpr import Sampler as GSampler                                                # set is absent: ml-100k instead)
        mod = importlib.reload(importlib.import_module('gbprmf'))
        sampler = GSampler(trasR, gsize, negSample, batch_size)
        max_iter = 30                                                                               # GBPRMF's default
        m = mod.GBPRMF(n_users, n_items, topN, rho, gsize, split_method, eval_metrics, reg, n_factors, batch_size)
        hyper = dict(n_factors=n_factors, batch_size=batch_size, n_neg=negSample, gsize=gsize, rho=rho, reg=reg, lr=0.1, topN=topN,
                     max_iter=max_iter)
        fname = 'e2e_gbpr_refgraph_golden.json'
    elif which in ('prigp', 'cplr'):
        # pl/testprigp.py:21-45 / pl/testcplr_u.py:21-47: these train() build their own sampler (sampler_prigp /
        # sampler_uitj_ranking threads) from the coefficient matrix of the preprocessing; no sampler argument
        topN, n_factors = 100, 100
        if which == 'prigp':
            topK, alpha, reg, batch_size = 5, 10, .1, 1000
            mod = importlib.reload(importlib.import_module('prigp'))
            m = mod.PRIGP(n_users, n_items, topK, topN, split_method, eval_metrics, alpha, reg, n_factors, batch_size)
            hyper = dict(topK=topK, alpha=alpha, reg=reg, n_factors=n_factors, batch_size=batch_size, lr=0.1, topN=topN, max_iter=max_iter)
        else:
            topK, reg, alpha, beta, gamma, batch_size = 200, .1, 1., 1., 1., 100
            mod = importlib.reload(importlib.import_module('cplr_u'))
            m = mod.CPLR(n_users, n_items, topK, topN, split_method, eval_metrics, alpha, beta, gamma, reg, n_factors, batch_size)
            hyper = dict(topK=topK, alpha=alpha, beta=beta, gamma=gamma, reg=reg, n_factors=n_factors, batch_size=batch_size, lr=0.1,
                         topN=topN, max_iter=max_iter)
        sampler = None
        fname = 'e2e_%s_refgraph_golden.json' % which
    elif which == 'wrmf':
        weight, reg, negRatio, n_factors, batch_size = 2., .1, 1, 100, 100                          # basic/testwrmf.py:22-30
        from sampler_rating import Sampler as RSampler
        mod = importlib.reload(importlib.import_module('wrmf'))
        sampler = RSampler(trasR, negRatio, batch_size)
        m = mod.WRMF(n_users, n_items, topN, split_method, eval_metrics, weight, reg, n_factors, batch_size)
        hyper = dict(n_factors=n_factors, batch_size=batch_size, negRatio=negRatio, weight=weight, reg=reg, lr=0.1, topN=topN,
                     max_iter=max_iter)
        fname = 'e2e_wrmf_refgraph_golden.json'
    else:
        margin, reg_cov, use_rank_weight, clip_norm, n_factors, batch_size, negSample = 1., 1., True, 1.0, 50, 50, 5   # testcml.py:26-34
        mod = importlib.reload(importlib.import_module('cml'))
        sampler = Sampler(trasR, n_neg=negSample, batch_size=batch_size)
        m = mod.CML(n_users, n_items, topN, split_method, eval_metrics, reg_cov, margin, use_rank_weight, clip_norm, n_factors,
                    batch_size, max_iter)
        hyper = dict(n_factors=n_factors, batch_size=batch_size, n_neg=negSample, reg_cov=reg_cov, margin=margin,
                     use_rank_weight=use_rank_weight, clip_norm=clip_norm, lr=0.1, topN=topN, max_iter=max_iter)
        fname = 'e2e_cml_refgraph_golden.json'
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()) as log:
        scores = m.train(fold + 1, trasR, tstsR, sampler) if sampler is not None else m.train(fold + 1, trasR, tstsR)
    hist, tail = [], []
    kv = lambda txt: {x.split('=')[0]: float(x.split('=')[1]) for x in txt.split()}
    for line in log.getvalue().splitlines():
        mt = re.search(r'iter=\s*(\d+):\s+TraLoss=([0-9.]+).*Tst[@0-9]*:(.*?)(\s+timecost.*)?$', line)
        if mt:
            hist.append(dict(epoch=int(mt.group(1)), TraLoss=float(mt.group(2)), **kv(mt.group(3))))
            continue
        mt = re.search(r'fold=\d+:\s+Tst@(\d+):(.*)$', line)
        if mt:
            tail.append(dict(topN=int(mt.group(1)), **kv(mt.group(2))))
    assert len(hist) == max_iter, len(hist)
    json.dump(dict(model=type(m).__name__, source='reference %s + sampler_ranking.py + IOUtil/Util, run on oracle/tf1_shim' % mod.__name__,
                   hyper=hyper, nnz=int(trasR.nnz), final_scores=[float(x) for x in scores], history=hist, tail=tail),
              open(os.path.join(OUT, fname), 'w'), indent=1)
    print('%s: %d epochs in %.0f s; epoch 10 / 20 / %d ndcg %.4f / %.4f / %.4f, final %s'
          % (fname, max_iter, time.time() - t0, max_iter, hist[9]['ndcg'], hist[19]['ndcg'], hist[-1]['ndcg'], ['%.4f' % x for x in scores]))


def coef_cases(topK=5):
    """The numpy preprocessing inside the reference's PRIGP / CPLR classes (prigp.py:64-90, cplr_u.py:66-97: user-user
    cosine similarities, the topK most similar users per user through np.argsort, the coefficient matrix).  The classes
    cannot be constructed without TensorFlow, so this is the first time the reference's own methods run: the stand-in
    builds the instance, ``__calsim__`` / ``__topk__`` / ``__calcoef__`` are called as train() calls them (prigp.py:173-175)
    on ml-100k fold 1 -> tests/golden/coef_refgraph_golden.npz (kept neighbours + similarities of every user, a flag where
    the cut falls inside a tie -- np.argsort's order is undefined there --, row sums and sample rows of both coefficient
    matrices, fp64 checksums)."""
    sys.path.insert(1, os.path.join(REF, 'utils'))
    from IOUtil import loadSparseR
    from Util import matBinarize
    trasR = lil_matrix(matBinarize(loadSparseR(943, 1682, '/root/reference/data/movielens/ml-100k/ratings__1_tra.txt'), 3))
    out = dict(topK=np.int64(topK))
    rows = np.array([0, 1, 7, 100, 400, 640, 941, 942])
    for name, module, cls in (('prigp', 'prigp', 'PRIGP'), ('cplr', 'cplr_u', 'CPLR')):
        tf.reset_default_graph()
        mod = importlib.reload(importlib.import_module(module))
        m = getattr(mod, cls)(943, 1682, topK=topK, n_factors=8)
        sim = m.__calsim__(trasR)
        kept = m.__topk__(sim.copy())
        setattr(m, '_%s__simMat' % cls, kept)
        coef = np.asarray(m.__calcoef__(trasR).todense())
        srt = np.sort(sim, axis=1)[:, ::-1]
        idx = np.full((943, topK), -1, np.int64)
        val = np.zeros((943, topK), sim.dtype)
        for u in range(943):
            nz = np.nonzero(kept[u])[0]
            nz = nz[np.lexsort((nz, -kept[u, nz]))]
            idx[u, :len(nz)], val[u, :len(nz)] = nz, kept[u, nz]
        out.update({name + '/nbr_idx': idx, name + '/nbr_sim': val, name + '/tie_at_cut': srt[:, topK - 1] == srt[:, topK],
                    name + '/sim_rows': sim[rows], name + '/sim_checksum': np.float64(sim.astype(np.float64).sum()),
                    name + '/sim_sq_checksum': np.float64((sim.astype(np.float64) ** 2).sum()),
                    name + '/coef_rows': coef[rows], name + '/coef_row_sums': coef.sum(1), name + '/coef_col_sums': coef.sum(0),
                    name + '/coef_nnz': np.int64((coef != 0).sum())})
        print('%-6s %s: similarities %s %s, %d users with a tie at the cut, coefficient matrix nnz %d, sum %.6f'
              % (name, cls, sim.shape, sim.dtype, int(out[name + '/tie_at_cut'].sum()), int(out[name + '/coef_nnz']), coef.sum()))
    out['rows'] = rows
    np.savez_compressed(os.path.join(OUT, 'coef_refgraph_golden.npz'), **out)
    print('coef_refgraph_golden.npz: %d arrays' % len(out))


def rating_e2e():
    """The worker() bodies of basic/testmf.py:28-48 and basic/testsvd.py:28-48 on ml-100k fold 1 from the reference's own
    modules (IOUtil.loadSparseR, sampler_rating.Sampler with negRatio 0 -- its batches are the file-order slices, shuffled
    inside the slice only, so the run is deterministic up to summation order --, MF.train() / SVD.train() on the TF1
    stand-in), started from the initial tables of the numpy oracle's runs (tests/golden/rating_golden.json: mf_ml100k,
    svd_ml100k_golden.json) and run for as many epochs -> tests/golden/rating_e2e_refgraph_golden.json: per-epoch mean loss
    and full-precision rmse / mae / mse (the reference's metrics/rating.py), next to the distance to the oracle's epochs."""
    from oracle.steps import truncated_normal
    sys.path.insert(1, os.path.join(REF, 'utils'))
    from IOUtil import loadSparseR
    from sampler_rating import Sampler
    d = '/root/reference/data/movielens/ml-100k/'
    n_users, n_items, fold = 943, 1682, 0
    eval_metrics, reg, range_of_ratings = ['rmse', 'mae', 'mse'], .1, (1, 5)
    trasR = loadSparseR(n_users, n_items, d + 'ratings__%d_tra.txt' % (fold + 1))
    tra_tuple = np.array([(user, item, trasR[user, item]) for user, item in np.asarray(trasR.nonzero()).T])
    tstsR = loadSparseR(n_users, n_items, d + 'ratings__%d_tst.txt' % (fold + 1))
    tst_tuple = np.array([(user, item, tstsR[user, item]) for user, item in np.asarray(tstsR.nonzero()).T])
    gold = dict(mf=json.load(open(os.path.join(OUT, 'rating_golden.json')))['mf_ml100k'],
                svd=json.load(open(os.path.join(OUT, 'svd_ml100k_golden.json'))))
    out = {}
    for which in ('mf', 'svd'):
        tf.reset_default_graph()
        g = gold[which]
        k, B, epochs = g['n_factors'], g['batch_size'], len(g['epochs'])
        init = np.random.default_rng(g['init_seed'])
        tf.INIT_OVERRIDE['user_embed'] = truncated_normal(init, (n_users, k))
        tf.INIT_OVERRIDE['item_embed'] = truncated_normal(init, (n_items, k))
        mod = importlib.reload(importlib.import_module(which))
        sampler = Sampler(trasR=trasR, negRatio=.0, batch_size=B)
        if which == 'mf':
            m = mod.MF(n_users, n_items, eval_metrics, range_of_ratings, reg, k, B, epochs)                 # testmf.py:43
        else:
            tf.INIT_OVERRIDE['kernel'] = truncated_normal(init, (k, k))
            m = mod.SVD(n_users, n_items, eval_metrics, range_of_ratings, reg, k, B, epochs)                # testsvd.py:40
        losses, scores_log = [], []
        real_run, real_eval = tf.Session.run, mod.evaluate

        def logging_run(self, fetches, feed_dict=None):
            o = real_run(self, fetches, feed_dict)
            if isinstance(fetches, tuple) and len(fetches) == 2:
                losses.append(float(o[1]))
            return o

        def logging_eval(*a, **kw):
            sc = real_eval(*a, **kw)
            scores_log.append([float(x) for x in sc])
            return sc
        tf.Session.run, mod.evaluate = logging_run, logging_eval
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                m.train(fold + 1, tra_tuple, tst_tuple, sampler)
        finally:
            tf.Session.run, mod.evaluate = real_run, real_eval
        nb = len(tra_tuple) // B
        assert len(losses) == nb * epochs and len(scores_log) == epochs
        eps, worst = [], 0.0
        for e in range(epochs):
            row = dict(loss=float(np.mean(losses[e * nb:(e + 1) * nb])), **dict(zip(eval_metrics, scores_log[e])))
            eps.append(row)
            worst = max(worst, max(abs(row[k_] - g['epochs'][e][k_]) / abs(row[k_]) for k_ in row))
        out[which] = dict(source='reference %s.py + sampler_rating.py + IOUtil, run on oracle/tf1_shim' % which, init_seed=g['init_seed'],
                          n_factors=k, reg=reg, batch_size=B, range_of_ratings=list(range_of_ratings), epochs=eps,
                          max_rel_distance_to_oracle_run=worst)
        print('%-4s %d epochs of %d minibatches: rmse %s; max relative distance to the oracle run %.2e'
              % (which, epochs, nb, ['%.4f' % r['rmse'] for r in eps], worst))
    json.dump(out, open(os.path.join(OUT, 'rating_e2e_refgraph_golden.json'), 'w'), indent=1)


def signatures():
    """The drop-in boundary, read off the reference itself: inspect.signature of every model class' constructor and train(),
    of every sampler's constructor and next_batch() -- possible for the TensorFlow-importing modules now that the stand-in
    lets them be imported -> tests/golden/signatures_golden.json (parameter names in order + their defaults)."""
    import inspect
    sys.path.insert(1, os.path.join(REF, 'utils'))

    def sig(f):
        out = []
        for n, p_ in inspect.signature(f).parameters.items():
            if n == 'self':
                continue
            d = None if p_.default is inspect.Parameter.empty else p_.default
            out.append([n, p_.default is not inspect.Parameter.empty, list(d) if isinstance(d, tuple) else d])
        return out
    out = {}
    for module, cls in (('bprmf', 'BPRMF'), ('cml', 'CML'), ('gbprmf', 'GBPRMF'), ('prigp', 'PRIGP'), ('cplr_u', 'CPLR'),
                        ('wrmf', 'WRMF'), ('mf', 'MF'), ('svd', 'SVD'), ('pop', 'PopRank'), ('itemcf', 'ItemCF'), ('usercf', 'UserCF')):
        mod = importlib.import_module(module)
        c = getattr(mod, cls)
        out['models/%s.%s' % (module, cls)] = dict(init=sig(c.__init__), train=sig(c.train),
                                                  close=hasattr(c, 'close'))
    for module in ('sampler_ranking', 'sampler_uij_ranking', 'sampler_gbpr', 'sampler_rating', 'sampler_prigp', 'sampler_uitj_ranking'):
        c = importlib.import_module(module).Sampler
        out['samplers/%s.Sampler' % module] = dict(init=sig(c.__init__), next_batch=sig(c.next_batch))
    for module, names in (('ranking', ('evaluateCV', 'evaluateLOOV')), ('rating', ('evaluate',)), ('IOUtil', ('loadSparseR',)),
                          ('Util', ('matBinarize',))):
        mod = importlib.import_module(module)
        for n in names:
            out['%s.%s' % (module, n)] = dict(call=sig(getattr(mod, n)))
    json.dump(out, open(os.path.join(OUT, 'signatures_golden.json'), 'w'), indent=1)
    print('signatures_golden.json: %d entries' % len(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'signatures':
        signatures()
        return
    if len(sys.argv) > 1 and sys.argv[1] == 'rating-e2e':
        rating_e2e()
        sys.stdout.flush()
        os._exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'coef':
        coef_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1].startswith('e2e'):
        e2e_reference(dict(e2e='bpr').get(sys.argv[1], sys.argv[1][4:]))          # e2e | e2e-cml | e2e-gbpr | e2e-wrmf
        sys.stdout.flush()
        os._exit(0)                       # the reference sampler's producer thread never ends
    tuple_cases()
    svd_cases()
    g = np.load(os.path.join(OUT, 'step_golden.npz'))
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    out, worst = {}, 0.0
    pn = lambda b: (b[0].astype(np.int32), b[1].astype(np.int32))
    cases = []
    for name in ('bpr', 'bpr_w3'):
        h = json.loads(str(g[name + '/hyper']))
        cases.append((name, 'bprmf', 'BPRMF', dict(reg=h['reg'], lr=h['lr']), pn))
    for name in ('cml', 'cml_norank_noreg'):
        h = json.loads(str(g[name + '/hyper']))
        cases.append((name, 'cml', 'CML', dict(reg_cov=h['reg_cov'], margin=h['margin'], use_rank_weight=h['use_rank_weight'],
                                               clip_norm=h['clip_norm'], lr=h['lr']), pn))
    for name in ('gbpr', 'gbpr_g1'):
        h = json.loads(str(g[name + '/hyper']))
        G = g[name + '/batch0/2'].shape[1]
        cases.append((name, 'gbprmf', 'GBPRMF', dict(rho=h['rho'], gsize=G, reg=h['reg'], lr=h['lr']),
                      lambda b: (b[0].astype(np.int32), b[1].astype(np.int32), b[2].astype(np.int32))))   # sampler_gbpr: pairs, negatives, group
    h = json.loads(str(g['wrmf/hyper']))
    cases.append(('wrmf', 'wrmf', 'WRMF', dict(weight=h['weight'], reg=h['reg'], lr=h['lr']),
                  lambda b: np.concatenate([b[0].astype(np.float64), b[1].astype(np.float64)[:, None]], 1)))
    for name, module, cls, kw, pack in cases:
        o, w = run_case(g, name, module, cls, kw, pack, names)
        out.update(o)
        worst = max(worst, w)
    np.savez_compressed(os.path.join(OUT, 'step_refgraph_golden.npz'), **out)
    print('step_refgraph_golden.npz: %d arrays; worst relative distance to step_golden.npz %.2e' % (len(out), worst))


if __name__ == '__main__':
    main()
