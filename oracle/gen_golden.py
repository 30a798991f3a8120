#!/usr/bin/env python
"""Generate tests/golden/* by RUNNING THE REFERENCE where it runs.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference; the GPU box does not have it):

    python oracle/gen_golden.py            # writes tests/golden/*.json|*.npz

What comes from where
---------------------
* ranking_golden.json  -- outputs of the reference's own /root/reference/src/metrics/ranking.py
  (imported as-is) on its ``__main__`` toy inputs (ranking.py:125,131) and on seeded random inputs.
* sampler_golden.json  -- shapes / dtypes / invariant checks of the reference's own samplers
  (/root/reference/src/samplers/sampler_{ranking,uij_ranking,gbpr,rating}.py imported as-is) on ml-100k fold 1.
* ml100k_fold1.npz     -- the bundled /root/reference/data/movielens/ml-100k/ratings__1_{tra,tst}.txt parsed by the
  reference's own loadSparseR + matBinarize (utils/IOUtil.py:7-16, Util.py:15-16), stored as int16/int8 triplets.
* step_golden.npz      -- TensorFlow is not installable, so the TF graphs of bprmf.py:52-88, cml.py:55-129,
  gbprmf.py:58-106, wrmf.py:52-88 are restated with torch autograd + torch.optim.Adagrad(lr,
  initial_accumulator_value=0.1, eps=0) (== TF1 AdagradOptimizer on summed sparse grads).  This is an
  INDEPENDENT restatement (autodiff, not the hand-derived gradients of oracle/steps.py).
* rating_golden.json   -- outputs of the reference's own /root/reference/src/metrics/rating.py (imported as-is) on its
  ``__main__`` toy input and seeded random inputs; plus the numpy oracle's 5-epoch MF run on ml-100k fold 1 with the
  hyper-parameters of basic/testmf.py:18-26 (scored with the reference's rating.py).
* svd_golden.npz / svd_ml100k_golden.json -- the SVD graph (svd.py:52-80) restated with torch autograd + Adagrad (two steps,
  two shapes) and the numpy oracle's 3-epoch ml-100k run with basic/testsvd.py's hyper-parameters.
* pop_golden.json      -- the reference's own PopRank (basic/models/pop.py, numpy only) run live on ml-100k fold 1: recommended
  lists + metric values (pins the masked top-N with its tie rule and the metrics end to end against the reference).
* tuple_golden.npz     -- PRIGP / CPLR graphs (prigp.py:92-137, cplr_u.py:99-144) restated with torch autograd + Adagrad.
* cf_golden.npz        -- the reference's own ItemCF / UserCF (basic/models/itemcf.py, usercf.py, numpy only) run live on ml-100k
  fold 1, stage by stage (similarities, neighbour choice, scores, lists, metric values).
* e2e_golden.json      -- oracle-trained ml-100k fold-1 metrics (reference hyper-parameters of testbprmf.py:21-30).
"""
import json
import os
import sys

import numpy as np

REF = '/root/reference/src'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)


def ref_import():
    for sub in ('metrics', 'samplers', 'utils'):
        sys.path.insert(0, os.path.join(REF, sub))
    import ranking as ref_ranking          # noqa
    import IOUtil                           # noqa
    import Util                             # noqa
    return ref_ranking, IOUtil, Util


def gen_ranking(ref_ranking):
    cases = []
    toyA = dict(yss_true=[[4, 2], [3, 1], [1]], yss_pred=[[3, 1, 2], [1, 2], [2, 3, 1]], k=3)
    toyB = [dict(yss_true=[[0, 1, 3, 4, 5, 8, 10, 12, 16, 18]], yss_pred=[list(range(20))], k=k) for k in (5, 10, 20)]
    rng = np.random.default_rng(2026)
    rnd = []
    for n_users, n_items, k in ((7, 30, 5), (40, 200, 10), (25, 60, 20), (10, 15, 12)):
        yt = [sorted(rng.choice(n_items, size=int(rng.integers(1, 12)), replace=False).tolist()) for _ in range(n_users)]
        yp = [rng.permutation(n_items)[:int(rng.integers(max(1, k - 3), k + 4))].tolist() for _ in range(n_users)]
        rnd.append(dict(yss_true=yt, yss_pred=yp, k=k))
    names = ['pre', 'recall', 'ndcg', 'map', 'mrr']
    for c in [toyA] + toyB + rnd:
        yt = [set(x) for x in c['yss_true']]
        c['cv'] = dict(zip(names, ref_ranking.evaluateCV(yt, c['yss_pred'], names, c['k'])))
        cases.append(c)
    loov = [dict(ys_true=[1, 5], yss_pred=[[3, 1, 2], [1, 2, 3]], k=3)]
    for n_users, n_items, k in ((30, 50, 5), (12, 20, 10)):
        loov.append(dict(ys_true=rng.integers(0, n_items, n_users).tolist(),
                         yss_pred=[rng.permutation(n_items)[:k + 2].tolist() for _ in range(n_users)], k=k))
    for c in loov:
        c['loov'] = dict(zip(['hr', 'arhr'], ref_ranking.evaluateLOOV(c['ys_true'], c['yss_pred'], ['hr', 'arhr'], c['k'])))
    unknown = ref_ranking.evaluateCV([{1}], [[1]], ['auc', 'pre'], 1)
    json.dump(dict(cv=cases, loov=loov, unknown_metric=unknown), open(os.path.join(OUT, 'ranking_golden.json'), 'w'))
    print('ranking_golden.json:', len(cases), 'cv cases,', len(loov), 'loov cases')


def gen_rating():
    """metrics/rating.py run live on its __main__ toy input (rating.py:33) and on seeded random inputs; plus the
    ml-100k fold-1 MF run of the numpy oracle (deterministic given the initial tables: sampler_rating with negRatio = 0
    walks the training tuples in file order)."""
    import rating as ref_rating            # /root/reference/src/metrics/rating.py, imported as-is
    from oracle import rating as orc, steps
    names = ['mae', 'mse', 'rmse', 'nope']
    cases = [dict(ys_true=[2.5, 1.5, 0], ys_pred=[1, 2, 1])]
    rng = np.random.default_rng(2026)
    for n in (1, 7, 300, 3000):
        t = rng.integers(1, 11, n) / 2.0
        p = np.clip(t + rng.normal(0, 1.0, n), 0.5, 5).astype(np.float32)
        cases.append(dict(ys_true=t.tolist(), ys_pred=[float(x) for x in p]))
    for c in cases:
        yt, yp = np.array(c['ys_true']), np.array(c['ys_pred'])
        c['scores'] = dict(zip(names, ref_rating.evaluate(yt, yp, names)))
        c['mae'], c['mse'], c['rmse'] = (float(ref_rating.mean_absolute_error(yt, yp)), float(ref_rating.mean_squared_error(yt, yp)),
                                         float(ref_rating.root_mean_squared_error(yt, yp)))
        mine = orc.evaluate(yt, yp, names)
        assert mine[3] is None and np.allclose(mine[:3], [c['scores'][m] for m in names[:3]], rtol=1e-14, atol=0)
    # MF on ml-100k fold 1, reference driver hyper-parameters (basic/testmf.py:18-26), 5 epochs, numpy oracle
    d = np.load(os.path.join(OUT, 'ml100k_fold1.npz'))
    tra = np.stack([d['tra_u'], d['tra_i'], d['tra_r']], 1).astype(np.float64)
    tst = np.stack([d['tst_u'], d['tst_i'], d['tst_r']], 1).astype(np.float64)
    nu, ni, k = 943, 1682, 100
    init = np.random.default_rng(7)
    U, V = steps.truncated_normal(init, (nu, k)), steps.truncated_normal(init, (ni, k))
    accU, accV = np.full_like(U, 0.1), np.full_like(V, 0.1)
    hist = orc.mf_train(U, V, accU, accV, tra, tst, ['rmse', 'mae', 'mse'], (1, 5), 0.1, 1000, 5)
    # the reference's own metrics module on the oracle's final predictions
    pred = orc.mf_predict(U, V, tst[:, :2], (1, 5))
    ref_final = ref_rating.evaluate(tst[:, 2], pred, ['rmse', 'mae', 'mse'])
    assert np.allclose(ref_final, hist[-1][1], rtol=1e-12)
    out = dict(cases=cases, mf_ml100k=dict(init_seed=7, n_factors=k, reg=0.1, batch_size=1000, range_of_ratings=[1, 5],
                                          epochs=[dict(loss=h[0], rmse=h[1][0], mae=h[1][1], mse=h[1][2]) for h in hist]))
    json.dump(out, open(os.path.join(OUT, 'rating_golden.json'), 'w'))
    print('rating_golden: %d metric cases; MF ml-100k epochs: %s' % (len(cases), ['%.4f' % h[1][0] for h in hist]))


def gen_pop(bins):
    """The reference's own PopRank (basic/models/pop.py, numpy only) run live on ml-100k fold 1 with testpop.py's
    settings: its recommended lists and metric values pin the masked top-N (ties -> lower item id) and the metrics."""
    sys.path.insert(0, os.path.join(REF, 'models', 'basic', 'models'))
    import pop as ref_pop
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    out = {}
    for topN in (10, 100):
        m = ref_pop.PopRank(943, 1682, topN, 'cv', names)
        scores = m.train(1, bins['tra'], bins['tst'])
        test_users = sorted(set(np.asarray(bins['tst'].nonzero()[0]).tolist()))
        lists = m._PopRank__recommend(bins['tra'], test_users)
        out['top%d' % topN] = dict(scores=dict(zip(names, [float(x) for x in scores])), test_users=[int(u) for u in test_users],
                                   lists=[[int(x) for x in l] for l in (lists if topN == 10 else lists[:60])])   # top-100: first 60 users
    m = ref_pop.PopRank(943, 1682, 10, 'loov', ['hr', 'arhr'])
    out['loov10'] = dict(scores=dict(zip(['hr', 'arhr'], [float(x) for x in m.train(1, bins['tra'], bins['tst'])])))
    json.dump(out, open(os.path.join(OUT, 'pop_golden.json'), 'w'))
    print('pop_golden:', out['top10']['scores'], out['loov10']['scores'])


def gen_tuples():
    """PRIGP / CPLR: TensorFlow is not installable, so the TF graphs of prigp.py:92-137 and cplr_u.py:99-144 are restated
    with torch autograd + torch.optim.Adagrad(initial_accumulator_value=0.1, eps=0) -- an INDEPENDENT restatement of the
    hand-derived gradients of oracle/steps.py (two steps each; PRIGP leaves item_bias out of the optimizer, :134)."""
    import torch
    from oracle.steps import truncated_normal
    torch.set_num_threads(1)
    rng = np.random.default_rng(11)
    out = {}

    def l2(t):
        return (t * t).sum() / 2

    def nls(x):
        return -torch.log(torch.sigmoid(x))

    def prigp_loss(P, t, h):
        U, V, b = P['U'], P['V'], P['b']
        u, m = U[t[:, 0]], t[:, 1:]
        x = (u[:, None, :] * V[m]).sum(-1) + b[m]
        return nls(x[:, 0] - x[:, 1]).sum() + h['alpha'] * nls(x[:, 2] - x[:, 3]).sum() + h['reg'] * (l2(u) + l2(V[m]) + l2(b[m]))

    def cplr_loss(P, t, c, h):
        U, V, b = P['U'], P['V'], P['b']
        u, m = U[t[:, 0]], t[:, 1:]
        x = (u[:, None, :] * V[m]).sum(-1) + b[m]
        ctj, cij = c[:, 1] + 1.0, c[:, 0] + 1.0
        cit = cij / ctj
        return (h['alpha'] * nls(cit * (x[:, 0] - x[:, 1])).sum() + h['beta'] * nls(ctj * (x[:, 1] - x[:, 2])).sum()
                + h['gamma'] * nls(cij * (x[:, 0] - x[:, 2])).sum() + h['reg'] * (l2(u) + l2(V[m]) + l2(b[m])))

    nu, ni, B = 50, 70, 100
    for name, d, width, h in (('prigp', 32, 5, dict(lr=0.1, reg=0.1, alpha=10.0)), ('prigp_d20', 12, 5, dict(lr=0.1, reg=0.01, alpha=1.0)),
                              ('cplr', 32, 4, dict(lr=0.1, reg=0.1, alpha=1.0, beta=0.5, gamma=2.0)),
                              ('cplr_d20', 12, 4, dict(lr=0.1, reg=0.01, alpha=1.0, beta=1.0, gamma=1.0))):
        params = dict(U=truncated_normal(rng, (nu, d)), V=truncated_normal(rng, (ni, d)), b=truncated_normal(rng, (ni,)))
        P = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
        trained = ['U', 'V'] if width == 5 else ['U', 'V', 'b']
        opt = torch.optim.Adagrad([P[k] for k in trained], lr=h['lr'], initial_accumulator_value=0.1, eps=0)
        for k, v in params.items():
            out['%s/init/%s' % (name, k)] = v
        for s_ in range(2):
            t = np.concatenate([rng.integers(0, nu, (B, 1)), rng.integers(0, ni, (B, width - 1))], axis=1)
            if width == 5:
                t[::7, 3], t[::7, 4] = t[::7, 1], t[::7, 2]        # the sampler's default t = i, k = j (sampler_prigp.py:36)
            c = (rng.random((B, 2)) * 3).astype(np.float32)
            c[::5] = 0
            opt.zero_grad()
            loss = prigp_loss(P, torch.tensor(t), h) if width == 5 else cplr_loss(P, torch.tensor(t), torch.tensor(c), h)
            loss.backward()
            opt.step()
            out['%s/loss%d' % (name, s_)] = np.float64(loss.item())
            out['%s/batch%d/tuples' % (name, s_)] = t
            out['%s/batch%d/coefs' % (name, s_)] = c
            for k in P:
                out['%s/step%d/%s' % (name, s_, k)] = P[k].detach().numpy().copy()
                if k in trained:
                    out['%s/step%d/acc%s' % (name, s_, k)] = opt.state[P[k]]['sum'].numpy().copy()
        out[name + '/hyper'] = np.array(json.dumps(h))
    np.savez_compressed(os.path.join(OUT, 'tuple_golden.npz'), **out)
    print('tuple_golden.npz:', len(out), 'arrays')


def gen_cf(nu, ni, bins):
    """The reference's own ItemCF / UserCF (basic/models/itemcf.py, usercf.py: numpy only) run live on ml-100k fold 1
    (testicf.py: topK = 5; usercf default topK = 50; topN = 10).  Their argsort is unstable, so exact lists are only
    defined where no two candidates tie at a cut; everything is therefore pinned STAGE BY STAGE, each stage fed with the
    reference's own previous-stage output: similarity values (bit-exact float32), neighbour choice (value multisets, and
    index sets on rows without a tie at the cut), scores (exact float64), top-N (exact where the scores do not tie)."""
    from oracle import neighbors as onb
    sys.path.insert(0, os.path.join(REF, 'models', 'basic', 'models'))
    import itemcf as ref_icf
    import usercf as ref_ucf
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    tra, tst = bins['tra'], bins['tst']
    test_users = list(set(np.asarray(tst.nonzero()[0])))
    out = {}
    # ---- ItemCF
    m = ref_icf.ItemCF(nu, ni, 5, 10, 'cv', names)
    sim = m.__calsim__(tra)
    assert sim.dtype == np.float32
    mine = onb.cosine_sim(tra.T.tocsr())
    assert np.array_equal(mine, sim), 'oracle item similarity differs from the reference'
    rows_i = [0, 1, 49, 99, 257, 600, 1200, 1681]
    out['icf_sim_rows'], out['icf_sim_rows_val'] = np.array(rows_i, np.int32), sim[rows_i]
    out['icf_sim_sum'], out['icf_sim_sqsum'] = np.float64(sim.astype(np.float64).sum()), np.float64((sim.astype(np.float64) ** 2).sum())
    tie = np.zeros(ni, dtype=bool)
    for i in range(ni):
        srt = np.sort(sim[i])[::-1]
        tie[i] = srt[4] == srt[5] and srt[4] > 0
    simK = m.__topk__(sim.copy())
    nbr_idx = np.full((ni, 5), -1, np.int32)
    nbr_val = np.zeros((ni, 5), np.float32)
    for i in range(ni):
        nz = np.nonzero(simK[i])[0]
        order = nz[np.lexsort((-nz, -simK[i, nz]))]
        nbr_idx[i, :len(order)], nbr_val[i, :len(order)] = order, simK[i, order]
    out['icf_nbr_idx'], out['icf_nbr_val'], out['icf_tie_at_cut'] = nbr_idx, nbr_val, tie
    m._ItemCF__simMat = simK
    users8 = [int(u) for u in sorted(test_users)[::115]][:8]
    pred = np.asarray(m.__predict__(tra, users8))
    assert np.array_equal(onb.item_scores(tra, users8, nbr_idx, nbr_val), pred), 'oracle item scores differ from the reference'
    out['icf_users8'], out['icf_pred8'] = np.array(users8, np.int32), pred
    lists = m._ItemCF__recommend(tra, sorted(test_users))
    out['icf_test_users'] = np.array(sorted(test_users), np.int32)
    out['icf_lists'] = np.array([[int(x) for x in l] + [-1] * (10 - len(l)) for l in lists], np.int32)
    out['icf_scores'] = np.array([float(x) for x in m.train(1, tra, tst)])
    # ---- UserCF
    u = ref_ucf.UserCF(nu, ni, 50, 10, 'cv', names)
    usim = u.__calsim__(tra)
    assert np.array_equal(onb.cosine_sim(tra.tocsr()), usim), 'oracle user similarity differs from the reference'
    rows_u = [0, 5, 100, 404, 700, 942]
    out['ucf_sim_rows'], out['ucf_sim_rows_val'] = np.array(rows_u, np.int32), usim[rows_u]
    out['ucf_sim_sum'] = np.float64(usim.astype(np.float64).sum())
    unbr = np.full((nu, 50), -1, np.int32)
    utie = np.zeros(nu, dtype=bool)
    for a in range(nu):
        inds = np.argsort(usim[a, :])[-50:]                      # usercf.py:37, the reference's own (unstable) choice
        inds = inds[usim[a, inds] > 0]
        inds = inds[np.lexsort((-inds, -usim[a, inds]))]
        unbr[a, :len(inds)] = inds
        srt = np.sort(usim[a])[::-1]
        utie[a] = srt[49] == srt[50] and srt[49] > 0
    out['ucf_nbr_idx'], out['ucf_tie_at_cut'] = unbr, utie
    u._UserCF__simMat = usim
    upred = np.asarray(u.__predict__(tra, users8))
    unbr_val = np.where(unbr >= 0, usim[np.arange(nu)[:, None], np.maximum(unbr, 0)], 0).astype(np.float32)
    assert np.array_equal(onb.user_scores(tra, users8, unbr, unbr_val), upred), 'oracle user scores differ from the reference'
    out['ucf_pred8'] = upred
    ulists = u._UserCF__recommend(tra, sorted(test_users))
    out['ucf_lists'] = np.array([[int(x) for x in l] + [-1] * (10 - len(l)) for l in ulists], np.int32)
    out['ucf_scores'] = np.array([float(x) for x in u.train(1, tra, tst)])
    np.savez_compressed(os.path.join(OUT, 'cf_golden.npz'), **out)
    print('cf_golden: ItemCF', dict(zip(names, out['icf_scores'].round(4))), 'rows with a tie at the cut: %d of %d;' % (tie.sum(), ni),
          'UserCF', dict(zip(names, out['ucf_scores'].round(4))), 'rows with a tie at the cut: %d of %d' % (utie.sum(), nu))


def gen_svd():
    """svd.py:52-80 restated with torch autograd + torch.optim.Adagrad (dense gradient on the kernel matrix, sparse rows
    on the tables: zero-gradient rows are no-ops) -> tests/golden/svd_golden.npz; plus the numpy oracle's ml-100k run
    with basic/testsvd.py's hyper-parameters (32 factors, batches of 100, reg .1, range (1, 5))."""
    import torch
    from oracle import rating as orc, steps
    torch.set_num_threads(1)
    rng = np.random.default_rng(11)
    out = {}
    nu, ni = 60, 90
    for name, d, B in (('svd', 32, 100), ('svd_d7', 7, 64)):
        params = dict(U=steps.truncated_normal(rng, (nu, d)), V=steps.truncated_normal(rng, (ni, d)),
                      K=steps.truncated_normal(rng, (d, d)))
        P = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
        opt = torch.optim.Adagrad(list(P.values()), lr=0.1, initial_accumulator_value=0.1, eps=0)
        O = {k: v.copy() for k, v in params.items()}
        A = {k: np.full_like(v, 0.1) for k, v in params.items()}
        for k, v in params.items():
            out['%s/init/%s' % (name, k)] = v
        for s in range(2):
            uir = np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B), rng.integers(1, 11, B) / 2.0], 1)
            opt.zero_grad()
            u, i = P['U'][torch.tensor(uir[:, 0].astype(np.int64))], P['V'][torch.tensor(uir[:, 1].astype(np.int64))]
            pred = ((u @ P['K']) * i).sum(1)
            l2 = lambda t: (t * t).sum() / 2
            loss = l2(pred - torch.tensor(uir[:, 2].astype(np.float32))) + 0.05 * (l2(u) + l2(i))
            loss.backward()
            opt.step()
            ol = steps.svd_step(O['U'], O['V'], O['K'], A['U'], A['V'], A['K'], uir, 0.1, 0.05)
            assert abs(ol - loss.item()) < 1e-5 * abs(ol)
            out['%s/batch%d' % (name, s)] = uir
            out['%s/loss%d' % (name, s)] = np.float64(loss.item())
            for k in P:
                got, acc = P[k].detach().numpy(), opt.state[P[k]]['sum'].numpy()
                assert np.allclose(O[k], got, rtol=1e-5, atol=1e-6) and np.allclose(A[k], acc, rtol=1e-5, atol=1e-6), (name, s, k)
                out['%s/step%d/%s' % (name, s, k)] = got.copy()
                out['%s/step%d/acc%s' % (name, s, k)] = acc.copy()
    np.savez_compressed(os.path.join(OUT, 'svd_golden.npz'), **out)
    d = np.load(os.path.join(OUT, 'ml100k_fold1.npz'))
    tra = np.stack([d['tra_u'], d['tra_i'], d['tra_r']], 1).astype(np.float64)
    tst = np.stack([d['tst_u'], d['tst_i'], d['tst_r']], 1).astype(np.float64)
    nu, ni, k, B = 943, 1682, 32, 100
    init = np.random.default_rng(8)
    U, V, K = steps.truncated_normal(init, (nu, k)), steps.truncated_normal(init, (ni, k)), steps.truncated_normal(init, (k, k))
    accU, accV, accK = np.full_like(U, 0.1), np.full_like(V, 0.1), np.full_like(K, 0.1)
    epochs = []
    for ep in range(3):
        losses = [steps.svd_step(U, V, K, accU, accV, accK, tra[b * B:(b + 1) * B], 0.1, 0.1) for b in range(len(tra) // B)]
        sc = orc.evaluate(tst[:, 2], orc.svd_predict(U, V, K, tst[:, :2], (1, 5)), ['rmse', 'mae', 'mse'])
        epochs.append(dict(loss=float(np.mean(losses)), rmse=float(sc[0]), mae=float(sc[1]), mse=float(sc[2])))
    json.dump(dict(init_seed=8, n_factors=k, reg=0.1, batch_size=B, range_of_ratings=[1, 5], epochs=epochs),
              open(os.path.join(OUT, 'svd_ml100k_golden.json'), 'w'))
    print('svd_golden.npz: %d arrays; SVD ml-100k epochs: %s' % (len(out), ['%.4f' % e['rmse'] for e in epochs]))


def load_ml100k(IOUtil, Util):
    from scipy.sparse import lil_matrix
    d = '/root/reference/data/movielens/ml-100k/'
    nu, ni = 943, 1682
    raw, bins = {}, {}
    for part in ('tra', 'tst'):
        sR = IOUtil.loadSparseR(nu, ni, d + 'ratings__1_%s.txt' % part)
        raw[part] = sR
        bins[part] = lil_matrix(Util.matBinarize(sR, 3))
    return nu, ni, raw, bins


def gen_ml100k(IOUtil, Util):
    nu, ni, raw, bins = load_ml100k(IOUtil, Util)
    arrs = {}
    for part in ('tra', 'tst'):
        coo = raw[part].tocoo()
        order = np.lexsort((coo.col, coo.row))
        arrs[part + '_u'] = coo.row[order].astype(np.int16)
        arrs[part + '_i'] = coo.col[order].astype(np.int16)
        arrs[part + '_r'] = coo.data[order].astype(np.int8)
    tst_users = sorted(set(np.asarray(bins['tst'].nonzero()[0]).tolist()))
    stats = dict(n_users=nu, n_items=ni, tra_rows=int(raw['tra'].nnz), tst_rows=int(raw['tst'].nnz),
                 tra_pos=int(bins['tra'].nnz), tst_pos=int(bins['tst'].nnz), n_test_users=len(tst_users),
                 max_train_size=int(max(len(r) for r in bins['tra'].rows)))
    np.savez_compressed(os.path.join(OUT, 'ml100k_fold1.npz'), **arrs)
    json.dump(stats, open(os.path.join(OUT, 'ml100k_fold1_stats.json'), 'w'))
    print('ml100k_fold1:', stats)
    return nu, ni, bins


def gen_sampler(bins):
    """Run the reference's own samplers (threads never stop -> caller must os._exit)."""
    import sampler_ranking, sampler_uij_ranking, sampler_gbpr, sampler_rating   # noqa
    from oracle import samplers as chk
    tra = bins['tra']
    out = {}
    s = sampler_ranking.Sampler(trasR=tra, n_neg=5, batch_size=100)
    nb = int(tra.nnz / 100)
    batches = [s.next_batch() for _ in range(nb)]
    # The reference queues a VIEW of its pair array (sampler_ranking.py:27,37) which the next epoch's in-place
    # shuffle (:24) can mutate before the queue's feeder thread pickles it: the last batch of an epoch can come out
    # with pairs that no longer match its negatives.  Record which batches are hit; everything else must be valid.
    invalid = [k for k, b in enumerate(batches) if not chk.negatives_are_valid(tra, np.array(b[0])[:, 0], b[1])]
    batches_ok = [b for k, b in enumerate(batches) if k not in invalid]
    pairs = np.concatenate([np.array(b[0]) for b in batches_ok])
    negs = np.concatenate([b[1] for b in batches_ok])
    out['ranking'] = dict(pairs_dtype=str(batches[0][0].dtype), negs_dtype=str(batches[0][1].dtype),
                          pairs_shape=list(batches[0][0].shape), negs_shape=list(batches[0][1].shape),
                          batches_per_epoch=nb, race_hit_batches=invalid,
                          negatives_valid=chk.negatives_are_valid(tra, pairs[:, 0], negs),
                          pairs_positive=chk.pairs_are_positives(tra, pairs),
                          neg_mean=float(negs.mean()), n_items=int(tra.shape[1]))
    s = sampler_uij_ranking.Sampler(trasR=tra, batch_size=100)
    b = s.next_batch()
    out['uij'] = dict(dtype=str(b.dtype), shape=list(b.shape),
                      negatives_valid=chk.negatives_are_valid(tra, b[:, 0], b[:, 2:3]))
    s = sampler_gbpr.Sampler(tra, 3, 5, 100)
    p, n, g = s.next_batch()
    out['gbpr'] = dict(pairs_dtype=str(p.dtype), negs_dtype=str(n.dtype), group_dtype=str(g.dtype),
                       group_shape=list(g.shape), group_valid=chk.group_members_are_valid(tra, p[:, 1], g),
                       negatives_valid=chk.negatives_are_valid(tra, p[:, 0], n))
    s = sampler_rating.Sampler(tra, 1, 100)
    b0, b1 = s.next_batch(), s.next_batch()
    pos0 = b0[b0[:, 2] > 0]
    first100 = chk._pairs_of(tra)[:100]
    out['rating'] = dict(dtype=str(b0.dtype), shape=list(b0.shape), n_pos=int((b0[:, 2] > 0).sum()),
                         positives_in_file_order=bool(set(map(tuple, pos0[:, :2].astype(int).tolist())) ==
                                                      set(map(tuple, first100.tolist()))),
                         neg_valid=chk.negatives_are_valid(tra, b0[b0[:, 2] == 0][:, 0].astype(int),
                                                           b0[b0[:, 2] == 0][:, 1:2].astype(int)),
                         second_batch_differs=bool(not np.array_equal(np.sort(b0, 0), np.sort(b1, 0))))
    json.dump(out, open(os.path.join(OUT, 'sampler_golden.json'), 'w'), indent=1)
    print('sampler_golden.json:', json.dumps(out)[:400], '...')


def gen_tuple_samplers(bins):
    """The reference's own sampler_prigp.Sampler / sampler_uitj_ranking.Sampler (numpy + a producer thread each) run live on
    ml-100k fold 1 with the coefficient matrices of the drivers' settings (testprigp.py: topK 5, neighbour counts;
    testcplr_u.py: topK 200, similarity sums divided by the row mean) -- the matrices come from the oracle's preprocessing,
    which is pinned to the reference's own (coef_refgraph_golden.npz).  One PRIGP epoch (44 batches of 1000) and 442 CPLR
    batches of 100 -> tests/golden/tuple_sampler_golden.json."""
    import sampler_prigp, sampler_uitj_ranking   # noqa
    from scipy.sparse import lil_matrix
    from oracle import samplers as chk
    from oracle import train_tuples
    tra = bins['tra']
    np.random.seed(2026)
    out = {}
    coef = train_tuples.coefficients(tra, 5, False)
    s = sampler_prigp.Sampler(tra, lil_matrix(coef), 1000)
    nb = int(tra.nnz / 1000)
    out['prigp'] = chk.tuple_sampler_stats(tra, coef, [s.next_batch() for _ in range(nb)], 'prigp')
    coefw = train_tuples.coefficients(tra, 200, True)
    s = sampler_uitj_ranking.Sampler(tra, lil_matrix(coefw), 100)
    out['cplr'] = chk.tuple_sampler_stats(tra, coefw, [s.next_batch() for _ in range(442)], 'cplr')
    json.dump(out, open(os.path.join(OUT, 'tuple_sampler_golden.json'), 'w'), indent=1)
    print('tuple_sampler_golden.json:', json.dumps(out))


def gen_pair_sampler_stats(bins):
    """The reference's sampler_ranking / sampler_gbpr / sampler_rating run live on ml-100k fold 1 (their producer threads,
    np.random seeded here): distribution statistics of one epoch (ranking: W = 5, B = 100; gbpr: G = 3, W = 5, B = 100) and of
    200 rating batches (negRatio 1, B = 100) -> tests/golden/pair_sampler_stats_golden.json.  Batches hit by the reference's
    view/shuffle race (see gen_sampler) are left out."""
    import sampler_ranking, sampler_gbpr, sampler_rating   # noqa
    from oracle import samplers as chk
    tra = bins['tra']
    np.random.seed(2026)
    nb = int(tra.nnz / 100)
    out = {}
    s = sampler_ranking.Sampler(trasR=tra, n_neg=5, batch_size=100)
    bs = [s.next_batch() for _ in range(nb)]
    bs = [b for b in bs if chk.negatives_are_valid(tra, np.array(b[0])[:, 0], b[1]) and chk.pairs_are_positives(tra, np.array(b[0]))]
    out['ranking'] = dict(batches_used=len(bs), **chk.pair_sampler_stats(tra, 'ranking', bs))
    s = sampler_gbpr.Sampler(tra, 3, 5, 100)
    bs = [s.next_batch() for _ in range(nb)]
    bs = [b for b in bs if chk.negatives_are_valid(tra, np.array(b[0])[:, 0], b[1]) and chk.pairs_are_positives(tra, np.array(b[0]))
          and chk.group_members_are_valid(tra, np.array(b[0])[:, 1], b[2])]
    out['gbpr'] = dict(batches_used=len(bs), **chk.pair_sampler_stats(tra, 'gbpr', bs))
    s = sampler_rating.Sampler(tra, 1, 100)
    bs = [s.next_batch() for _ in range(200)]
    out['rating'] = dict(batches_used=len(bs), **chk.pair_sampler_stats(tra, 'rating', bs))
    json.dump(out, open(os.path.join(OUT, 'pair_sampler_stats_golden.json'), 'w'), indent=1)
    print('pair_sampler_stats_golden.json:', json.dumps(out))


# ---------------------------------------------------------------- torch-autograd restatement of the TF graphs
def _torch_models():
    import torch

    def l2(t):
        return (t * t).sum() / 2                                   # tf.nn.l2_loss

    def bpr_loss(P, pairs, negs, h):                               # bprmf.py:52-75
        U, V = P['U'], P['V']
        u, i, j = U[pairs[:, 0]], V[pairs[:, 1]], V[negs]
        ui = (u * i).sum(1)
        uj = (u[:, None, :] * j).sum(-1)
        emb = (-torch.log(torch.sigmoid(ui[:, None] - uj))).sum()
        return emb + h['reg'] * (l2(u) + l2(i) + l2(j))

    def cml_loss(P, pairs, negs, h):                               # cml.py:55-109
        U, V = P['U'], P['V']
        u, i, j = U[pairs[:, 0]], V[pairs[:, 1]], V[negs]
        dp = ((u - i) ** 2).sum(1)
        dn = ((u[:, None, :] - j) ** 2).sum(-1)
        closest = torch.amin(dn, 1)                                # reduce_min: grad split evenly among ties
        lp = torch.relu(dp - closest + h['margin'])
        if h['use_rank_weight']:
            imp = ((dp[:, None] - dn + h['margin']) > 0).float()
            lp = lp * torch.log(imp.mean(1) * V.shape[0] + 1.0)
        loss = lp.sum()
        if h['reg_cov'] > 0:
            loss = loss + h['reg_cov'] * (l2(u) + l2(i) + l2(j))
        return loss

    def gbpr_loss(P, pairs, negs, group, h):                       # gbprmf.py:58-93
        U, V, b = P['U'], P['V'], P['b']
        u, i, js, g = U[pairs[:, 0]], V[pairs[:, 1]], V[negs], U[group]
        ib, jb = b[pairs[:, 1]], b[negs]
        ui_u = (u * i).sum(-1)
        ui_g = (g * i[:, None, :]).sum((1, 2)) / float(group.shape[1])
        ui = h['rho'] * ui_g + (1 - h['rho']) * ui_u + ib
        uj = (u[:, None, :] * js).sum(-1) + jb
        emb = (-torch.log(torch.sigmoid(ui[:, None] - uj))).sum()
        return emb + h['reg'] * (l2(u) + l2(g) + l2(i) + l2(jb))

    def wrmf_loss(P, ui, r, h):                                    # wrmf.py:52-75
        U, V = P['U'], P['V']
        u, i = U[ui[:, 0]], V[ui[:, 1]]
        pred = (u * i).sum(1)
        return l2((pred - r) * float(np.sqrt(h['weight']))) + h['reg'] * (l2(u) + l2(i))

    return torch, bpr_loss, cml_loss, gbpr_loss, wrmf_loss


def gen_steps():
    torch, bpr_loss, cml_loss, gbpr_loss, wrmf_loss = _torch_models()
    from oracle.steps import truncated_normal
    torch.set_num_threads(1)
    rng = np.random.default_rng(7)
    out = {}

    def run(name, params, loss_fn, batches, h, clip=None, n_steps=2):
        P = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
        opt = torch.optim.Adagrad(list(P.values()), lr=h['lr'], initial_accumulator_value=0.1, eps=0)
        for k, v in params.items():
            out['%s/init/%s' % (name, k)] = v
        for s in range(n_steps):
            opt.zero_grad()
            loss = loss_fn(P, *[torch.tensor(x) for x in batches[s]], h)
            loss.backward()
            opt.step()
            if clip is not None:                                   # cml.py:119-129: whole tables, every step
                with torch.no_grad():
                    for k in ('U', 'V'):
                        n = P[k].norm(dim=1, keepdim=True)
                        P[k].mul_(clip / torch.maximum(n, torch.tensor(clip)))
            out['%s/loss%d' % (name, s)] = np.float64(loss.item())
            for k in P:
                out['%s/step%d/%s' % (name, s, k)] = P[k].detach().numpy().copy()
                out['%s/step%d/acc%s' % (name, s, k)] = opt.state[P[k]]['sum'].numpy().copy()
            for bi, x in enumerate(batches[s]):
                out['%s/batch%d/%d' % (name, s, bi)] = x
        out[name + '/hyper'] = np.array(json.dumps(h))

    nu, ni = 60, 90      # small tables so duplicate rows (and u in its own group) are the norm at B=100
    for name, d, W in (('bpr', 100, 1), ('bpr_w3', 20, 3)):
        U, V = truncated_normal(rng, (nu, d)), truncated_normal(rng, (ni, d))
        bt = [(np.stack([rng.integers(0, nu, 100), rng.integers(0, ni, 100)], 1), rng.integers(0, ni, (100, W)))
              for _ in range(2)]
        run(name, dict(U=U, V=V), bpr_loss, bt, dict(lr=0.1, reg=0.1))
    for name, d, W, rw, rc in (('cml', 50, 5, True, 1.0), ('cml_norank_noreg', 12, 4, False, 0.0)):
        U = (0.1 * rng.standard_normal((nu, d))).astype(np.float32)
        V = (0.1 * rng.standard_normal((ni, d))).astype(np.float32)
        if name == 'cml_norank_noreg':
            U *= 4
            V *= 4
        bt = [(np.stack([rng.integers(0, nu, 50), rng.integers(0, ni, 50)], 1), rng.integers(0, ni, (50, W)))
              for _ in range(2)]
        run(name, dict(U=U, V=V), cml_loss, bt,
            dict(lr=0.1, reg_cov=rc, margin=1.0 if rw else 0.5, use_rank_weight=rw, clip_norm=1.0), clip=1.0)
    for name, d, W, G in (('gbpr', 64, 5, 3), ('gbpr_g1', 100, 5, 1)):
        U, V = truncated_normal(rng, (nu, d)), truncated_normal(rng, (ni, d))
        b = truncated_normal(rng, (ni,))
        bt = [(np.stack([rng.integers(0, nu, 100), rng.integers(0, ni, 100)], 1), rng.integers(0, ni, (100, W)),
               rng.integers(0, nu, (100, G))) for _ in range(2)]
        run(name, dict(U=U, V=V, b=b), gbpr_loss, bt, dict(lr=0.1, reg=0.01, rho=0.4))
    U, V = truncated_normal(rng, (nu, 100)), truncated_normal(rng, (ni, 100))
    bt = [(np.stack([rng.integers(0, nu, 200), rng.integers(0, ni, 200)], 1),
           (rng.random(200) < 0.5).astype(np.float32)) for _ in range(2)]
    run('wrmf', dict(U=U, V=V), wrmf_loss, bt, dict(lr=0.1, reg=0.1, weight=2.0))
    np.savez_compressed(os.path.join(OUT, 'step_golden.npz'), **out)
    print('step_golden.npz:', len(out), 'arrays')


def gen_e2e(nu, ni, bins, ref_ranking):
    """ml-100k fold 1, BPRMF with the reference driver's hyper-parameters (testbprmf.py:21-30), trained by the
    numpy oracle; evaluated with the reference's own ranking.py."""
    from oracle import steps, samplers, scoring
    tra, tst = bins['tra'], bins['tst']
    rng = np.random.default_rng(2026)
    d, B, W, reg, lr, epochs, topn = 100, 100, 1, 0.1, 0.1, 20, 10
    U, V = steps.truncated_normal(rng, (nu, d)), steps.truncated_normal(rng, (ni, d))
    aU, aV = np.full_like(U, 0.1), np.full_like(V, 0.1)
    gen = samplers.ranking_batches(tra, W, B, seed=2026)
    nb = int(tra.nnz / B)
    test_users = sorted(set(np.asarray(tst.nonzero()[0]).tolist()))
    yss_true = [set(tst.rows[u]) for u in test_users]
    train_sets = [set(tra.rows[u]) for u in test_users]
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    hist = []
    for ep in range(epochs):
        for _ in range(nb):
            p, n = next(gen)
            steps.bpr_step(U, V, aU, aV, p, n, lr, reg)
        if ep in (9, 19):
            top = scoring.topn_masked(scoring.scores_f64(U[test_users], V), train_sets, topn)
            sc = ref_ranking.evaluateCV(yss_true, [r.tolist() for r in top], names, topn)
            hist.append(dict(epoch=ep + 1, **dict(zip(names, sc))))
            print('e2e oracle BPRMF epoch', ep + 1, hist[-1])
    json.dump(dict(model='BPRMF', hyper=dict(n_factors=d, batch_size=B, n_neg=W, reg=reg, lr=lr, topN=topn),
                   history=hist), open(os.path.join(OUT, 'e2e_golden.json'), 'w'), indent=1)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    what = set(sys.argv[1:]) or {'ranking', 'ml100k', 'steps', 'sampler', 'e2e', 'rating', 'svd', 'pop', 'cf', 'tuples'}
    ref_ranking, IOUtil, Util = ref_import()
    if 'ranking' in what:
        gen_ranking(ref_ranking)
    if 'steps' in what:
        gen_steps()
    if 'rating' in what:
        gen_rating()
    if 'svd' in what:
        gen_svd()
    if 'tuples' in what:
        gen_tuples()
    if what & {'ml100k', 'sampler', 'e2e', 'pop', 'cf', 'tuple-samplers', 'sampler-stats'}:
        nu, ni, bins = gen_ml100k(IOUtil, Util)
        if 'pop' in what:
            gen_pop(bins)
        if 'cf' in what:
            gen_cf(nu, ni, bins)
        if 'e2e' in what:
            gen_e2e(nu, ni, bins, ref_ranking)
        if 'sampler' in what:
            gen_sampler(bins)
        if 'tuple-samplers' in what:
            gen_tuple_samplers(bins)
        if 'sampler-stats' in what:
            gen_pair_sampler_stats(bins)
    sys.stdout.flush()
    os._exit(0)      # the reference's sampler threads never stop (sampler_ranking.py:23)
