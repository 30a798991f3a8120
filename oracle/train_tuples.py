"""CPU oracle: PRIGP / CPLR trained end to end the way the reference's train() does it.  TEST INFRASTRUCTURE.

Restates /root/reference/src/models/pl/models/prigp.py:172-228 and cplr_u.py:178-276 from the oracle's own pieces:
preprocessing (oracle.neighbors: __calsim__ / __topk__ / __calcoef__; CPLR divides every coefficient row by its mean,
cplr_u.py:193-196), the tuple samplers (oracle.samplers.prigp_batches / uitj_batches), int(nnz / batch_size) steps per epoch
(oracle.steps.prigp_step / cplr_step, constant learning rate: the `lr *= .98` of the loop only changes the printed value,
the optimizer was built before it), and after every epoch the masked top-N of U V^T + b for the test users scored with
oracle.ranking.evaluateCV.

PINNED at trajectory level against the reference's own driver bodies run on the TF-1.x stand-in
(tests/golden/e2e_prigp_refgraph_golden.json, e2e_cplr_refgraph_golden.json; tests/test_oracle_tuples.py).
"""
import numpy as np

from . import neighbors, ranking, samplers, scoring, steps

NAMES = ['pre', 'recall', 'map', 'mrr', 'ndcg']


def coefficients(tra, topK, weighted):
    """train()'s first lines: the dense float64 coefficient matrix; CPLR's per-row normalisation included."""
    idx, val = neighbors.topk_neighbors(neighbors.cosine_sim(tra.tocsr()), topK)
    dense = np.zeros((tra.shape[0], tra.shape[0]), np.float32)
    r, c = np.nonzero(idx >= 0)
    dense[r, idx[r, c]] = val[r, c]
    coef = neighbors.coef_matrix(tra, dense, weighted)
    if weighted:                                            # cplr_u.py:193-196: row / (row sum / row nnz)
        nnz = (coef != 0).sum(1)
        ave = np.divide(coef.sum(1), nnz, out=np.zeros(len(coef)), where=nnz > 0)
        coef[ave > 0] /= ave[ave > 0, None]
    return coef


def run(which, tra, tst, hyper, seed=0, epochs=None, eval_epochs=None):
    """Returns [{epoch, TraLoss, pre, recall, map, mrr, ndcg}] for the epochs in eval_epochs (default: all)."""
    from scipy.sparse import csr_matrix
    nu, ni = tra.shape
    d, B, topn = hyper['n_factors'], hyper['batch_size'], hyper['topN']
    epochs = epochs or hyper['max_iter']
    rng = np.random.default_rng(seed)
    U, V = steps.truncated_normal(rng, (nu, d)), steps.truncated_normal(rng, (ni, d))
    b = steps.truncated_normal(rng, (ni,))
    aU, aV, ab = np.full_like(U, 0.1), np.full_like(V, 0.1), np.full_like(b, 0.1)
    coef = csr_matrix(coefficients(tra, hyper['topK'], which == 'cplr'))
    gen = (samplers.prigp_batches(tra, coef, B, seed) if which == 'prigp' else samplers.uitj_batches(tra, coef, B, seed))
    nb = int(tra.nnz / B)
    test_users = sorted(set(np.asarray(tst.nonzero()[0]).tolist()))
    yss_true = [set(tst.rows[u]) for u in test_users]
    train_sets = [set(tra.rows[u]) for u in test_users]
    hist = []
    for ep in range(epochs):
        losses = []
        for _ in range(nb):
            if which == 'prigp':
                losses.append(steps.prigp_step(U, V, b, aU, aV, next(gen), hyper['lr'], hyper['reg'], hyper['alpha']))
            else:
                t, c = next(gen)
                losses.append(steps.cplr_step(U, V, b, aU, aV, ab, t, c, hyper['lr'], hyper['reg'], hyper['alpha'],
                                              hyper['beta'], hyper['gamma']))
        if eval_epochs is None or ep + 1 in eval_epochs:
            s = U[test_users].astype(np.float64) @ V.astype(np.float64).T + b.astype(np.float64)[None, :]
            top = scoring.topn_masked(s, train_sets, topn)
            sc = ranking.evaluateCV(yss_true, [r.tolist() for r in top], NAMES, topn)
            hist.append(dict(epoch=ep + 1, TraLoss=float(np.mean(losses)), **dict(zip(NAMES, sc))))
    return hist
