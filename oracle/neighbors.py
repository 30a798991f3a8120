"""CPU oracle: the neighbourhood models and the user-similarity preprocessing.  TEST INFRASTRUCTURE.

Restates, vectorised in numpy,
  /root/reference/src/models/basic/models/itemcf.py:19-66   (__calsim__, __topk__, __predict__, __recommend)
  /root/reference/src/models/basic/models/usercf.py:19-67
  /root/reference/src/models/pl/models/prigp.py:60-90, cplr_u.py:64-97  (__calsim__, __topk__, __calcoef__)
PINNED against the reference run live on ml-100k fold 1 (oracle/gen_golden.py cf -> tests/golden/cf_golden.npz): the
similarity matrices are bit-identical to the reference's float32 ones; scores and lists are identical wherever numpy's
UNSTABLE argsort (itemcf.py:33,59) leaves no choice, i.e. when no two candidates tie at a cut.  Ties here go to the HIGHER
index first (what a stable ascending argsort followed by [-K:] / [::-1] gives).
"""
import numpy as np


def cosine_sim(R):
    """__calsim__ over the ROWS of the scipy matrix R (pass R.T for itemcf): float32, sim[a, b] = (c / |lo|) / |hi| with
    lo < hi (the order of the in-place row / column divisions, itemcf.py:21-26), zero diagonal."""
    R = R.tocsr().astype(np.float32)
    c = np.asarray((R @ R.T).todense(), dtype=np.float32)
    den = np.sqrt(np.asarray(R.multiply(R).sum(axis=1), dtype=np.float32).reshape(-1)).astype(np.float32)
    n = R.shape[0]
    safe = np.where(den > 0, den, np.float32(1)).astype(np.float32)
    lo = np.minimum.outer(np.arange(n), np.arange(n))
    hi = np.maximum.outer(np.arange(n), np.arange(n))
    sim = ((c / safe[lo]).astype(np.float32) / safe[hi]).astype(np.float32)
    np.fill_diagonal(sim, 0)
    return sim


def topk_neighbors(sim, K, tie_high=True):
    """__topk__ (itemcf.py:29-40): per row the K largest POSITIVE similarities, (sim desc, index desc | asc);
    returns (idx int32 [n, K] padded with -1, val float32 [n, K])."""
    n = sim.shape[0]
    K = min(K, n)
    idx = np.full((n, K), -1, dtype=np.int32)
    val = np.zeros((n, K), dtype=np.float32)
    cols = np.arange(sim.shape[1])
    for a in range(n):
        row = sim[a]
        order = np.lexsort((-cols if tie_high else cols, -row))       # primary: -sim ascending; ties: index order
        order = order[row[order] > 0][:K]
        idx[a, :len(order)] = order
        val[a, :len(order)] = row[order]
    return idx, val


def item_scores(R, users, nbr_idx, nbr_sim):
    """itemcf.py:42-50: score[u, j] = sum_{i in items(u)} simK[i, j] * r_ui, float32 products accumulated in float64."""
    R = R.tocsr()
    out = np.zeros((len(users), R.shape[1]), dtype=np.float64)
    for t, u in enumerate(users):
        for e in range(R.indptr[u], R.indptr[u + 1]):
            i, r = R.indices[e], np.float32(R.data[e])
            ok = nbr_idx[i] >= 0
            np.add.at(out[t], nbr_idx[i][ok], (nbr_sim[i][ok] * r).astype(np.float32).astype(np.float64))
    return out


def user_scores(R, users, nbr_idx, nbr_sim):
    """usercf.py:31-44: score[u, :] = sum over the neighbours v with sim > 0 of sim(u, v) * R[v, :]."""
    R = R.tocsr()
    out = np.zeros((len(users), R.shape[1]), dtype=np.float64)
    for t, u in enumerate(users):
        for v, s in zip(nbr_idx[u], nbr_sim[u]):
            if v >= 0 and s > 0:
                sl = slice(R.indptr[v], R.indptr[v + 1])
                out[t, R.indices[sl]] += (np.float32(s) * R.data[sl].astype(np.float32)).astype(np.float32).astype(np.float64)
    return out


def topn_dense(scores, mask_sets, N, tie_high=True):
    """itemcf.py:52-66: the N best unmasked columns by (score desc, index desc | asc)."""
    out = []
    cols = np.arange(scores.shape[1])
    for t in range(scores.shape[0]):
        order = np.lexsort((-cols if tie_high else cols, -scores[t]))
        out.append([int(j) for j in order if j not in mask_sets[t]][:N])
    return out


def coef_matrix(R, sim_topk_dense, weighted):
    """__calcoef__: prigp.py:83-90 (weighted=False: number of top-K neighbours that hold the item) / cplr_u.py:89-97
    (weighted=True: their similarities summed), float64 dense [n_users, n_items]; rows of users without training items stay 0."""
    R = R.tocsr()
    B = (R > 0).astype(np.float64) if not weighted else R.astype(np.float64)
    W = (sim_topk_dense != 0).astype(np.float64) if not weighted else sim_topk_dense.astype(np.float64)
    out = np.asarray(W @ B.todense())
    has = np.diff(R.indptr) > 0
    out[~has] = 0
    return out
