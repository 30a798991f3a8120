"""CPU oracle: the reference's batch samplers, restated without threads/queues.  TEST INFRASTRUCTURE.

Restates the ``__sample_function__`` bodies of
  /root/reference/src/samplers/sampler_ranking.py:22-37      -> (pairs[B,2] int32, negs[B,W] int64)
  /root/reference/src/samplers/sampler_uij_ranking.py:22-38  -> uij[B,3] int64
  /root/reference/src/samplers/sampler_gbpr.py:25-43         -> (+ group[B,G] int64, np.random.choice w/ replacement)
  /root/reference/src/samplers/sampler_rating.py:22-39       -> [B + int(B*negRatio), 3] float64, positives in file order
  /root/reference/src/samplers/sampler_prigp.py:22-52        -> uijtk[B,5] int64 (collaborative pair t, k from the coefficient rows)
  /root/reference/src/samplers/sampler_uitj_ranking.py:22-40 -> (uitj[B,4] int64, coefs[B,2] float64)
as plain generators over a seeded ``numpy.random.Generator`` (the reference is unseeded and runs in a
producer thread; distribution, shapes, dtypes and the epoch structure are what is restated).

Also the invariant checkers used on the on-device sampler's output.
PINNED on invariants/dtypes against the live reference samplers (tests/golden/sampler_golden.json).
"""
import numpy as np


def _pairs_of(trasR):
    coo = trasR.tocoo()
    order = np.lexsort((coo.col, coo.row))
    return np.stack([coo.row[order], coo.col[order]], axis=1)      # == np.array(trasR.nonzero()).T for lil


def _posmask_fn(trasR):
    csr = trasR.tocsr()
    csr.sort_indices()
    indptr, indices = csr.indptr, csr.indices
    allkeys = np.repeat(np.arange(csr.shape[0], dtype=np.int64), np.diff(indptr)) * csr.shape[1] + indices
    if len(allkeys) == 0:
        allkeys = np.array([-1], dtype=np.int64)

    def is_pos(u, j):
        u = np.asarray(u, dtype=np.int64)
        j = np.asarray(j, dtype=np.int64)
        key = u * csr.shape[1] + j
        pos = np.searchsorted(allkeys, key)
        pos = np.minimum(pos, len(allkeys) - 1)
        return allkeys[pos] == key
    return is_pos


def _draw_negs(rng, is_pos, users, n_items, shape):
    negs = rng.integers(0, n_items, size=shape)
    ub = np.broadcast_to(np.asarray(users).reshape((-1,) + (1,) * (len(shape) - 1)), shape)
    bad = is_pos(ub.reshape(-1), negs.reshape(-1)).reshape(shape)
    while bad.any():                                                # sampler_ranking.py:35-36
        negs[bad] = rng.integers(0, n_items, size=int(bad.sum()))
        bad2 = np.zeros_like(bad)
        bad2[bad] = is_pos(ub[bad], negs[bad])
        bad = bad2
    return negs


def ranking_batches(trasR, n_neg=5, batch_size=100, seed=0):
    rng = np.random.default_rng(seed)
    pairs = _pairs_of(trasR).astype(np.int32)
    is_pos = _posmask_fn(trasR)
    n_items = trasR.shape[1]
    while True:
        rng.shuffle(pairs)                                          # :24
        for k in range(int(len(pairs) / batch_size)):               # :25 (tail dropped)
            pb = pairs[k * batch_size:(k + 1) * batch_size]
            yield pb.copy(), _draw_negs(rng, is_pos, pb[:, 0], n_items, (len(pb), n_neg)).astype(np.int64)


def uij_batches(trasR, batch_size=100, seed=0):
    for pb, nb in ranking_batches(trasR, 1, batch_size, seed):
        yield np.concatenate([pb.astype(np.int64), nb], axis=1)


def gbpr_batches(trasR, gsize=2, n_neg=5, batch_size=100, seed=0):
    rng = np.random.default_rng(seed + 7919)
    csc = trasR.tocsc()
    csc.sort_indices()
    for pb, nb in ranking_batches(trasR, n_neg, batch_size, seed):
        items = pb[:, 1].astype(np.int64)
        deg = (csc.indptr[items + 1] - csc.indptr[items])
        pick = (rng.random((len(pb), gsize)) * deg[:, None]).astype(np.int64)
        group = csc.indices[csc.indptr[items][:, None] + pick].astype(np.int64)   # sampler_gbpr.py:41
        yield pb, nb, group


def rating_batches(trasR, negRatio=0.0, batch_size=500, seed=0):
    rng = np.random.default_rng(seed)
    pairs = _pairs_of(trasR)
    csr = trasR.tocsr()
    vals = np.asarray(csr[pairs[:, 0], pairs[:, 1]]).reshape(-1)
    uir = np.concatenate([pairs.astype(np.float64), vals[:, None].astype(np.float64)], axis=1)
    is_pos = _posmask_fn(trasR)
    n_users, n_items = trasR.shape
    num_neg = int(batch_size * negRatio)
    while True:
        for k in range(int(len(uir) / batch_size)):                 # sampler_rating.py:24 -- never shuffled globally
            batch = uir[k * batch_size:(k + 1) * batch_size]
            if num_neg > 0:
                users = rng.integers(0, n_users, size=num_neg)
                negs = _draw_negs(rng, is_pos, users, n_items, (num_neg,))
                neg_rows = np.stack([users, negs, np.zeros(num_neg)], axis=1).astype(np.float64)
                batch = np.concatenate([batch, neg_rows])
            batch = batch.copy()
            rng.shuffle(batch)                                      # :38 in-batch shuffle
            yield batch


def _coef_rows(coefMat):
    """The coefficient matrix (prigp.py:83-90 / cplr_u.py:89-97) as sorted CSR arrays without stored zeros."""
    csr = coefMat.tocsr().astype(np.float64)
    csr.eliminate_zeros()
    csr.sort_indices()
    return csr


def _choice_in_rows(rng, indptr, indices, users):
    """np.random.choice(list(row_set[u]), 1)[0] for every u: uniform over the row's stored columns; also the CSR slot."""
    lo = indptr[users].astype(np.int64)
    n = (indptr[users + 1] - indptr[users]).astype(np.int64)
    slot = lo + np.minimum((rng.random(len(users)) * n).astype(np.int64), n - 1)
    return indices[slot].astype(np.int64), slot


def prigp_batches(trasR, coefMat, batch_size=100, seed=0):
    """sampler_prigp.py:22-52.  Per epoch the positives are shuffled and cut into whole batches (:24-25); per row: j uniform
    outside the user's positives (:30-34); without coefficients (t, k) = (i, j) (:36); otherwise t uniform in the user's
    coefficient row (:38) and k uniform outside it (:39-41); when the row holds more than one DISTINCT value and a standard
    normal draw is below nnz(row) / n_items (:43 -- np.random.randn, not rand, so the branch is taken with probability
    Phi(nnz / n_items) >= 1/2), k is redrawn inside the row until its coefficient differs from t's (:44-46) and the pair is
    ordered so that t carries the larger coefficient (:47-48)."""
    rng = np.random.default_rng(seed)
    pairs = _pairs_of(trasR).astype(np.int64)
    is_pos = _posmask_fn(trasR)
    coef = _coef_rows(coefMat)
    in_coef = _posmask_fn(coef)
    indptr, indices, vals = coef.indptr, coef.indices, coef.data
    n_items = trasR.shape[1]
    nnz_row = np.diff(indptr)
    distinct = np.array([len(set(vals[indptr[u]:indptr[u + 1]])) for u in range(coef.shape[0])])   # user_coefItemset_vals
    while True:
        rng.shuffle(pairs)
        for b in range(int(len(pairs) / batch_size)):
            pb = pairs[b * batch_size:(b + 1) * batch_size]
            u, i = pb[:, 0], pb[:, 1]
            j = _draw_negs(rng, is_pos, u, n_items, (len(pb),)).astype(np.int64)
            t, k = i.copy(), j.copy()
            has = np.nonzero(nnz_row[u] > 0)[0]
            if len(has):
                uh = u[has]
                th, slot_t = _choice_in_rows(rng, indptr, indices, uh)
                kh = _draw_negs(rng, in_coef, uh, n_items, (len(has),)).astype(np.int64)
                inside = (distinct[uh] > 1) & (rng.standard_normal(len(has)) < nnz_row[uh] / float(n_items))
                todo = np.nonzero(inside)[0]
                slot_k = np.zeros(len(has), dtype=np.int64)
                while len(todo):                                                   # :44-46
                    kh[todo], slot_k[todo] = _choice_in_rows(rng, indptr, indices, uh[todo])
                    todo = todo[vals[slot_t[todo]] == vals[slot_k[todo]]]
                swap = inside & (vals[slot_t] < vals[np.where(inside, slot_k, slot_t)])
                th[swap], kh[swap] = kh[swap], th[swap]
                t[has], k[has] = th, kh
            yield np.stack([u, i, j, t, k], axis=1)


def uitj_batches(trasR, coefMat, batch_size=1000, seed=0):
    """sampler_uitj_ranking.py:22-40.  ui[u] = the positives, ut[u] = coefficient columns that are not positives (:12-13).
    Every row draws its user uniformly among those with positives, collaborative items and room for a negative (:27-28),
    i uniform in ui[u], t uniform in ut[u], j uniform outside both (:29-33); coefs = (coef[u, i], coef[u, t]) (:35)."""
    rng = np.random.default_rng(seed)
    n_users, n_items = trasR.shape
    tra = trasR.tocsr().astype(np.float64)
    tra.eliminate_zeros()
    tra.sort_indices()
    coef = _coef_rows(coefMat)
    collab = coef - coef.multiply(tra != 0)                # ut: the coefficient entries outside the positives
    collab = collab.tocsr()
    collab.eliminate_zeros()
    collab.sort_indices()
    is_pos, is_collab = _posmask_fn(tra), _posmask_fn(collab)
    npos, ncol = np.diff(tra.indptr), np.diff(collab.indptr)
    eligible = np.nonzero((npos > 0) & (ncol > 0) & (npos + ncol < n_items))[0]
    dense_lookup = coef.todok() if n_users * n_items > 50_000_000 else None
    coef_dense = None if dense_lookup is not None else coef.toarray()
    taken = lambda uu, jj: is_pos(uu, jj) | is_collab(uu, jj)
    while True:
        u = eligible[(rng.random(batch_size) * len(eligible)).astype(np.int64)]
        i, _ = _choice_in_rows(rng, tra.indptr, tra.indices, u)
        t, _ = _choice_in_rows(rng, collab.indptr, collab.indices, u)
        j = _draw_negs(rng, taken, u, n_items, (batch_size,)).astype(np.int64)
        if coef_dense is not None:
            c = np.stack([coef_dense[u, i], coef_dense[u, t]], axis=1)
        else:
            c = np.array([[dense_lookup.get((a, b), 0.0), dense_lookup.get((a, d), 0.0)] for a, b, d in zip(u, i, t)])
        yield np.stack([u, i, t, j], axis=1).astype(np.int64), c.astype(np.float64)


# ---------------------------------------------------------------- invariant checkers
def negatives_are_valid(trasR, users, negs):
    """Every sampled negative is outside the user's training set and inside [0, n_items)."""
    is_pos = _posmask_fn(trasR)
    users = np.asarray(users).reshape(-1, 1)
    negs = np.asarray(negs).reshape(len(users), -1)
    ub = np.broadcast_to(users, negs.shape)
    ok_range = (negs >= 0).all() and (negs < trasR.shape[1]).all()
    return bool(ok_range and not is_pos(ub.reshape(-1), negs.reshape(-1)).any())


def pairs_are_positives(trasR, pairs):
    is_pos = _posmask_fn(trasR)
    return bool(is_pos(pairs[:, 0], pairs[:, 1]).all())


def epoch_covers_each_pair_once(trasR, epoch_pairs, batch_size):
    """One epoch = B*int(nnz/B) DISTINCT training pairs (sampler_ranking.py:24-27)."""
    nnz = trasR.nnz
    want = batch_size * int(nnz / batch_size)
    keys = epoch_pairs[:, 0].astype(np.int64) * trasR.shape[1] + epoch_pairs[:, 1].astype(np.int64)
    return len(keys) == want and len(np.unique(keys)) == want and pairs_are_positives(trasR, epoch_pairs)


def group_members_are_valid(trasR, items, group):
    """Every group member has the pair's positive item in their training set (sampler_gbpr.py:15,41)."""
    is_pos = _posmask_fn(trasR)
    ib = np.broadcast_to(np.asarray(items).reshape(-1, 1), group.shape)
    return bool(is_pos(group.reshape(-1), ib.reshape(-1)).all())


def tuple_sampler_stats(tra, coef, batches, kind):
    """Distribution statistics of PRIGP (uijtk) / CPLR (uitj, coefs) batches: what tests compare between the reference's
    samplers run live, the oracle's restatement and the device samplers.  ``coef``: dense float64 [n_users, n_items]."""
    pos = np.asarray(tra.todense()) > 0
    nz = coef != 0
    ni = tra.shape[1]
    if kind == 'prigp':
        b = np.concatenate(batches)
        u, i, j, t, k = b.T
        has = nz.sum(1)[u] > 0
        distinct = np.array([len(set(row[row != 0])) for row in coef])
        can = has & (distinct[u] > 1)
        inside = has & nz[u, k]
        return dict(rows=int(len(b)), dtype=str(b.dtype), shape=list(batches[0].shape),
                    pairs_positive=bool(pos[u, i].all()), j_not_positive=bool(not pos[u, j].any()),
                    no_coef_rows_copy_ij=bool(np.array_equal(t[~has], i[~has]) and np.array_equal(k[~has], j[~has])),
                    t_in_coef_row=bool(nz[u[has], t[has]].all()),
                    inside_pairs_ordered=bool((coef[u[inside], t[inside]] > coef[u[inside], k[inside]]).all()),
                    distinct_pairs=int(len(set(zip(u.tolist(), i.tolist())))),
                    frac_rows_with_coef=float(has.mean()), inside_rate=float(inside[can].mean()),
                    mean_coef_t_inside=float(coef[u[inside], t[inside]].mean()), mean_coef_k_inside=float(coef[u[inside], k[inside]].mean()),
                    mean_coef_t_outside=float(coef[u[has & ~inside], t[has & ~inside]].mean()),
                    frac_t_is_positive=float(pos[u[has], t[has]].mean()), mean_j=float(j.mean()) / ni,
                    mean_k_outside=float(k[has & ~inside].mean()) / ni)
    uitj = np.concatenate([x[0] for x in batches])
    c = np.concatenate([x[1] for x in batches])
    u, i, t, j = uitj.T
    npos, ncol = pos.sum(1), (nz & ~pos).sum(1)
    eligible = (npos > 0) & (ncol > 0) & (npos + ncol < ni)
    return dict(rows=int(len(uitj)), dtype=str(uitj.dtype), coef_dtype=str(c.dtype), shape=list(batches[0][0].shape),
                coef_shape=list(batches[0][1].shape), i_positive=bool(pos[u, i].all()),
                t_collaborative=bool((nz[u, t] & ~pos[u, t]).all()), j_outside_both=bool(not (pos[u, j] | nz[u, j]).any()),
                coefs_are_matrix_entries=bool(np.array_equal(c[:, 0], coef[u, i]) and np.array_equal(c[:, 1], coef[u, t])),
                users_all_eligible=bool(eligible[u].all()), n_eligible=int(eligible.sum()),
                distinct_users=int(len(np.unique(u))), mean_user_degree=float(npos[u].mean()),
                mean_coef_i=float(c[:, 0].mean()), mean_coef_t=float(c[:, 1].mean()), mean_j=float(j.mean()) / ni)


def pair_sampler_stats(tra, kind, batches):
    """Distribution statistics of the hot-path samplers' batches (a16, a18, a19): what tests compare between the reference's
    samplers run live, the oracle's restatements and the device samplers.  ``batches`` as next_batch() returns them."""
    nu, ni = tra.shape
    pos = np.asarray(tra.todense()) > 0
    deg_u, deg_i = pos.sum(1), pos.sum(0)
    if kind in ('ranking', 'gbpr'):
        pairs = np.concatenate([np.asarray(b[0]) for b in batches]).astype(np.int64)
        negs = np.concatenate([np.asarray(b[1]) for b in batches]).astype(np.int64)
        u, i = pairs[:, 0], pairs[:, 1]
        out = dict(rows=int(len(pairs)), pairs_positive=bool(pos[u, i].all()), negatives_valid=bool(not pos[u[:, None], negs].any()),
                   distinct_pairs=int(len(set(zip(u.tolist(), i.tolist())))),
                   mean_neg=float(negs.mean()) / ni, frac_neg_low_half=float((negs < ni // 2).mean()),
                   mean_neg_popularity=float(deg_i[negs].mean()), mean_pair_user_degree=float(deg_u[u].mean()))
        if kind == 'gbpr':
            grp = np.concatenate([np.asarray(b[2]) for b in batches]).astype(np.int64)
            out.update(group_valid=bool(pos[grp, i[:, None]].all()), frac_group_is_user=float((grp == u[:, None]).mean()),
                       expected_frac_group_is_user=float((1.0 / deg_i[i]).mean()), mean_group_user_degree=float(deg_u[grp].mean()))
        return out
    rows = np.concatenate([np.asarray(b) for b in batches])
    uu, ii, rr = rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64), rows[:, 2]
    neg = rr == 0
    return dict(rows=int(len(rows)), batch_rows=int(len(batches[0])), positives_per_batch=int((np.asarray(batches[0])[:, 2] > 0).sum()),
                positives_positive=bool(pos[uu[~neg], ii[~neg]].all()), negatives_valid=bool(not pos[uu[neg], ii[neg]].any()),
                distinct_positive_pairs=int(len(set(zip(uu[~neg].tolist(), ii[~neg].tolist())))),
                mean_neg_user=float(uu[neg].mean()) / nu, mean_neg_item=float(ii[neg].mean()) / ni,
                mean_neg_user_degree=float(deg_u[uu[neg]].mean()), mean_neg_item_popularity=float(deg_i[ii[neg]].mean()))


# tolerances for comparing two samples' pair_sampler_stats: >= 4 standard errors of the DIFFERENCE of two samples of one
# ml-100k epoch (220 500 negatives) / 200 rating batches (20 000 negatives; item popularity there has std 41 -> SE 0.29 per
# sample, user degree std 44 -> SE 0.31)
PAIR_STATS_TOL = dict(mean_neg=0.005, frac_neg_low_half=0.006, mean_neg_popularity=0.5, mean_pair_user_degree=1.0,
                      frac_group_is_user=0.003, expected_frac_group_is_user=0.001, mean_group_user_degree=1.0,
                      mean_neg_user=0.012, mean_neg_item=0.012, mean_neg_user_degree=2.0, mean_neg_item_popularity=1.8)


def compare_pair_stats(got, want, scale=1.0):
    """None if ``got`` agrees with ``want`` (booleans and per-batch counts equal, every row a distinct positive pair,
    distribution statistics within ``scale`` x PAIR_STATS_TOL), else a description of the first difference."""
    for k, v in want.items():
        if k in ('rows', 'batches_used', 'distinct_pairs', 'distinct_positive_pairs'):
            continue
        if isinstance(v, float):
            if abs(got[k] - v) > scale * PAIR_STATS_TOL[k]:
                return '%s: %r vs %r' % (k, got[k], v)
        elif got[k] != v:
            return '%s: %r vs %r' % (k, got[k], v)
    if 'distinct_pairs' in got and got['distinct_pairs'] != got['rows']:
        return 'an epoch repeats a pair'
    return None
