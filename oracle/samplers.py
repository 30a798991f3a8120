"""CPU oracle: the reference's batch samplers, restated without threads/queues.  TEST INFRASTRUCTURE.

Restates the ``__sample_function__`` bodies of
  /root/reference/src/samplers/sampler_ranking.py:22-37      -> (pairs[B,2] int32, negs[B,W] int64)
  /root/reference/src/samplers/sampler_uij_ranking.py:22-38  -> uij[B,3] int64
  /root/reference/src/samplers/sampler_gbpr.py:25-43         -> (+ group[B,G] int64, np.random.choice w/ replacement)
  /root/reference/src/samplers/sampler_rating.py:22-39       -> [B + int(B*negRatio), 3] float64, positives in file order
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
