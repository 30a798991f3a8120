"""On-device samplers: reproducible for a fixed seed, every negative verified against the CSR positives, epoch
structure and dtypes of the reference samplers (SURVEY.md Appendix C invariants)."""
import numpy as np
import pytest

from oracle import samplers as chk

pytestmark = pytest.mark.gpu


def _samplers():
    from collaborativefilteringusingtensorflow_b200.samplers import sampler_gbpr, sampler_ranking, sampler_rating, sampler_uij_ranking
    return sampler_ranking, sampler_uij_ranking, sampler_gbpr, sampler_rating


def test_ranking_sampler_epoch_invariants(ml100k):
    sr = _samplers()[0]
    tra = ml100k['tra']
    s = sr.Sampler(trasR=tra, n_neg=5, batch_size=100, seed=7)
    assert s.batches_per_epoch == 442
    ep = [s.next_batch() for _ in range(442)]
    assert ep[0][0].dtype == np.int32 and ep[0][1].dtype == np.int64
    assert ep[0][0].shape == (100, 2) and ep[0][1].shape == (100, 5)
    pairs, negs = np.concatenate([b[0] for b in ep]), np.concatenate([b[1] for b in ep])
    assert chk.epoch_covers_each_pair_once(tra, pairs, 100)
    assert chk.negatives_are_valid(tra, pairs[:, 0], negs)
    # next epoch: a different permutation, again each pair at most once
    ep2 = np.concatenate([s.next_batch()[0] for _ in range(442)])
    assert chk.epoch_covers_each_pair_once(tra, ep2, 100) and not np.array_equal(ep2, pairs)


def test_same_seed_same_stream_and_chunking_invariance(ml100k):
    sr = _samplers()[0]
    tra = ml100k['tra']
    a, b, c = (sr.Sampler(tra, 3, 128, seed=s) for s in (5, 5, 6))
    ca = a.next_chunk(500)                     # crosses an epoch boundary (345 batches per epoch)
    parts = [b.next_chunk(n) for n in (1, 7, 300, 192)]
    for k in range(2):
        whole = ca[k].cpu().numpy()
        pieces = np.concatenate([p[k].cpu().numpy() for p in parts])
        np.testing.assert_array_equal(whole, pieces)
    assert not np.array_equal(ca[0].cpu().numpy(), c.next_chunk(500)[0].cpu().numpy())
    a.seek(0, 0)
    np.testing.assert_array_equal(a.next_chunk(500)[1].cpu().numpy(), ca[1].cpu().numpy())


def test_negatives_are_uniform_over_the_complement(ml100k):
    sr = _samplers()[0]
    tra = ml100k['tra']
    s = sr.Sampler(tra, 20, 100, seed=3)
    pairs, negs = (x.cpu().numpy() for x in s.next_chunk(442))
    assert chk.negatives_are_valid(tra, pairs[:, 0], negs)
    u = int(np.bincount(pairs[:, 0]).argmax())                      # the most frequent user
    mine = negs[pairs[:, 0] == u].reshape(-1)
    comp = np.setdiff1d(np.arange(tra.shape[1]), np.array(sorted(tra.rows[u])))
    counts = np.bincount(mine, minlength=tra.shape[1])[comp]
    expect = len(mine) / len(comp)
    chi2 = ((counts - expect) ** 2 / expect).sum()
    assert chi2 < len(comp) + 6 * np.sqrt(2 * len(comp))           # chi-square, ~6 sigma
    assert abs(negs.mean() - (tra.shape[1] - 1) / 2) < 40


def test_uij_sampler(ml100k):
    su = _samplers()[1]
    tra = ml100k['tra']
    s = su.Sampler(tra, batch_size=100, seed=1)
    b = s.next_batch()
    assert b.dtype == np.int64 and b.shape == (100, 3)
    assert chk.pairs_are_positives(tra, b[:, :2]) and chk.negatives_are_valid(tra, b[:, 0], b[:, 2:3])


def test_gbpr_sampler(ml100k):
    sg = _samplers()[2]
    tra = ml100k['tra']
    s = sg.Sampler(tra, 3, 5, 100, seed=2)
    p, n, g = s.next_batch()
    assert (p.dtype, n.dtype, g.dtype) == (np.int32, np.int64, np.int64) and g.shape == (100, 3)
    P, N, G = (x.cpu().numpy() for x in s.next_chunk(400))
    assert chk.negatives_are_valid(tra, P[:, 0], N) and chk.group_members_are_valid(tra, P[:, 1], G)
    # with replacement, uniformly over the item's users: for a popular item all of its users show up
    tr_t = tra.transpose().tolil()
    i = int(np.bincount(P[:, 1]).argmax())
    seen = set(G[P[:, 1] == i].reshape(-1).tolist())
    assert seen <= set(tr_t.rows[i]) and len(seen) > 0.5 * min(len(tr_t.rows[i]), (P[:, 1] == i).sum())


def test_rating_sampler(ml100k):
    sr = _samplers()[3]
    u, i, r = ml100k['tra_raw']
    from scipy.sparse import coo_matrix
    raw = coo_matrix((r.astype(np.float32), (u, i)), shape=ml100k['tra'].shape).tolil()
    for mat, want_vals in ((ml100k['tra'], {1.0}), (raw, {1.0, 2.0, 3.0, 4.0, 5.0})):
        s = sr.Sampler(mat, 1, 100, seed=4)
        b0, b1 = s.next_batch(), s.next_batch()
        assert b0.dtype == np.float64 and b0.shape == (200, 3)
        pos = b0[b0[:, 2] > 0]
        assert len(pos) == 100 and set(np.unique(pos[:, 2]).tolist()) <= want_vals
        first100 = chk._pairs_of(mat)[:100]
        assert set(map(tuple, pos[:, :2].astype(int).tolist())) == set(map(tuple, first100.tolist()))   # file order
        neg = b0[b0[:, 2] == 0]
        assert chk.negatives_are_valid(mat, neg[:, 0].astype(int), neg[:, 1:2].astype(int))
        assert not np.array_equal(pos[:, :2].astype(int), first100)        # shuffled inside the batch
        pos1 = b1[b1[:, 2] > 0]
        assert set(map(tuple, pos1[:, :2].astype(int).tolist())) == set(map(tuple, chk._pairs_of(mat)[100:200].tolist()))
    s0 = sr.Sampler(ml100k['tra'], 0.0, 500, seed=4)
    assert s0.next_batch().shape == (500, 3)


def test_degenerate_user_is_flagged_not_spun_on():
    from scipy.sparse import lil_matrix
    sr = _samplers()[0]
    m = lil_matrix((3, 8), dtype=np.float32)
    m[0, :] = 1                                   # user 0 likes everything: the reference would loop forever
    m[1, 2] = 1
    s = sr.Sampler(m, 2, 3, seed=0)
    with pytest.raises(RuntimeError, match='every item'):
        for _ in range(5):
            s.next_batch()


def test_pair_set_membership_gives_the_same_batches_as_the_bisection(monkeypatch):
    """The negatives' rejection test through the hash set of the training pairs (cf_pair_set_build; default for large
    training sets) must reproduce the bisection of the user's CSR row bit for bit: same draws, same answers."""
    import numpy as np
    import torch
    from scipy.sparse import lil_matrix
    from collaborativefilteringusingtensorflow_b200.samplers import _base, sampler_gbpr, sampler_ranking
    rng = np.random.default_rng(11)
    nu, ni = 400, 90                      # dense rows: many rejections
    tra = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        tra[u, rng.choice(ni, size=int(rng.integers(1, 70)), replace=False)] = 1
    outs = {}
    for mode in ('hash', 'bisect'):
        monkeypatch.setattr(_base.DeviceSamplerBase, 'PAIR_SET_MIN_NNZ', 0 if mode == 'hash' else 1 << 40)
        s = sampler_ranking.Sampler(tra, 4, 128, seed=5)
        g = sampler_gbpr.Sampler(tra, 3, 2, 64, seed=6)
        outs[mode] = [t.cpu() for t in s.next_chunk(7)] + [t.cpu() for t in g.next_chunk(5)]
        assert bool(s._pair_set) == (mode == 'hash')
        s.check_flags()
    for a, b in zip(outs['hash'], outs['bisect']):
        assert torch.equal(a, b)
