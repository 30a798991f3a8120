"""oracle.scoring (two forms agree) and oracle.samplers (invariants; reference-sampler facts from the golden file)."""
import json
import os

import numpy as np

from oracle import samplers, scoring
from conftest import GOLDEN


def test_masked_topn_equals_reference_two_stage_form():
    rng = np.random.default_rng(3)
    U, V = rng.standard_normal((17, 12)).astype(np.float32), rng.standard_normal((40, 12)).astype(np.float32)
    V[7] = V[3]                                    # exact tie -> lower index first
    b = rng.standard_normal(40).astype(np.float32)
    train = [set(rng.choice(40, size=int(rng.integers(0, 9)), replace=False).tolist()) for _ in range(17)]
    for kind in (scoring.DOT, scoring.DOT_BIAS, scoring.NEG_SQDIST):
        s = scoring.scores_f64(U, V, kind, b)
        a = scoring.topn_masked(s, train, 10)
        r = scoring.recommend_reference_form(s, train, 10)
        assert [row.tolist() for row in a] == r
    s = scoring.scores_f64(U, V)
    assert np.allclose(s, U.astype(np.float64) @ V.astype(np.float64).T, atol=1e-12)
    top = scoring.topn_masked(s, [set()] * 17, 40)
    for t in range(17):
        assert list(top[t]).index(3) + 1 == list(top[t]).index(7)


def test_reference_sampler_facts():
    g = json.load(open(os.path.join(GOLDEN, 'sampler_golden.json')))
    assert g['ranking']['pairs_dtype'] == 'int32' and g['ranking']['negs_dtype'] == 'int64'
    assert g['ranking']['batches_per_epoch'] == 442 and g['ranking']['negatives_valid'] and g['ranking']['pairs_positive']
    # the reference's view/shuffle race only ever hits the tail of an epoch
    assert all(k >= 440 for k in g['ranking']['race_hit_batches'])
    assert g['uij'] == dict(dtype='int64', shape=[100, 3], negatives_valid=True)
    assert g['gbpr']['group_dtype'] == 'int64' and g['gbpr']['group_valid'] and g['gbpr']['negatives_valid']
    assert g['rating']['dtype'] == 'float64' and g['rating']['shape'] == [200, 3] and g['rating']['n_pos'] == 100
    assert g['rating']['positives_in_file_order'] and g['rating']['neg_valid']


def test_oracle_samplers_invariants(ml100k):
    tra = ml100k['tra']
    assert tra.nnz == ml100k['stats']['tra_pos'] == 44243
    gen = samplers.ranking_batches(tra, 5, 100, seed=1)
    nb = int(tra.nnz / 100)
    ep = [next(gen) for _ in range(nb)]
    pairs, negs = np.concatenate([b[0] for b in ep]), np.concatenate([b[1] for b in ep])
    assert ep[0][0].dtype == np.int32 and ep[0][1].dtype == np.int64 and ep[0][1].shape == (100, 5)
    assert samplers.epoch_covers_each_pair_once(tra, pairs, 100)
    assert samplers.negatives_are_valid(tra, pairs[:, 0], negs)
    p, n, g = next(samplers.gbpr_batches(tra, 3, 5, 100, seed=2))
    assert g.shape == (100, 3) and samplers.group_members_are_valid(tra, p[:, 1], g)
    b = next(samplers.rating_batches(tra, 1, 100, seed=3))
    assert b.shape == (200, 3) and b.dtype == np.float64 and (b[:, 2] > 0).sum() == 100
    assert set(map(tuple, b[b[:, 2] > 0][:, :2].astype(int).tolist())) == set(map(tuple, samplers._pairs_of(tra)[:100].tolist()))
    u = next(samplers.uij_batches(tra, 100, seed=4))
    assert u.shape == (100, 3) and u.dtype == np.int64


def test_oracle_samplers_match_the_reference_samplers_run_live(ml100k):
    """tests/golden/pair_sampler_stats_golden.json: sampler_ranking / sampler_gbpr / sampler_rating of the reference run live
    (oracle/gen_golden.py sampler-stats): one epoch each (W = 5, B = 100; G = 3) and 200 rating batches at negRatio 1.  The
    oracle's restatements draw from the same distributions: negatives uniform over the non-positives (mean id, share in the
    lower half of the catalogue, mean popularity), every positive once per epoch (so the pairs' mean user degree is
    degree-weighted), group members uniform over the item's users WITH replacement and including the user itself at rate
    mean(1 / deg(i)), rating negatives uniform over users (not degree-weighted)."""
    gold = json.load(open(os.path.join(GOLDEN, 'pair_sampler_stats_golden.json')))
    tra = ml100k['tra']
    nb = int(tra.nnz / 100)
    gen = samplers.ranking_batches(tra, 5, 100, seed=31)
    assert samplers.compare_pair_stats(samplers.pair_sampler_stats(tra, 'ranking', [next(gen) for _ in range(nb)]), gold['ranking']) is None
    gen = samplers.gbpr_batches(tra, 3, 5, 100, seed=32)
    assert samplers.compare_pair_stats(samplers.pair_sampler_stats(tra, 'gbpr', [next(gen) for _ in range(nb)]), gold['gbpr']) is None
    gen = samplers.rating_batches(tra, 1, 100, seed=33)
    assert samplers.compare_pair_stats(samplers.pair_sampler_stats(tra, 'rating', [next(gen) for _ in range(200)]), gold['rating']) is None
    # the check has teeth: degree-weighted negative users (a plausible mis-restatement of sampler_rating.py:31) are caught
    bad = dict(gold['rating'], mean_neg_user_degree=float((np.asarray(tra.todense()) > 0).sum(1).astype(float).dot(
        (np.asarray(tra.todense()) > 0).sum(1)) / tra.nnz))
    assert samplers.compare_pair_stats(samplers.pair_sampler_stats(tra, 'rating', [next(gen) for _ in range(200)]), bad) is not None
