"""The numpy oracle of the neighbourhood models (oracle/neighbors.py) against the reference's own ItemCF / UserCF run live
(tests/golden/cf_golden.npz, written by oracle/gen_golden.py cf).  The reference's argsort is unstable, so lists are only
defined where nothing ties at a cut; every stage is checked on the reference's own previous-stage output."""
import os

import numpy as np
import pytest

from oracle import neighbors as onb
from oracle import ranking as orank

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def cf():
    return np.load(os.path.join(GOLDEN, 'cf_golden.npz'))


def test_similarities_are_bit_identical_to_the_reference(cf, ml100k):
    tra = ml100k['tra']
    isim = onb.cosine_sim(tra.T.tocsr())
    assert isim.dtype == np.float32 and np.array_equal(isim[cf['icf_sim_rows']], cf['icf_sim_rows_val'])
    assert float(isim.astype(np.float64).sum()) == float(cf['icf_sim_sum'])
    assert float((isim.astype(np.float64) ** 2).sum()) == float(cf['icf_sim_sqsum'])
    usim = onb.cosine_sim(tra.tocsr())
    assert np.array_equal(usim[cf['ucf_sim_rows']], cf['ucf_sim_rows_val']) and float(usim.astype(np.float64).sum()) == float(cf['ucf_sim_sum'])


def test_neighbour_choice_matches_the_reference_up_to_ties(cf, ml100k):
    tra = ml100k['tra']
    idx, val = onb.topk_neighbors(onb.cosine_sim(tra.T.tocsr()), 5)
    assert np.array_equal(val, cf['icf_nbr_val'])                       # the kept similarity VALUES never depend on the tie order
    clean = ~cf['icf_tie_at_cut']
    assert np.array_equal(np.sort(idx[clean], 1), np.sort(cf['icf_nbr_idx'][clean], 1))
    assert clean.sum() > 1000
    uidx, uval = onb.topk_neighbors(onb.cosine_sim(tra.tocsr()), 50)
    uclean = ~cf['ucf_tie_at_cut']
    assert np.array_equal(np.sort(uidx[uclean], 1), np.sort(cf['ucf_nbr_idx'][uclean], 1)) and uclean.sum() > 600


def test_scores_and_lists_from_the_reference_neighbours(cf, ml100k):
    tra = ml100k['tra']
    users8 = cf['icf_users8']
    pred = onb.item_scores(tra, users8, cf['icf_nbr_idx'], cf['icf_nbr_val'])
    assert np.array_equal(pred, cf['icf_pred8'])
    where = {int(u): k for k, u in enumerate(cf['icf_test_users'])}
    masks = [set(tra.rows[u]) for u in users8]
    mine = onb.topn_dense(pred, masks, 10)
    for t, u in enumerate(users8):
        ref = [int(x) for x in cf['icf_lists'][where[int(u)]] if x >= 0]
        assert [pred[t, j] for j in mine[t]] == [pred[t, j] for j in ref]       # same score sequence, whatever the tie order
        top11 = np.sort(pred[t][[j for j in range(pred.shape[1]) if j not in masks[t]]])[::-1][:11]
        if len(set(top11.tolist())) == 11:
            assert mine[t] == ref                                                # no tie near the cut: the same list


def test_end_to_end_metrics_are_close_to_the_reference(cf, ml100k):
    """Ties at the cuts (463 of 1682 item rows, 225 of 943 user rows on this fold) change which of several equally similar
    neighbours is kept, so the end metrics agree to a few 1e-3, not to the last digit."""
    tra, tst = ml100k['tra'], ml100k['tst']
    test_users = [int(u) for u in cf['icf_test_users']]
    truth = [set(tst.rows[u]) for u in test_users]
    masks = [set(tra.rows[u]) for u in test_users]
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    idx, val = onb.topk_neighbors(onb.cosine_sim(tra.T.tocsr()), 5)
    got = orank.evaluateCV(truth, onb.topn_dense(onb.item_scores(tra, test_users, idx, val), masks, 10), names, 10)
    assert np.allclose(got, cf['icf_scores'], atol=0.01), (got, cf['icf_scores'])
    uidx, uval = onb.topk_neighbors(onb.cosine_sim(tra.tocsr()), 50)
    got = orank.evaluateCV(truth, onb.topn_dense(onb.user_scores(tra, test_users, uidx, uval), masks, 10), names, 10)
    assert np.allclose(got, cf['ucf_scores'], atol=0.01), (got, cf['ucf_scores'])
