"""Neighbourhood kernels (cf_neighbors, cf_neighbor_scores, cf_topk_dense) and the ItemCF / UserCF classes against the
reference's own run on ml-100k fold 1 (tests/golden/cf_golden.npz) and against the numpy oracle on random matrices.
Stage by stage, each stage fed with the reference's previous-stage output, because the reference's unstable argsort leaves
ties at a cut undefined (see tests/test_oracle_neighbors.py)."""
import os

import numpy as np
import pytest

from oracle import neighbors as onb

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def cf():
    return np.load(os.path.join(GOLDEN, 'cf_golden.npz'))


def _dense_from(idx, sim, n):
    out = np.zeros((idx.shape[0], n), np.float32)
    r, c = np.nonzero(idx >= 0)
    out[r, idx[r, c]] = sim[r, c]
    return out


def test_similarity_values_bit_identical_to_the_reference(cf, ml100k):
    from collaborativefilteringusingtensorflow_b200 import neighbors
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    tra = DeviceCSR.from_scipy(ml100k['tra'], 'cuda:0', with_values=True)
    for ent, rows, vals, total in ((tra.transpose(), cf['icf_sim_rows'], cf['icf_sim_rows_val'], cf['icf_sim_sum']),
                                   (tra, cf['ucf_sim_rows'], cf['ucf_sim_rows_val'], cf['ucf_sim_sum'])):
        n = ent.shape[0]
        idx, sim = neighbors.cosine_topk(ent, n - 1)                    # every neighbour: the whole similarity matrix
        dense = _dense_from(idx.cpu().numpy(), sim.cpu().numpy(), n)
        assert np.array_equal(dense[rows], vals)                        # float32 bit-exact against the reference's rows
        assert float(dense.astype(np.float64).sum()) == float(total)    # and the whole matrix by its exact fp64 checksum
        s = sim.cpu().numpy()
        assert (s[:, :-1] >= s[:, 1:]).all()                            # ordered by similarity


def test_neighbour_choice_matches_the_reference_up_to_ties(cf, ml100k):
    from collaborativefilteringusingtensorflow_b200 import neighbors
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    tra = DeviceCSR.from_scipy(ml100k['tra'], 'cuda:0', with_values=True)
    idx, sim = (x.cpu().numpy() for x in neighbors.cosine_topk(tra.transpose(), 5))
    assert np.array_equal(sim, cf['icf_nbr_val'])
    clean = ~cf['icf_tie_at_cut']
    assert np.array_equal(np.sort(idx[clean], 1), np.sort(cf['icf_nbr_idx'][clean], 1))
    oi, ov = onb.topk_neighbors(onb.cosine_sim(ml100k['tra'].T.tocsr()), 5)
    assert np.array_equal(idx, oi) and np.array_equal(sim, ov)         # the oracle's tie rule, every row
    uidx, usim = (x.cpu().numpy() for x in neighbors.cosine_topk(tra, 50))
    uclean = ~cf['ucf_tie_at_cut']
    assert np.array_equal(np.sort(uidx[uclean], 1), np.sort(cf['ucf_nbr_idx'][uclean], 1))


def test_scores_from_the_reference_neighbours_are_exact(cf, ml100k):
    import torch
    from collaborativefilteringusingtensorflow_b200 import neighbors
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    tra = DeviceCSR.from_scipy(ml100k['tra'], 'cuda:0', with_values=True)
    users = torch.from_numpy(cf['icf_users8']).cuda()
    got = neighbors.neighbor_scores(tra, users, torch.from_numpy(cf['icf_nbr_idx']).cuda(), torch.from_numpy(cf['icf_nbr_val']).cuda(), 'item')
    assert np.array_equal(got.cpu().numpy(), cf['icf_pred8'])                               # itemcf.py:42-50, float64 exact
    host = ml100k['tra'].tocsr()
    unbr = cf['ucf_nbr_idx']
    usim = onb.cosine_sim(host)
    uval = np.where(unbr >= 0, usim[np.arange(unbr.shape[0])[:, None], np.maximum(unbr, 0)], 0).astype(np.float32)
    got = neighbors.neighbor_scores(tra, users, torch.from_numpy(unbr).cuda(), torch.from_numpy(uval).cuda(), 'user')
    assert np.array_equal(got.cpu().numpy(), cf['ucf_pred8'])                               # usercf.py:31-44


def test_lists_from_the_reference_scores(cf, ml100k):
    import torch
    from collaborativefilteringusingtensorflow_b200 import neighbors
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    tra_host = ml100k['tra']
    tra = DeviceCSR.from_scipy(tra_host, 'cuda:0')
    users8 = cf['icf_users8']
    where = {int(u): k for k, u in enumerate(cf['icf_test_users'])}
    for pred, lists in ((cf['icf_pred8'], cf['icf_lists']), (cf['ucf_pred8'], cf['ucf_lists'])):
        got = neighbors.topk_dense(torch.from_numpy(pred.copy()).cuda(), 10, torch.from_numpy(users8).cuda(), tra).cpu().numpy()
        masks = [set(tra_host.rows[u]) for u in users8]
        assert [list(r) for r in got] == onb.topn_dense(pred, masks, 10)                    # oracle, same tie rule: every list
        for t, u in enumerate(users8):
            ref = [int(x) for x in lists[where[int(u)]] if x >= 0]
            assert [pred[t, j] for j in got[t]] == [pred[t, j] for j in ref]               # reference: same score sequence
            assert not (set(got[t].tolist()) & masks[t])


@pytest.mark.parametrize('cls,key', [('ItemCF', 'icf'), ('UserCF', 'ucf')])
def test_train_end_to_end_close_to_the_reference(cf, ml100k, cls, key):
    import collaborativefilteringusingtensorflow_b200 as pkg
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    m = getattr(pkg, cls)(943, 1682, 5 if key == 'icf' else 50, 10, 'cv', names)            # testicf.py: topK = 5; usercf.py default 50
    got = m.train(1, ml100k['tra'], ml100k['tst'])
    assert np.allclose(got, cf[key + '_scores'], atol=0.01), (got, cf[key + '_scores'])     # ties at the cuts: see the oracle test
    hr = getattr(pkg, cls)(943, 1682, 5 if key == 'icf' else 50, 10, 'loov', ['hr', 'arhr']).train(1, ml100k['tra'], ml100k['tst'])
    assert hr[0] > 0 and hr[1] > 0


@pytest.mark.parametrize('n_rows,n_cols,deg,K,binary', [(40, 30, 6, 5, True), (300, 120, 20, 50, True), (64, 500, 3, 8, False),
                                                        (17, 9, 9, 40, True), (200, 64, 1, 4, True)])
def test_random_matrices_against_the_oracle(n_rows, n_cols, deg, K, binary):
    import torch
    from scipy.sparse import lil_matrix
    from collaborativefilteringusingtensorflow_b200 import neighbors
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    rng = np.random.default_rng(n_rows + n_cols)
    R = lil_matrix((n_rows, n_cols), dtype=np.float32)
    for u in range(n_rows):
        k = int(rng.integers(0, deg + 1))                                    # some empty rows
        if k:
            cols = rng.choice(n_cols, size=min(k, n_cols), replace=False)
            R[u, cols] = 1.0 if binary else rng.integers(1, 6, len(cols)).astype(np.float32)
    csr = DeviceCSR.from_scipy(R, 'cuda:0', with_values=True)
    K = min(K, n_rows, n_cols)
    idx, sim = (x.cpu().numpy() for x in neighbors.cosine_topk(csr, K))
    osim = onb.cosine_sim(R.tocsr())
    oi, ov = onb.topk_neighbors(osim, K)
    if binary:
        assert np.array_equal(idx, oi) and np.array_equal(sim, ov)
    else:   # non-binary values: the fp32 dot products are summed in another order than scipy's
        np.testing.assert_allclose(_dense_from(idx, sim, n_rows), _dense_from(oi, ov, n_rows), rtol=1e-5, atol=1e-7)
    users = torch.arange(n_rows, dtype=torch.int32).cuda()
    for mode, fn, nbr in (('user', onb.user_scores, (oi, ov)),):
        got = neighbors.neighbor_scores(csr, users, torch.from_numpy(nbr[0]).cuda(), torch.from_numpy(nbr[1]).cuda(), mode).cpu().numpy()
        want = fn(R, list(range(n_rows)), nbr[0], nbr[1])
        if binary:
            assert np.array_equal(got, want)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-12)
    ii, iv = onb.topk_neighbors(onb.cosine_sim(R.T.tocsr()), K)
    got = neighbors.neighbor_scores(csr, users, torch.from_numpy(ii).cuda(), torch.from_numpy(iv).cuda(), 'item').cpu().numpy()
    want = onb.item_scores(R, list(range(n_rows)), ii, iv)
    np.testing.assert_allclose(got, want, rtol=1e-12) if not binary else np.testing.assert_array_equal(got, want)
    lists = neighbors.topk_dense(torch.from_numpy(want.copy()).cuda(), min(5, n_cols), users, csr).cpu().numpy()
    masks = [set(R.rows[u]) for u in range(n_rows)]
    ref = onb.topn_dense(want, masks, min(5, n_cols))
    assert [[int(x) for x in r if x >= 0] for r in lists] == ref
