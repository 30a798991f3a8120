"""WRMF weighted-ALS half-sweeps (tensor-core Gram + batched Cholesky) vs a dense float64 solve of the same normal
equations (oracle/als.py); the objective never increases; ml-100k end metric lands in the WRMF ballpark."""
import numpy as np
import pytest

from oracle import als as orc

pytestmark = pytest.mark.gpu


def _rand_matrix(rng, nu, ni, deg):
    from scipy.sparse import lil_matrix
    m = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        k = int(rng.integers(0, deg + 1))
        if k:
            m[u, rng.choice(ni, size=k, replace=False)] = 1
    return m


@pytest.mark.parametrize('nu,ni,d,weight,reg', [(70, 90, 16, 2.0, 0.1), (200, 333, 128, 5.0, 0.5), (64, 1000, 100, 1.0, 0.1),
                                               (300, 50, 20, 40.0, 1.0)])
def test_half_sweeps_match_dense_float64_solve(nu, ni, d, weight, reg):
    from collaborativefilteringusingtensorflow_b200 import WRMF
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    rng = np.random.default_rng(nu + ni)
    R = _rand_matrix(rng, nu, ni, min(ni, 30))
    m = WRMF(nu, ni, weight=weight, reg=reg, n_factors=d, verbose=False, seed=1, solver='als')
    csr = DeviceCSR.from_scipy(R, m.device)
    st = m.state_dict()
    U0, V0 = st['U'].cpu().numpy().astype(np.float64), st['V'].cpu().numpy().astype(np.float64)
    Rd = np.asarray(R.todense(), dtype=np.float64)
    obj0 = orc.objective(U0, V0, Rd, weight, reg)
    m.engine.als_half_sweep('users', csr)
    U1 = m.state_dict()['U'].cpu().numpy()
    want = orc.half_sweep(V0, R.rows, weight, reg)
    np.testing.assert_allclose(U1, want, rtol=2e-3, atol=2e-4 * np.abs(want).max())
    obj1 = orc.objective(U1, V0, Rd, weight, reg)
    m.engine.als_half_sweep('items', csr.transpose())
    V1 = m.state_dict()['V'].cpu().numpy()
    want_v = orc.half_sweep(U1.astype(np.float64), R.transpose().tolil().rows, weight, reg)
    np.testing.assert_allclose(V1, want_v, rtol=2e-3, atol=2e-4 * np.abs(want_v).max())
    obj2 = orc.objective(U1, V1, Rd, weight, reg)
    assert obj1 <= obj0 * (1 + 1e-6) and obj2 <= obj1 * (1 + 1e-6), (obj0, obj1, obj2)


def test_wrmf_als_trains_on_ml100k(ml100k):
    from collaborativefilteringusingtensorflow_b200 import WRMF
    names = ['pre', 'recall', 'map', 'mrr', 'ndcg']
    w = WRMF(943, 1682, 10, 'cv', names, 20., 10., 64, 100, 8, verbose=False, seed=1, solver='als')
    s = w.train(1, ml100k['tra'], ml100k['tst'], None)
    assert s[names.index('ndcg')] > 0.45, s        # the SGD reference path reaches 0.51 after 50 epochs


@pytest.mark.parametrize('P', [1, 3])
def test_row_sharded_sweep_equals_whole_sweep(P):
    """The multi-GPU sweep's stages on one device: partial Grams of P row slices summed (what the all_reduce does) and
    the rows solved range by range (cf_als_gram / cf_als_solve_rows) against cf_als_half_sweep on the whole table."""
    import torch
    from collaborativefilteringusingtensorflow_b200 import WRMF
    from collaborativefilteringusingtensorflow_b200.dist import DistributedALS
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    nu, ni, d = 257, 400, 64
    rng = np.random.default_rng(5)
    R = _rand_matrix(rng, nu, ni, 25)
    a = WRMF(nu, ni, weight=4.0, reg=0.3, n_factors=d, verbose=False, seed=2, solver='als')
    b = WRMF(nu, ni, weight=4.0, reg=0.3, n_factors=d, verbose=False, seed=2, solver='als')
    b.load_state_dict(a.state_dict())
    csr = DeviceCSR.from_scipy(R, a.device)
    a.engine.als_half_sweep('users', csr)
    eng = b.engine
    G = torch.zeros(128, 128, device=eng.device)
    for p in range(P):
        lo, hi = DistributedALS.row_range(ni, P, p)
        eng.als_gram(eng.V[lo:hi], G)
    for p in range(P):
        lo, hi = DistributedALS.row_range(nu, P, p)
        eng.als_solve_rows(eng.U[lo:hi], eng.V, csr.select_rows(torch.arange(lo, hi, device=eng.device)), G)
    got, want = b.state_dict()['U'].cpu().numpy(), a.state_dict()['U'].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * np.abs(want).max())   # fp32 Gram summed in another order
    if P == 1:   # the world-1 driver object runs the same stages
        c = WRMF(nu, ni, weight=4.0, reg=0.3, n_factors=d, verbose=False, seed=2, solver='als')
        c.load_state_dict(a.state_dict())
        a.engine.als_half_sweep('items', csr.transpose())
        DistributedALS(c.engine, csr, csr.transpose()).half_sweep('items')
        np.testing.assert_allclose(c.state_dict()['V'].cpu().numpy(), a.state_dict()['V'].cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('d,weight', [(128, 3.0), (40, 11.0)])
def test_short_rows_low_rank_path_and_long_rows_full_solve_in_one_sweep(d, weight):
    """Every row-length class of the whitened solve (empty; <= 16 and 17..32: one warp per row; 33..128: n x n system from a
    tcgen05 Gram; > 128: 128 x 128 system from a chunked tcgen05 Gram) must equal the dense float64 solve of the same normal
    equations, and so must CF_ALS_DIRECT=1 (the direct 128 x 128 solve of every row)."""
    import os
    from scipy.sparse import lil_matrix
    from collaborativefilteringusingtensorflow_b200 import WRMF
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    rng = np.random.default_rng(d)
    nu, ni, reg = 150, 400, 0.25
    R = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        k = [0, 1, 2, 15, 16, 17, 31, 32, 33, 40, 63, 64, 65, 127, 128, 129, 192, 193, 200, 390][u % 20] if u < 80 else int(rng.integers(0, 150))
        if k:
            R[u, rng.choice(ni, size=k, replace=False)] = 1
    csr = None
    outs = {}
    for mode in ('mixed', 'direct'):
        if mode == 'direct':
            os.environ['CF_ALS_DIRECT'] = '1'
        try:
            m = WRMF(nu, ni, weight=weight, reg=reg, n_factors=d, verbose=False, seed=1, solver='als')
            csr = DeviceCSR.from_scipy(R, m.device)
            V0 = m.state_dict()['V'].cpu().numpy().astype(np.float64)
            m.engine.als_half_sweep('users', csr)
            outs[mode] = m.state_dict()['U'].cpu().numpy()
        finally:
            os.environ.pop('CF_ALS_DIRECT', None)
    want = orc.half_sweep(V0, R.rows, weight, reg)
    for mode in outs:
        np.testing.assert_allclose(outs[mode], want, rtol=2e-3, atol=2e-4 * np.abs(want).max(), err_msg=mode)
    assert np.abs(outs['mixed'][np.diff(R.tocsr().indptr) == 0]).max() == 0.0       # nothing observed: x = 0
