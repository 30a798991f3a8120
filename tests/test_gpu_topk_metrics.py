"""Full-catalog masked top-K (bit-exact indices vs the fp64 oracle, deterministic tie-break) and ranking metrics
(vs the reference-run golden vectors and the oracle, 1e-6)."""
import numpy as np
import pytest

from oracle import ranking as orc_rank
from oracle import scoring

pytestmark = pytest.mark.gpu
NAMES = ['pre', 'recall', 'ndcg', 'map', 'mrr']


def _model(kind, nu, ni, d, seed=3):
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF, WRMF
    cls = dict(bpr=BPRMF, cml=CML, gbpr=GBPRMF, wrmf=WRMF)[kind]
    return cls(nu, ni, n_factors=d, verbose=False, seed=seed)


def _rand_train(rng, nu, ni, max_deg):
    from scipy.sparse import lil_matrix
    m = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        k = int(rng.integers(0, max_deg + 1))
        if k:
            m[u, rng.choice(ni, size=min(k, ni), replace=False)] = 1
    return m


KINDS = [('bpr', scoring.DOT), ('gbpr', scoring.DOT_BIAS), ('cml', scoring.NEG_SQDIST), ('wrmf', scoring.DOT)]


@pytest.mark.parametrize('kind,okind', KINDS)
@pytest.mark.parametrize('nu,ni,d,K', [(40, 300, 20, 10), (64, 9000, 100, 100), (16, 20000, 128, 308), (10, 50, 7, 50)])
def test_topk_bit_exact_vs_oracle(kind, okind, nu, ni, d, K):
    rng = np.random.default_rng(nu + ni + d)
    m = _model(kind, nu, ni, d)
    st = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    # duplicate some item rows so that exact score ties exist (tie -> lower item id first)
    V = st['V']
    V[ni // 2:ni // 2 + 5] = V[3]
    sd = dict(U=st['U'], V=V)
    if kind == 'gbpr':
        b = st['b']
        b[ni // 2:ni // 2 + 5] = b[3]
        sd['b'] = b
    m.load_state_dict(sd)
    tra = _rand_train(rng, nu, ni, min(ni - 1, 60))
    users = rng.permutation(nu)[:max(1, nu // 2)].astype(np.int64)
    got = m.recommend(users, K, tra)
    S = scoring.scores_f64(st['U'][users], V, okind, st.get('b'))
    train_sets = [set(tra.rows[u]) for u in users]
    want = scoring.topn_masked(S, train_sets, K)
    for t in range(len(users)):
        w = [int(x) for x in want[t] if x >= 0]
        assert got[t] == w, 'user %d: first mismatch at %s' % (users[t], next(i for i, (a, c) in enumerate(zip(got[t] + [-9], w + [-8])) if a != c))
    # the dense score matrix itself is bit-identical too
    np.testing.assert_array_equal(m.predict(users), S)
    # and equals the reference's two-stage form
    assert got == scoring.recommend_reference_form(S, train_sets, K)[:len(got)] or K > ni - 60


def test_topk_short_rows_and_item_shards_merge():
    """A user with almost everything masked gets -1 padding; per-shard top-K + merge == single-shot top-K."""
    import torch
    from collaborativefilteringusingtensorflow_b200 import _lib
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    from scipy.sparse import lil_matrix
    rng = np.random.default_rng(0)
    nu, ni, d, K = 12, 5000, 64, 100
    m = _model('bpr', nu, ni, d)
    tra = lil_matrix((nu, ni), dtype=np.float32)
    tra[0, :ni - 7] = 1
    tra[1, rng.choice(ni, 50, replace=False)] = 1
    rec = m.recommend(np.arange(nu), K, tra)
    assert len(rec[0]) == 7 and set(rec[0]) == set(range(ni - 7, ni)) and len(rec[1]) == K
    csr = DeviceCSR.from_scipy(tra, m.device)
    whole_i, whole_v = m.engine.topk(None, K, csr, return_values=True)
    P = 4
    bounds = [0, 1000, 1001, 3500, ni]
    idx = torch.empty(P, nu, K, dtype=torch.int32, device=m.device)
    val = torch.empty(P, nu, K, dtype=torch.float64, device=m.device)
    for p in range(P):
        idx[p], val[p] = m.engine.topk(None, K, csr, return_values=True, item_range=(bounds[p], bounds[p + 1]))
    out_i = torch.empty(nu, K, dtype=torch.int32, device=m.device)
    out_v = torch.empty(nu, K, dtype=torch.float64, device=m.device)
    lib = _lib.lib()
    _lib.check(lib.cf_topk_merge(idx.data_ptr(), val.data_ptr(), P, nu, K, out_i.data_ptr(), out_v.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream), 'merge')
    assert torch.equal(out_i, whole_i) and torch.equal(out_v, whole_v)


def test_metrics_golden(ranking_golden):
    from collaborativefilteringusingtensorflow_b200.metrics import ranking
    for c in ranking_golden['cv']:
        yt = [set(x) for x in c['yss_true']]
        got = ranking.evaluateCV(yt, c['yss_pred'], NAMES, c['k'])
        for n, v in zip(NAMES, got):
            assert abs(v - c['cv'][n]) < 1e-6, (n, v, c['cv'][n])
        assert abs(ranking.precision_k_score(yt, c['yss_pred'], c['k']) - c['cv']['pre']) < 1e-6
        assert abs(ranking.ndcg_k_score(yt, c['yss_pred'], c['k']) - c['cv']['ndcg']) < 1e-6
    for c in ranking_golden['loov']:
        got = ranking.evaluateLOOV(c['ys_true'], c['yss_pred'], ['hr', 'arhr'], c['k'])
        assert abs(got[0] - c['loov']['hr']) < 1e-6 and abs(got[1] - c['loov']['arhr']) < 1e-6
    assert ranking.evaluateCV([{1}], [[1]], ['auc', 'pre'], 1) == [None, 1.0]
    for bad in (([{1}], [[1], [2]], 3), ([], [], 3), ([{1}], [[1]], 0)):
        with pytest.raises(ValueError):
            ranking.recall_k_score(*bad)


def test_metrics_vs_oracle_random_and_duplicates():
    from collaborativefilteringusingtensorflow_b200.metrics import ranking
    rng = np.random.default_rng(12)
    for _ in range(10):
        n_users, n_items, k = int(rng.integers(1, 300)), int(rng.integers(5, 400)), int(rng.integers(1, 60))
        yt = [set(rng.choice(n_items, size=int(rng.integers(1, 5)), replace=False).tolist()) for _ in range(n_users)]
        yp = [rng.integers(0, n_items, size=int(rng.integers(1, k + 5))).tolist() for _ in range(n_users)]   # may repeat ids
        a, b = ranking.evaluateCV(yt, yp, NAMES, k), orc_rank.evaluateCV(yt, yp, NAMES, k)
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)
        ys = rng.integers(0, n_items, n_users).tolist()
        np.testing.assert_allclose(ranking.evaluateLOOV(ys, yp, ['hr', 'arhr'], k),
                                   orc_rank.evaluateLOOV(ys, yp, ['hr', 'arhr'], k), atol=1e-6)
