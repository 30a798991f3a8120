"""Tensor-core top-K (tcgen05 / TMA candidate pass + exact re-rank) == the exact fp64 kernel, bit for bit, and its raw
fp16 GEMM scores match a torch matmul of the fp16-rounded operands."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(kind, nu, ni, d, seed=3, std=0.1):
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF
    cls = dict(bpr=BPRMF, cml=CML, gbpr=GBPRMF)[kind]
    return cls(nu, ni, n_factors=d, verbose=False, seed=seed, init_stddev=std)


def _train_csr(rng, nu, ni, deg, device):
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    from scipy.sparse import lil_matrix
    m = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        k = int(rng.integers(0, deg + 1))
        if k:
            m[u, rng.choice(ni, size=k, replace=False)] = 1
    return DeviceCSR.from_scipy(m, device)


@pytest.mark.parametrize('d', [128, 100, 64])
def test_raw_gemm_scores_match_torch_fp16(d):
    import torch
    nu, ni = 300, 1000
    m = _model('bpr', nu, ni, d)
    users = torch.arange(0, nu, dtype=torch.int32, device=m.device)
    _, _, dbg = m.engine.topk(users, 10, None, return_values=True, method='tensor', debug_scores=True)
    U = m.engine.U[:, :d].to(torch.float16).float()
    V = m.engine.V[:, :d].to(torch.float16).float()
    want = U @ V.T
    got = dbg[:, :ni]
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), float((got - want).abs().max())


@pytest.mark.parametrize('kind', ['bpr', 'gbpr', 'cml'])
@pytest.mark.parametrize('nu,ni,d,K,T', [(900, 5000, 128, 100, 700), (300, 70001, 100, 10, 300), (2100, 1300, 64, 50, 2100),
                                        (64, 20000, 20, 200, 40)])
def test_tensor_topk_is_bit_identical_to_exact(kind, nu, ni, d, K, T):
    import torch
    rng = np.random.default_rng(nu + ni)
    m = _model(kind, nu, ni, d, std=0.3 if kind != 'cml' else 0.1)
    # exact ties: duplicate item rows (+ bias)
    with torch.no_grad():
        m.engine.V[ni // 2:ni // 2 + 4] = m.engine.V[5]
        if kind == 'gbpr':
            m.engine.b[ni // 2:ni // 2 + 4] = m.engine.b[5]
    csr = _train_csr(rng, nu, ni, 60, m.device)
    users = torch.from_numpy(rng.permutation(nu)[:T].astype(np.int32)).to(m.device)
    ei, ev = m.engine.topk(users, K, csr, return_values=True, method='exact')
    ti, tv = m.engine.topk(users, K, csr, return_values=True, method='tensor')
    assert torch.equal(ei, ti), 'first mismatch at %s' % (torch.nonzero(ei != ti)[:3].tolist(),)
    assert torch.equal(ev, tv)
    st = m.engine.tc_stats.cpu().numpy()
    assert st[0] <= max(2, T // 50), 'the tensor path handed %d of %d rows to the exact fallback' % (st[0], T)


def test_degenerate_scores_fall_back_to_exact_rows():
    """All-equal scores: every item is within 2 eps of the K-th best -> candidate overflow -> exact kernel for those rows."""
    import torch
    nu, ni, d, K = 260, 4000, 64, 20
    m = _model('bpr', nu, ni, d)
    with torch.no_grad():
        m.engine.V[:, :d] = m.engine.V[0, :d]          # identical items: ties everywhere, lower id wins
    users = torch.arange(nu, dtype=torch.int32, device=m.device)
    ti = m.engine.topk(users, K, None, method='tensor')
    assert torch.equal(ti, torch.arange(K, dtype=torch.int32, device=m.device).expand(nu, K))


def test_values_outside_fp16_range_fall_back_to_exact():
    import torch
    nu, ni, d, K = 300, 3000, 64, 10
    m = _model('bpr', nu, ni, d)
    with torch.no_grad():
        m.engine.V[7, 3] = 1e6          # not representable in fp16
    users = torch.arange(nu, dtype=torch.int32, device=m.device)
    ei = m.engine.topk(users, K, None, method='exact')
    ti = m.engine.topk(users, K, None, method='tensor')
    assert torch.equal(ei, ti) and int(m.engine.tc_stats[0].item()) == nu


def test_wide_kernel_equals_the_exact_kernel(monkeypatch):
    """CF_TC_WIDE=1: the 256-item-tile kernel (one N = 256 tcgen05.mma per tile, one accumulator per M tile) returns the
    same lists and fp64 scores as the exact kernel (the default for every d is the 128-item-tile kernel)."""
    import numpy as np
    import torch
    from scipy.sparse import lil_matrix
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    monkeypatch.setenv('CF_TC_WIDE', '1')
    rng = np.random.default_rng(3)
    for cls, nu, ni, d, K in ((BPRMF, 300, 5000, 128, 100), (CML, 257, 3001, 64, 37), (BPRMF, 64, 777, 20, 200)):
        m = cls(nu, ni, n_factors=d, verbose=False, seed=2)
        tra = lil_matrix((nu, ni), dtype=np.float32)
        for u in range(nu):
            tra[u, rng.choice(ni, 25, replace=False)] = 1
        csr = DeviceCSR.from_scipy(tra, m.device)
        users = torch.arange(nu, dtype=torch.int32, device=m.device)
        ti, tv = m.engine.topk(users, K, csr, return_values=True, method='tensor')
        ei, ev = m.engine.topk(users, K, csr, return_values=True, method='exact')
        assert torch.equal(ti, ei) and torch.equal(tv, ev)


def test_paired_kernel_equals_the_exact_kernel(monkeypatch):
    """CF_TC_PAIR=1: the cta_group::2 kernel (clusters of two CTAs, one M = 256, N = 256 tcgen05.mma per k-step, two candidate
    buffers per row and split) returns the same lists and fp64 scores as the exact kernel."""
    import numpy as np
    import torch
    from scipy.sparse import lil_matrix
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    monkeypatch.setenv('CF_TC_PAIR', '1')
    rng = np.random.default_rng(5)
    for cls, nu, ni, d, K in ((BPRMF, 300, 5000, 128, 100), (CML, 257, 3001, 64, 37), (BPRMF, 64, 777, 20, 200),
                              (GBPRMF, 130, 40000, 200, 50), (CML, 1000, 20011, 128, 100)):
        m = cls(nu, ni, n_factors=d, verbose=False, seed=2)
        tra = lil_matrix((nu, ni), dtype=np.float32)
        for u in range(nu):
            tra[u, rng.choice(ni, 25, replace=False)] = 1
        csr = DeviceCSR.from_scipy(tra, m.device)
        users = torch.arange(nu, dtype=torch.int32, device=m.device)
        ti, tv = m.engine.topk(users, K, csr, return_values=True, method='tensor')
        ei, ev = m.engine.topk(users, K, csr, return_values=True, method='exact')
        assert torch.equal(ti, ei) and torch.equal(tv, ev), (cls.__name__, nu, ni, d, K)
        assert int(m.engine.tc_stats[0].item()) == 0       # no row fell back to the exact kernel


@pytest.mark.parametrize('kind,nu,ni,d,K,T,masked', [('cml', 700, 30011, 128, 1000, 600, True), ('bpr', 300, 9000, 64, 201, 300, True),
                                                    ('gbpr', 520, 12345, 100, 500, 520, True), ('bpr', 260, 5000, 128, 1000, 260, False),
                                                    ('cml', 300, 1500, 32, 1000, 300, True), ('bpr', 64, 700, 20, 1024, 64, True)])
def test_topk_above_200_runs_in_rounds_and_equals_exact(kind, nu, ni, d, K, T, masked):
    """K in (200, 1024] (the CML tail re-recommends at topN = 1000, cml.py:203-211): rounds of 200 over a mask that grows by
    the earlier rounds' results; lists and fp64 scores equal the exact kernel's, including exact ties, catalogues with
    fewer unmasked items than K (-1 padding) and a query list that repeats users."""
    import torch
    rng = np.random.default_rng(nu + ni + K)
    m = _model(kind, nu, ni, d, std=0.3 if kind != 'cml' else 0.1)
    with torch.no_grad():
        m.engine.V[ni // 2:ni // 2 + 4] = m.engine.V[5]
        if kind == 'gbpr':
            m.engine.b[ni // 2:ni // 2 + 4] = m.engine.b[5]
    csr = _train_csr(rng, nu, ni, 300 if ni < 2000 else 60, m.device) if masked else None
    u = rng.permutation(nu)[:T].astype(np.int32)
    u[-3:] = u[:3]                                   # repeated query users
    users = torch.from_numpy(u).to(m.device)
    ei, ev = m.engine.topk(users, min(K, ni), csr, return_values=True, method='exact')
    ti, tv = m.engine.topk(users, min(K, ni), csr, return_values=True, method='tensor')
    assert torch.equal(ei, ti), 'first mismatch at %s' % (torch.nonzero(ei != ti)[:3].tolist(),)
    assert torch.equal(ev, tv)
    st = m.engine.tc_stats.cpu().numpy()
    assert st[0] <= max(3, T // 50), 'the tensor path handed %d of %d rows to the exact fallback' % (st[0], T)


def test_rounds_hand_overflowed_rows_to_the_exact_kernel():
    """Identical items: every sweep overflows its candidate buffers -> the rows turn sticky and the exact kernel writes all K."""
    import torch
    nu, ni, d, K = 260, 4000, 64, 450
    m = _model('bpr', nu, ni, d)
    with torch.no_grad():
        m.engine.V[:, :d] = m.engine.V[0, :d]
    users = torch.arange(nu, dtype=torch.int32, device=m.device)
    ti = m.engine.topk(users, K, None, method='tensor')
    assert torch.equal(ti, torch.arange(K, dtype=torch.int32, device=m.device).expand(nu, K))
    assert int(m.engine.tc_stats[0].item()) == nu
