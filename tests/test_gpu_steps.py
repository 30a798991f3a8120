"""Parity of the fused CUDA step (through the model classes / C ABI) with
  (a) tests/golden/step_golden.npz -- torch-autograd restatement of the TF graphs + TF1-semantics Adagrad, and
  (b) the numpy oracle (oracle/steps.py) on seeded random batches at more shapes.
Tolerance: north_star's "within 1e-5 relative (fp32)", plus atol 1e-6 for near-zero elements."""
import json

import numpy as np
import pytest

from oracle import steps

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def _close(a, b, what='', stress=False):
    """Golden-vector tests: rtol 1e-5 + atol 1e-6 (north_star: 'within 1e-5 relative (fp32)').
    Stress shapes (hundreds of duplicate gradients per row, rank-weighted CML coefficients ~10): the fp32 sum of the
    duplicates is order-dependent at ~eps * sum|g_i| in ANY implementation (TF's segment-sum included), so elements
    near zero are compared at 1e-5 of THEIR OWN ROW's norm (1e-5 relative per row vector); Adagrad accumulators hold g^2
    (twice the relative error)."""
    a, b = np.asarray(a), np.asarray(b)
    rtol = 5e-5 if 'acc' in what else RTOL
    if not stress:
        np.testing.assert_allclose(a, b, rtol=rtol, atol=ATOL, err_msg=what)
        return
    b2 = b.reshape(b.shape[0], -1).astype(np.float64)
    a2 = a.reshape(b2.shape).astype(np.float64)
    atol = np.maximum(ATOL, 1e-5 * np.sqrt((b2 ** 2).sum(1)))[:, None]
    bad = np.abs(a2 - b2) > rtol * np.abs(b2) + atol
    assert not bad.any(), '%s: %d elements off, worst |diff| %.3g at row %d (row norm %.3g)' % (
        what, int(bad.sum()), float(np.abs(a2 - b2).max()), int(np.abs(a2 - b2).max(1).argmax()),
        float(np.sqrt((b2 ** 2).sum(1))[np.abs(a2 - b2).max(1).argmax()]))


def _mk(kind, nu, ni, d, **kw):
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF, WRMF
    cls = dict(bpr=BPRMF, cml=CML, gbpr=GBPRMF, wrmf=WRMF)[kind]
    return cls(nu, ni, n_factors=d, verbose=False, seed=1, **kw)


def _state(m):
    return {k: v.cpu().numpy() for k, v in m.state_dict().items()}


@pytest.mark.parametrize('name', ['bpr', 'bpr_w3'])
def test_bpr_golden(step_golden, name):
    g = step_golden
    h = json.loads(str(g[name + '/hyper']))
    U0, V0 = g[name + '/init/U'], g[name + '/init/V']
    m = _mk('bpr', U0.shape[0], V0.shape[0], U0.shape[1], reg=h['reg'], lr=h['lr'])
    m.load_state_dict(dict(U=U0, V=V0))
    for s in range(2):
        loss = m.step(g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)])
        st = _state(m)
        for k in ('U', 'V', 'accU', 'accV'):
            _close(st[k], g['%s/step%d/%s' % (name, s, k)], '%s step %d %s' % (name, s, k))
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * abs(loss)


@pytest.mark.parametrize('name', ['cml', 'cml_norank_noreg'])
def test_cml_golden(step_golden, name):
    g = step_golden
    h = json.loads(str(g[name + '/hyper']))
    U0, V0 = g[name + '/init/U'], g[name + '/init/V']
    m = _mk('cml', U0.shape[0], V0.shape[0], U0.shape[1], reg_cov=h['reg_cov'], margin=h['margin'],
            use_rank_weight=h['use_rank_weight'], clip_norm=h['clip_norm'], lr=h['lr'])
    m.load_state_dict(dict(U=U0, V=V0))
    for s in range(2):
        loss = m.step(g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)])
        st = _state(m)
        for k in ('U', 'V', 'accU', 'accV'):
            _close(st[k], g['%s/step%d/%s' % (name, s, k)], '%s step %d %s' % (name, s, k))
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * max(1.0, abs(loss))
        assert np.linalg.norm(st['U'], axis=1).max() <= h['clip_norm'] * (1 + 1e-6)


@pytest.mark.parametrize('name', ['gbpr', 'gbpr_g1'])
def test_gbpr_golden(step_golden, name):
    g = step_golden
    h = json.loads(str(g[name + '/hyper']))
    U0, V0, b0 = g[name + '/init/U'], g[name + '/init/V'], g[name + '/init/b']
    G = g['%s/batch0/2' % name].shape[1]
    m = _mk('gbpr', U0.shape[0], V0.shape[0], U0.shape[1], rho=h['rho'], gsize=G, reg=h['reg'], lr=h['lr'])
    m.load_state_dict(dict(U=U0, V=V0, b=b0))
    for s in range(2):
        loss = m.step(*[g['%s/batch%d/%d' % (name, s, k)] for k in range(3)])
        st = _state(m)
        for k in ('U', 'V', 'b', 'accU', 'accV', 'accb'):
            _close(st[k], g['%s/step%d/%s' % (name, s, k)], '%s step %d %s' % (name, s, k))
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * abs(loss)


def test_wrmf_golden(step_golden):
    g, name = step_golden, 'wrmf'
    h = json.loads(str(g[name + '/hyper']))
    U0, V0 = g[name + '/init/U'], g[name + '/init/V']
    m = _mk('wrmf', U0.shape[0], V0.shape[0], U0.shape[1], weight=h['weight'], reg=h['reg'], lr=h['lr'])
    m.load_state_dict(dict(U=U0, V=V0))
    for s in range(2):
        ui, r = g['%s/batch%d/0' % (name, s)], g['%s/batch%d/1' % (name, s)]
        uir = np.concatenate([ui.astype(np.float64), r[:, None].astype(np.float64)], axis=1)   # the sampler_rating layout
        loss = m.step(uir)
        st = _state(m)
        for k in ('U', 'V', 'accU', 'accV'):
            _close(st[k], g['%s/step%d/%s' % (name, s, k)], '%s step %d %s' % (name, s, k))
        assert abs(loss - float(g['%s/loss%d' % (name, s)])) < 1e-4 * abs(loss)


# ---------------------------------------------------------------- oracle parity at more shapes / sizes
SHAPES = [  # (d, B, W, nu, ni): exercises 8/16/32-lane groups, 2 and 4 vectors per lane, W tiles > 8, ragged last warp
    (20, 100, 1, 943, 1682), (50, 50, 5, 300, 500), (64, 257, 2, 2000, 3000), (100, 100, 1, 943, 1682),
    (128, 4096, 5, 20000, 10000), (128, 1000, 11, 500, 400), (200, 333, 3, 1000, 1000), (300, 64, 9, 200, 300),
    (7, 33, 4, 50, 60), (512, 40, 2, 100, 100),
    # more entries than one shared-memory tile holds: (d<=32: 6 per tile) (d=512: 5 per tile) (33 negatives > 30 lanes)
    (20, 64, 8, 100, 120), (512, 32, 9, 100, 100), (16, 50, 20, 80, 90), (128, 64, 33, 300, 200),
]


def _rand_batch(rng, nu, ni, B, W):
    return (np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B)], 1).astype(np.int32),
            rng.integers(0, ni, (B, W)).astype(np.int64))


@pytest.mark.parametrize('d,B,W,nu,ni', SHAPES)
@pytest.mark.parametrize('optimizer', ['adagrad', 'sgd'])
def test_bpr_vs_oracle(d, B, W, nu, ni, optimizer):
    rng = np.random.default_rng(d * 1000 + B)
    m = _mk('bpr', nu, ni, d, reg=0.05, lr=0.1, optimizer=optimizer)
    P = _state(m)
    opt = steps.ADAGRAD if optimizer == 'adagrad' else steps.SGD
    for s in range(2):
        pairs, negs = _rand_batch(rng, nu, ni, B, W)
        loss = m.step(pairs, negs)
        ol = steps.bpr_step(P['U'], P['V'], P['accU'], P['accV'], pairs, negs, 0.1, 0.05, opt)
        st = _state(m)
        for k in ('U', 'V') + (('accU', 'accV') if optimizer == 'adagrad' else ()):
            _close(st[k], P[k], 'bpr d=%d step %d %s' % (d, s, k), stress=True)
        assert abs(loss - ol) < 1e-4 * abs(ol)


@pytest.mark.parametrize('d,B,W,nu,ni', SHAPES)
def test_cml_vs_oracle(d, B, W, nu, ni):
    rng = np.random.default_rng(d * 1000 + B + 1)
    m = _mk('cml', nu, ni, d, reg_cov=1.0, margin=1.0, use_rank_weight=True, clip_norm=1.0, lr=0.1,
            init_stddev=0.3 / np.sqrt(d) * 3)
    P = _state(m)
    for s in range(2):
        pairs, negs = _rand_batch(rng, nu, ni, B, W)
        # pairs whose hinge / impostor argument sits within rounding distance of the relu / indicator kink are dropped
        # (fp32 summation order decides their side in any implementation); the rest of the minibatch is compared
        keep = steps.cml_forward(P['U'], P['V'], pairs, negs, 1.0, True, ni)['kink_per_pair'] >= 1e-5
        assert keep.mean() > 0.9
        pairs, negs = np.ascontiguousarray(pairs[keep]), np.ascontiguousarray(negs[keep])
        loss = m.step(pairs, negs)
        ol = steps.cml_step(P['U'], P['V'], P['accU'], P['accV'], pairs, negs, 0.1, 1.0, 1.0, True, 1.0)
        st = _state(m)
        for k in ('U', 'V', 'accU', 'accV'):
            _close(st[k], P[k], 'cml d=%d step %d %s' % (d, s, k), stress=True)
        assert abs(loss - ol) < 1e-4 * max(1.0, abs(ol))


@pytest.mark.parametrize('d,B,W,nu,ni', SHAPES[:8] + SHAPES[10:])
@pytest.mark.parametrize('G', [1, 3, 6])
def test_gbpr_vs_oracle(d, B, W, nu, ni, G):
    rng = np.random.default_rng(d * 1000 + B + G)
    m = _mk('gbpr', nu, ni, d, rho=0.4, gsize=G, reg=0.01, lr=0.1)
    P = _state(m)
    for s in range(2):
        pairs, negs = _rand_batch(rng, nu, ni, B, W)
        group = rng.integers(0, nu, (B, G)).astype(np.int64)
        group[::7, 0] = pairs[::7, 0]                 # the user inside its own group (sampler_gbpr.py:41 allows it)
        loss = m.step(pairs, negs, group)
        ol = steps.gbpr_step(P['U'], P['V'], P['b'], P['accU'], P['accV'], P['accb'], pairs, negs, group, 0.1, 0.01, 0.4)
        st = _state(m)
        for k in ('U', 'V', 'b', 'accU', 'accV', 'accb'):
            _close(st[k], P[k], 'gbpr d=%d G=%d step %d %s' % (d, G, s, k), stress=True)
        assert abs(loss - ol) < 1e-4 * abs(ol)


@pytest.mark.parametrize('d,B,nu,ni', [(10, 200, 943, 1682), (100, 200, 943, 1682), (128, 5000, 3000, 2000)])
def test_wrmf_vs_oracle(d, B, nu, ni):
    rng = np.random.default_rng(d + B)
    m = _mk('wrmf', nu, ni, d, weight=2.0, reg=0.1, lr=0.1)
    P = _state(m)
    for s in range(2):
        uir = np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B), (rng.random(B) < 0.5)], 1).astype(np.float64)
        loss = m.step(uir)
        ol = steps.wrmf_step(P['U'], P['V'], P['accU'], P['accV'], uir, 0.1, 0.1, 2.0)
        st = _state(m)
        for k in ('U', 'V', 'accU', 'accV'):
            _close(st[k], P[k], 'wrmf step %d %s' % (s, k), stress=True)
        assert abs(loss - ol) < 1e-4 * abs(ol)


def test_multi_minibatch_call_equals_step_by_step():
    """cf_train_steps with n_batches > 1 == the same minibatches one call at a time (bprmf.py:143-148)."""
    rng = np.random.default_rng(9)
    nu, ni, d, B, W, nb = 500, 700, 64, 128, 3, 6
    a, b = _mk('bpr', nu, ni, d, reg=0.02), _mk('bpr', nu, ni, d, reg=0.02)
    b.load_state_dict(a.state_dict())
    pairs, negs = _rand_batch(rng, nu, ni, B * nb, W)
    la = a.engine.train_batches(pairs, negs, batch_size=B).cpu().numpy()
    lb = [b.step(pairs[k * B:(k + 1) * B], negs[k * B:(k + 1) * B]) for k in range(nb)]
    np.testing.assert_allclose(la, lb, rtol=1e-6)
    sa, sb = _state(a), _state(b)
    for k in sa:
        _close(sa[k], sb[k], k)


def test_hogwild_without_duplicates_equals_sync():
    """With no repeated row in the minibatch the racy mode has nothing to race on."""
    rng = np.random.default_rng(11)
    nu, ni, d, B = 4000, 9000, 128, 1000
    a, b = _mk('bpr', nu, ni, d, reg=0.02), _mk('bpr', nu, ni, d, reg=0.02, update='hogwild')
    b.load_state_dict(a.state_dict())
    items = rng.permutation(ni)[:3 * B].reshape(B, 3)
    pairs = np.stack([rng.permutation(nu)[:B], items[:, 0]], 1)
    negs = items[:, 1:]
    a.step(pairs, negs)
    b.step(pairs, negs)
    sa, sb = _state(a), _state(b)
    for k in sa:
        _close(sa[k], sb[k], k)


def test_out_of_range_index_is_reported():
    m = _mk('bpr', 50, 60, 16)
    before = _state(m)
    with pytest.raises(RuntimeError, match='out of range'):
        m.step(np.array([[1, 2], [50, 3]]), np.array([[4], [5]]))
    after = _state(m)
    np.testing.assert_array_equal(before['U'], after['U'])
    m.step(np.array([[1, 2], [49, 3]]), np.array([[4], [5]]))      # the workspace is usable again


def test_empty_and_ragged_inputs_raise():
    m = _mk('bpr', 50, 60, 16)
    with pytest.raises(ValueError):
        m.step(np.zeros((0, 2), dtype=np.int32), np.zeros((0, 1), dtype=np.int64))
    with pytest.raises(ValueError):
        m.step(np.zeros((4, 2), dtype=np.int32), np.zeros((3, 1), dtype=np.int64))


@pytest.mark.parametrize('kind,W,G', [('bpr', 1, 0), ('cml', 1, 0), ('gbpr', 5, 3), ('gbpr', 5, 1)])
@pytest.mark.parametrize('d', [128, 100, 68, 64, 40])
@pytest.mark.parametrize('optimizer', ['adagrad', 'sgd'])
def test_specialised_step_kernel_equals_the_generic_one(monkeypatch, kind, W, G, d, optimizer):
    """cf_step_fast.cu (unrolled + software-pipelined; taken for one negative per pair and for GBPR with 5 negatives and a group
    of 3 or 1, at 32 < ld <= 128) runs the generic kernel's arithmetic operation for operation: on a minibatch without
    repeated rows the tables and accumulators are bit-identical; with repeats only the order of the staged red.adds differs
    (as it does from run to run)."""
    nu, ni, B = 5000 * (2 + G), 5000 * (1 + W), 5000
    kw = dict(reg=0.05) if kind == 'bpr' else dict(reg=0.01, rho=0.4, gsize=G) if kind == 'gbpr' else \
        dict(reg_cov=1.0, margin=1.0, use_rank_weight=True, clip_norm=1.0, init_stddev=0.9 / np.sqrt(d))
    rng = np.random.default_rng(d + W)
    uperm = rng.permutation(nu)
    uniq = [np.stack([uperm[:B], rng.permutation(5000)], 1).astype(np.int32),
            (5000 + rng.permutation(5000 * W)).reshape(B, W).astype(np.int64)]                 # no row occurs twice
    dup = [np.stack([rng.integers(0, 300, B), rng.integers(0, 200, B)], 1).astype(np.int32),
           rng.integers(0, 200, (B, W)).astype(np.int64)]                                       # almost every row repeats
    if G:
        uniq.append(uperm[B:B + B * G].reshape(B, G).astype(np.int64))
        dup.append(rng.integers(0, 300, (B, G)).astype(np.int64))
        dup[2][::7, 0] = dup[0][::7, 0]                                                         # the user inside its own group
    states, losses = [], []
    for generic in (True, False):
        if generic:
            monkeypatch.setenv('CF_STEP_GENERIC', '1')
        else:
            monkeypatch.delenv('CF_STEP_GENERIC')
        m = _mk(kind, nu, ni, d, lr=0.1, optimizer=optimizer, **kw)
        l1 = m.step(*uniq)
        s1 = _state(m)
        l2 = m.step(*dup)
        l3 = m.step(*[x[:777] for x in uniq])          # ragged tail: fewer pairs than groups in the grid
        states.append((s1, _state(m)))
        losses.append((l1, l2, l3))
    (g1, g2), (f1, f2) = states
    for k in g1:
        assert np.array_equal(g1[k], f1[k]), '%s differs on a minibatch without repeated rows' % k
        _close(f2[k], g2[k], '%s W=%d d=%d %s' % (kind, W, d, k), stress=True)
    for a, b in zip(*losses):
        assert abs(a - b) <= 1e-6 * abs(a)
