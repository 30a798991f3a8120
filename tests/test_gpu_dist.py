"""Single-GPU checks of the multi-GPU building blocks: the exchange-mode step + owner-side apply (world = 1 must equal the
plain fused step), and item-sharded (item % P) top-K + merge emulated on one device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _state(m):
    return {k: v.cpu().numpy() for k, v in m.state_dict().items()}


@pytest.mark.parametrize('transport', ['nccl', 'peer', 'peer-push', 'fetch', 'auto', 'replicate'])
@pytest.mark.parametrize('kind', ['bpr', 'cml'])
def test_exchange_mode_world1_equals_fused_step(kind, transport):
    import torch
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML
    from collaborativefilteringusingtensorflow_b200.dist import DistributedTrainer
    nu, ni, d, B, W = 400, 300, 128, 512, 3
    mk = (lambda: BPRMF(nu, ni, n_factors=d, reg=0.05, verbose=False, seed=4)) if kind == 'bpr' else \
         (lambda: CML(nu, ni, n_factors=d, reg_cov=1.0, margin=1.0, init_stddev=0.05, verbose=False, seed=4))   # row norms < clip at init
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    rng = np.random.default_rng(0)

    class NoSampler(object):
        batch_size = B
    tr = DistributedTrainer(b, NoSampler(), ni, 1, 0, item_transport=transport)
    for s in range(3):
        pairs = np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B)], 1).astype(np.int32)
        negs = rng.integers(0, ni, (B, W)).astype(np.int32)
        la = a.step(pairs, negs)
        lb = float(tr.step_chunk(torch.from_numpy(pairs).cuda(), torch.from_numpy(negs).cuda(), B)[0].item())
        b.engine.check_flags()
        assert abs(la - lb) < 1e-5 * abs(la)
        sa, sb = _state(a), _state(b)
        for k in sa:
            np.testing.assert_allclose(sb[k], sa[k], rtol=5e-5, atol=2e-6 if kind == 'bpr' else 1e-5,
                                       err_msg='%s step %d %s' % (kind, s, k))   # CML: rank-weighted coefficients ~10, fp32 sums regrouped


def test_item_mod_sharded_topk_merge_equals_single_shot():
    import torch
    from collaborativefilteringusingtensorflow_b200 import BPRMF, _lib
    from collaborativefilteringusingtensorflow_b200.engine import FactorEngine
    from collaborativefilteringusingtensorflow_b200.sparse import DeviceCSR
    from scipy.sparse import lil_matrix
    rng = np.random.default_rng(1)
    nu, ni, d, K, P = 20, 3001, 64, 50, 3
    m = BPRMF(nu, ni, n_factors=d, verbose=False, seed=5)
    tra = lil_matrix((nu, ni), dtype=np.float32)
    for u in range(nu):
        tra[u, rng.choice(ni, 40, replace=False)] = 1
    whole_i, whole_v = m.engine.topk(None, K, DeviceCSR.from_scipy(tra, m.device), return_values=True)
    idx = torch.empty(P, nu, K, dtype=torch.int32, device=m.device)
    val = torch.empty(P, nu, K, dtype=torch.float64, device=m.device)
    for p in range(P):
        shard = FactorEngine('bpr', nu, len(range(p, ni, P)), d, m.device, seed=0)
        shard.U.copy_(m.engine.U)
        shard.V.copy_(m.engine.V[p::P])
        local_mask = lil_matrix((nu, shard.n_items), dtype=np.float32)
        for u in range(nu):
            cols = [c // P for c in tra.rows[u] if c % P == p]
            if cols:
                local_mask[u, cols] = 1
        li, lv = shard.topk(None, K, DeviceCSR.from_scipy(local_mask, m.device), return_values=True)
        idx[p], val[p] = torch.where(li >= 0, li * P + p, li), lv
    out_i = torch.empty(nu, K, dtype=torch.int32, device=m.device)
    out_v = torch.empty(nu, K, dtype=torch.float64, device=m.device)
    _lib.check(_lib.lib().cf_topk_merge(idx.data_ptr(), val.data_ptr(), P, nu, K, out_i.data_ptr(), out_v.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), 'merge')
    assert torch.equal(out_i, whole_i) and torch.equal(out_v, whole_v)


@pytest.mark.parametrize('W', [4, 40])      # 40 negatives do not fit one tile of the d=128 kernel (30 entries): two tiles, CML two passes
@pytest.mark.parametrize('kind', ['bpr', 'cml'])
def test_peer_pull_from_three_shards_equals_fetched_rows(kind, W):
    """The peer-pull step (item rows read from their owners' shards by GLOBAL id: item i = row i // P of shard i % P)
    against the fetched-rows exchange step on the same minibatch: same user update, same gradient rows.  The three
    "peers" are three tensors of this process -- the kernel only sees pointers."""
    import torch
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, _lib
    from collaborativefilteringusingtensorflow_b200.dist import ItemExchange
    nu, ni, d, B, P = 300, 1001, 128, 2048 if W == 4 else 512, 3
    mk = (lambda: BPRMF(nu, ni, n_factors=d, reg=0.05, verbose=False, seed=9)) if kind == 'bpr' else \
         (lambda: CML(nu, ni, n_factors=d, reg_cov=1.0, margin=1.0, init_stddev=0.05, verbose=False, seed=9))
    rng = np.random.default_rng(3)
    pairs = torch.from_numpy(np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B)], 1).astype(np.int32)).cuda()
    negs = torch.from_numpy(rng.integers(0, ni, (B, W)).astype(np.int32)).cuda()
    items = torch.cat([pairs[:, 1:2], negs], 1).to(torch.int64)
    plan = ItemExchange(1, 0).plan_local(items, ni)                    # dedupe only: occ_local -> row of the compact buffers
    out = {}
    for mode in ('fetch', 'pull'):
        m = mk()
        eng = m.engine
        shards = [eng.V[p::P].contiguous() for p in range(P)]
        Gbuf = torch.zeros(plan.n_req, eng.ld, device=eng.device)
        a = _lib.StepArgs()
        if mode == 'pull':
            lp, ln = pairs.contiguous(), negs.contiguous()
            gp, gn = plan.occ_local[:, 0].contiguous(), plan.occ_local[:, 1:].contiguous()
            a.V, a.n_items, a.n_peers = _lib.ptr(eng.V), ni, P
            for p in range(P):
                a.peerV[p] = shards[p].data_ptr()
            a.gslot_pos, a.gslot_neg = _lib.ptr(gp), _lib.ptr(gn)
        else:
            Vbuf = eng.V.index_select(0, plan.req_global)
            lp = torch.stack([pairs[:, 0], plan.occ_local[:, 0]], 1).contiguous()
            ln = plan.occ_local[:, 1:].contiguous()
            a.V, a.n_items = _lib.ptr(Vbuf), plan.n_req
        a.U, a.accU, a.accV = _lib.ptr(eng.U), _lib.ptr(eng.accU), _lib.ptr(eng.accV)
        a.n_users, a.d, a.ld = eng.n_users, eng.d, eng.ld
        a.pairs, a.negs = _lib.ptr(lp), _lib.ptr(ln)
        a.B, a.W, a.G, a.n_batches = B, W, 0, 1
        a.model, a.optimizer, a.update = eng.model_id, 0, _lib.UPDATE_SYNC
        h = eng.hyper
        a.use_rank_weight = int(bool(h['use_rank_weight']))
        a.lr, a.reg, a.margin, a.clip_norm, a.rho, a.weight = h['lr'], h['reg'], h['margin'], h['clip_norm'], h['rho'], h['weight']
        ws = eng._workspace(B, W, 0)
        a.metaU, a.metaV = _lib.ptr(ws['metaU']), _lib.ptr(ws['metaV'])
        a.slotU, a.slotV, a.slot_row = _lib.ptr(ws['slotU']), _lib.ptr(ws['slotV']), _lib.ptr(ws['slot_row'])
        a.staging, a.staging_rows = _lib.ptr(ws['staging']), ws['staging'].shape[0]
        a.counters = _lib.ptr(eng.counters)
        loss = torch.zeros(1, dtype=torch.float64, device=eng.device)
        a.loss, a.gradV, a.rank_items = _lib.ptr(loss), _lib.ptr(Gbuf), ni
        _lib.check(_lib.lib().cf_train_steps(a, torch.cuda.current_stream().cuda_stream), 'cf_train_steps')
        eng.check_flags()
        out[mode] = (eng.U.cpu().numpy(), Gbuf.cpu().numpy(), float(loss.item()))
    assert np.abs(out['fetch'][1]).max() > 0
    np.testing.assert_allclose(out['pull'][0], out['fetch'][0], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(out['pull'][1], out['fetch'][1], rtol=2e-5, atol=5e-5 if kind == 'cml' else 4e-6)   # red.add order differs (run to run, too)
    assert abs(out['pull'][2] - out['fetch'][2]) <= 1e-6 * abs(out['fetch'][2])


def test_peer_pull_rejects_out_of_range_global_ids():
    import torch
    from collaborativefilteringusingtensorflow_b200 import BPRMF
    from collaborativefilteringusingtensorflow_b200.dist import DistributedTrainer
    m = BPRMF(50, 40, n_factors=32, verbose=False, seed=1)

    class NoSampler(object):
        batch_size = 8
    tr = DistributedTrainer(m, NoSampler(), 40, 1, 0, item_transport='peer')
    pairs = torch.tensor([[k, k] for k in range(8)], dtype=torch.int32).cuda()
    negs = torch.full((8, 2), 3, dtype=torch.int32).cuda()
    before = m.engine.U.clone()
    pairs[5, 0] = 50                                                  # user id past the table
    tr.step_chunk(pairs, negs, 8)
    with pytest.raises(RuntimeError, match='out of range'):
        m.engine.check_flags()
    assert torch.equal(before, m.engine.U)


def test_distributed_evaluate_world1_equals_evaluate():
    from collaborativefilteringusingtensorflow_b200.dist import distributed_evaluate
    from collaborativefilteringusingtensorflow_b200.metrics.ranking import evaluateCV, evaluateLOOV
    rng = np.random.default_rng(2)
    T, n, k = 200, 500, 10
    truth = [set(rng.choice(n, int(rng.integers(1, 12)), replace=False).tolist()) for _ in range(T)]
    pred = [rng.choice(n, 20, replace=False).tolist() for _ in range(T)]
    names = ['pre', 'recall', 'ndcg', 'map', 'mrr', 'nope']
    assert distributed_evaluate(truth, pred, names, k) == evaluateCV(truth, pred, names, k)
    one = [int(next(iter(t))) for t in truth]
    assert distributed_evaluate(one, pred, ['hr', 'arhr', 'pre'], k, 'loov') == evaluateLOOV(one, pred, ['hr', 'arhr', 'pre'], k)
    with pytest.raises(ValueError):
        distributed_evaluate([], [], names, k)
    with pytest.raises(ZeroDivisionError):
        distributed_evaluate(truth[:3] + [set()], pred[:4], ['map'], k)


def _mk4(kind, nu, ni, d):
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF, WRMF
    if kind == 'bpr':
        return BPRMF(nu, ni, n_factors=d, reg=0.05, verbose=False, seed=6)
    if kind == 'cml':
        return CML(nu, ni, n_factors=d, reg_cov=1.0, margin=1.0, init_stddev=0.05, verbose=False, seed=6)
    if kind == 'gbpr':
        return GBPRMF(nu, ni, rho=0.4, gsize=3, reg=0.01, n_factors=d, verbose=False, seed=6)
    return WRMF(nu, ni, weight=2.0, reg=0.1, n_factors=d, verbose=False, seed=6)


@pytest.mark.parametrize('halves', [1, 2])
@pytest.mark.parametrize('kind,d,W', [('bpr', 128, 4), ('cml', 64, 4), ('gbpr', 64, 4), ('gbpr', 20, 4), ('wrmf', 100, 4),
                                      ('gbpr', 64, 13), ('cml', 128, 33)])     # the last two: entries span two tiles
def test_replicated_mode_equals_fused_step(kind, d, W, halves):
    """Gradient-only step into dense tables + cf_apply_dense (what every rank of ReplicatedTrainer runs) against the
    plain fused step.  halves=2 accumulates two half-batches into the same dense tables before the apply -- the sum the
    all_reduce forms over two ranks -- and must equal ONE step on the whole batch."""
    import torch
    from collaborativefilteringusingtensorflow_b200.dist import ReplicatedTrainer
    nu, ni, B, G = 300, 200, 1024 if W == 4 else 256, 3
    a, b = _mk4(kind, nu, ni, d), _mk4(kind, nu, ni, d)
    b.load_state_dict(a.state_dict())
    rng = np.random.default_rng(8)

    class NoSampler(object):
        batch_size = B
    tr = ReplicatedTrainer(b, NoSampler())
    for s in range(3):
        pairs = np.stack([rng.integers(0, nu, B), rng.integers(0, ni, B)], 1).astype(np.int32)
        negs = rng.integers(0, ni, (B, W)).astype(np.int32)
        group = rng.integers(0, nu, (B, G)).astype(np.int32)
        ratings = (rng.random(B) < 0.5).astype(np.float32)
        if kind == 'wrmf':
            la = a.step(np.concatenate([pairs, ratings[:, None]], 1).astype(np.float64))
        elif kind == 'gbpr':
            la = a.step(pairs, negs, group)
        else:
            la = a.step(pairs, negs)
        lb = 0.0
        h = B // halves
        for k in range(halves):
            sl = slice(k * h, (k + 1) * h)
            kw = dict(group=group[sl]) if kind == 'gbpr' else (dict(ratings=ratings[sl]) if kind == 'wrmf' else {})
            last = k == halves - 1
            if last:
                lb += float(tr.step_chunk(torch.from_numpy(pairs[sl]).cuda(), None if kind == 'wrmf' else torch.from_numpy(negs[sl]).cuda(), **kw)[0].item())
            else:   # another rank's contribution: gradients only
                lb += float(b.engine.train_batches(pairs[sl], None if kind == 'wrmf' else negs[sl], batch_size=h,
                                                   grad_tables=(tr.gU, tr.gV, tr.gb), **kw)[0].item())
        b.engine.check_flags()
        assert abs(la - lb) < 2e-5 * abs(la), (la, lb)
        assert float(tr.flat.abs().max().item()) == 0.0           # the applied rows were re-zeroed
        sa, sb = _state(a), _state(b)
        for k in sa:
            scale = float(np.abs(sa[k]).max())
            np.testing.assert_allclose(sb[k], sa[k], rtol=5e-5, atol=1e-4 * max(scale, 1.0) if kind == 'cml' else 4e-6 * max(scale, 1.0),   # CML: ~25 rank-weighted (x10) gradients per item row, fp32 sums regrouped
                                       err_msg='%s step %d %s' % (kind, s, k))


def test_owner_apply_reads_gradient_rows_in_place_from_segments():
    """cf_apply_rows with the gradient rows left in three "requesters'" buffers (owner-pull segments) against the same
    rows concatenated into one receive buffer: identical tables (rows received from several requesters are summed)."""
    import torch
    from collaborativefilteringusingtensorflow_b200 import CML, _lib
    n_rows, d, counts = 500, 128, [300, 0, 450]
    rng = np.random.default_rng(4)
    out = []
    ids = [np.sort(rng.choice(n_rows, c, replace=False)).astype(np.int32) for c in counts]
    grads = [(0.1 * rng.standard_normal((c + 7, d))).astype(np.float32) for c in counts]      # 7 rows of other owners first
    for mode in ('concat', 'segments'):
        m = CML(10, n_rows, n_factors=d, init_stddev=0.05, verbose=False, seed=3)
        eng = m.engine
        rows = torch.from_numpy(np.concatenate(ids)).cuda()
        n = int(rows.numel())
        bufs = [torch.from_numpy(g).cuda() for g in grads]
        cat = torch.cat([b[7:] for b in bufs]).contiguous()
        ap = _lib.ApplyArgs()
        ap.table, ap.acc, ap.n_rows, ap.d, ap.ld = _lib.ptr(eng.V), _lib.ptr(eng.accV), n_rows, d, eng.ld
        ap.rows, ap.n, ap.ldg = _lib.ptr(rows), n, eng.ld
        ap.model, ap.optimizer, ap.lr, ap.clip_norm = eng.model_id, 0, 0.1, 1.0
        meta = torch.zeros(n_rows, dtype=torch.int32, device='cuda')
        slot = torch.zeros(n_rows, dtype=torch.int32, device='cuda')
        slot_row = torch.full((n,), -1, dtype=torch.int32, device='cuda')
        staging = torch.zeros(n, eng.ld + 4, device='cuda')
        ap.meta, ap.slot, ap.slot_row, ap.staging, ap.staging_rows = _lib.ptr(meta), _lib.ptr(slot), _lib.ptr(slot_row), _lib.ptr(staging), n
        ap.counters = _lib.ptr(eng.counters)
        if mode == 'concat':
            ap.grads = _lib.ptr(cat)
        else:
            ap.n_segs, start = 3, 0
            for q in range(3):
                ap.seg_start[q] = start
                ap.seg_grads[q] = bufs[q].data_ptr() + 7 * eng.ld * 4
                start += counts[q]
            ap.seg_start[3] = start
            ap.first_seg = 2
        _lib.check(_lib.lib().cf_apply_rows(ap, torch.cuda.current_stream().cuda_stream), 'cf_apply_rows')
        eng.check_flags()
        assert int(meta.abs().sum().item()) == 0 and float(staging.abs().max().item()) == 0.0
        out.append((eng.V.cpu().numpy(), eng.accV.cpu().numpy()))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
