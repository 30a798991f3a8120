"""Single-GPU checks of the multi-GPU building blocks: the exchange-mode step + owner-side apply (world = 1 must equal the
plain fused step), and item-sharded (item % P) top-K + merge emulated on one device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _state(m):
    return {k: v.cpu().numpy() for k, v in m.state_dict().items()}


@pytest.mark.parametrize('kind', ['bpr', 'cml'])
def test_exchange_mode_world1_equals_fused_step(kind):
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
    tr = DistributedTrainer(b, NoSampler(), ni, 1, 0)
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
