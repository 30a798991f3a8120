"""Drop-in check against the reference's OWN train() loop: oracle/gen_refgraph_golden.py imported the reference's model
files unmodified (bprmf.py, cml.py, gbprmf.py, basic/models/wrmf.py), ran ``Model(...).train(fold, trasR, tstsR, sampler)``
on the TF1 stand-in of oracle/tf1_shim with a sampler that hands out recorded minibatches, and stored the tables, the
Adagrad accumulators, the per-step losses and the metric values train() RETURNED.  Here the product's classes get the
same constructor arguments, the same initial tables, the same sampler object and the same train() call -- through the C ABI
on the GPU -- and must return the same metric values and end with the same state.
Tolerance: 1e-5 relative + 1e-6 (north_star, fp32) on tables / accumulators / losses; metric values to 1e-9 (they are
means of small rationals: any difference would be a different recommended list)."""
import numpy as np
import pytest

import refgraph_cases as R

pytestmark = pytest.mark.gpu


def _model(name, c):
    from collaborativefilteringusingtensorflow_b200 import BPRMF, CML, GBPRMF, WRMF
    h, ev = c['hyper'], c['ev']
    nu, ni, d = c['init']['U'].shape[0], c['init']['V'].shape[0], c['init']['U'].shape[1]
    common = dict(topN=ev['topN'], split_method='cv', eval_metrics=list(ev['metrics']), n_factors=d,
                  batch_size=len(c['batches'][0][0]), max_iter=1, lr=h['lr'], verbose=False, seed=3)
    kind = R.KIND[name]
    if kind == 'bpr':
        return BPRMF(nu, ni, reg=h['reg'], **common)
    if kind == 'cml':
        return CML(nu, ni, reg_cov=h['reg_cov'], margin=h['margin'], use_rank_weight=h['use_rank_weight'],
                   clip_norm=h['clip_norm'], **common)
    if kind == 'gbpr':
        return GBPRMF(nu, ni, rho=h['rho'], gsize=c['batches'][0][2].shape[1], reg=h['reg'], **common)
    return WRMF(nu, ni, weight=h['weight'], reg=h['reg'], **common)


@pytest.mark.parametrize('name', R.CASES)
def test_train_returns_what_the_reference_train_returned(name):
    c = R.case(*R.load(), name)
    m = _model(name, c)
    m.load_state_dict(c['init'])
    scores = m.train(1, c['tra'].tocsr(), c['tst'].tocsr(), R.RecordedSampler(R.KIND[name], c['batches']))
    np.testing.assert_allclose(scores, c['ev']['scores'], rtol=0, atol=1e-9)
    st = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    for k, want in c['final'].items():
        rtol = 5e-5 if k.startswith('acc') else 1e-5          # accumulators hold g^2: twice the relative error
        np.testing.assert_allclose(st[k], want, rtol=rtol, atol=1e-6, err_msg='%s %s' % (name, k))


@pytest.mark.parametrize('name', ['svd', 'svd_d7'])
def test_svd_train_returns_what_the_reference_train_returned(name):
    """basic/models/svd.py through its own train() (recorded [B, 3] minibatches, its clipped-prediction evaluation with the
    reference's metrics/rating.py) against the product's SVD.train() with the same arguments."""
    import json
    from collaborativefilteringusingtensorflow_b200 import SVD
    z = R.golden('svd', 'refgraph')
    ev = json.loads(str(z[name + '/eval']))
    U0, V0, K0 = (z['%s/init/%s' % (name, k)] for k in 'UVK')
    batches = [[z['%s/batch%d' % (name, s)]] for s in range(2)]
    B = len(batches[0][0])
    m = SVD(U0.shape[0], V0.shape[0], eval_metrics=list(ev['metrics']), range_of_ratings=tuple(ev['range_of_ratings']), reg=0.05,
            n_factors=U0.shape[1], batch_size=B, max_iter=1, lr=0.1, verbose=False, seed=3)
    m.load_state_dict(dict(U=U0, V=V0, K=K0))

    class Recorded(object):
        k = 0

        def next_batch(self):
            self.k += 1
            return batches[self.k - 1][0]
    scores = m.train(1, np.zeros((2 * B, 3)), np.asarray(ev['tst']), Recorded())
    np.testing.assert_allclose(scores, ev['scores'], rtol=1e-5)
    st = {k: v.cpu().numpy() for k, v in m.state_dict().items()}
    for k in ('U', 'V', 'K', 'accU', 'accV', 'accK'):
        np.testing.assert_allclose(st[k], z['%s/step1/%s' % (name, k)], rtol=5e-5 if k.startswith('acc') else 1e-5, atol=1e-6,
                                   err_msg='%s %s' % (name, k))
