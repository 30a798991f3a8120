"""The numpy restatement of the PRIGP / CPLR minibatch steps (oracle/steps.py: hand-derived gradients + TF1 Adagrad on the
summed row gradients) against an independent torch-autograd restatement of the same TF graphs (tests/golden/
tuple_golden.npz, oracle/gen_golden.py tuples).  Also against the reference's own prigp.py / cplr_u.py graphs run on the TF-1.x stand-in (tuple_refgraph_golden.npz).  TensorFlow itself is not installable."""
import json
import os

import numpy as np
import pytest

from oracle import steps

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module', params=['autograd', 'refgraph'])
def tg(request):
    """'autograd': the torch-autograd restatement; 'refgraph': the reference's own prigp.py / cplr_u.py graphs run on the TF1
    stand-in (oracle/gen_refgraph_golden.py)."""
    import refgraph_cases
    return refgraph_cases.golden('tuple', request.param)


@pytest.mark.parametrize('name', ['prigp', 'prigp_d20', 'cplr', 'cplr_d20'])
def test_hand_derived_steps_equal_autograd(tg, name):
    h = json.loads(str(tg[name + '/hyper']))
    P = {k: tg['%s/init/%s' % (name, k)].copy() for k in ('U', 'V', 'b')}
    acc = {k: np.full_like(P[k], 0.1) for k in P}
    for s in range(2):
        t, c = tg['%s/batch%d/tuples' % (name, s)], tg['%s/batch%d/coefs' % (name, s)]
        if name.startswith('prigp'):
            loss = steps.prigp_step(P['U'], P['V'], P['b'], acc['U'], acc['V'], t, h['lr'], h['reg'], h['alpha'])
        else:
            loss = steps.cplr_step(P['U'], P['V'], P['b'], acc['U'], acc['V'], acc['b'], t, c, h['lr'], h['reg'], h['alpha'], h['beta'], h['gamma'])
        assert abs(loss - float(tg['%s/loss%d' % (name, s)])) <= 2e-5 * abs(loss)
        for k in P:
            np.testing.assert_allclose(P[k], tg['%s/step%d/%s' % (name, s, k)], rtol=1e-5, atol=1e-6, err_msg='%s step %d %s' % (name, s, k))
            key = '%s/step%d/acc%s' % (name, s, k)
            if key in tg.files:
                np.testing.assert_allclose(acc[k], tg[key], rtol=2e-5, atol=1e-6, err_msg=key)
    if name.startswith('prigp'):
        assert np.array_equal(P['b'], tg[name + '/init/b'])          # prigp.py:134: the bias is not in the optimizer's var_list
