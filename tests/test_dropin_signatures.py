"""The drop-in boundary, checked against the reference itself: tests/golden/signatures_golden.json holds inspect.signature of
every reference model class (constructor, train, close), sampler (constructor, next_batch), metric entry point and loader
function, read off the reference's own modules (oracle/gen_refgraph_golden.py signatures; the TensorFlow-importing ones through
the TF-1.x stand-in).  The product's classes must take the SAME parameters in the SAME order with the SAME defaults -- the
reference drivers pass them positionally (testbprmf.py:44, testcml.py:49, testgbprmf.py:48, testwrmf.py:43) -- and may add
only keyword-only or trailing defaulted parameters (seed, verbose, optimizer, ...).  CPU only: nothing is constructed."""
import importlib
import inspect
import json
import os

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PKG = 'collaborativefilteringusingtensorflow_b200'
SIG = json.load(open(os.path.join(GOLDEN, 'signatures_golden.json')))
WHERE = {'models/bprmf.BPRMF': 'models.pl.models.bprmf', 'models/cml.CML': 'models.pl.models.cml',
         'models/gbprmf.GBPRMF': 'models.pl.models.gbprmf', 'models/prigp.PRIGP': 'models.pl.models.prigp',
         'models/cplr_u.CPLR': 'models.pl.models.cplr_u', 'models/wrmf.WRMF': 'models.basic.models.wrmf',
         'models/mf.MF': 'models.basic.models.mf', 'models/svd.SVD': 'models.basic.models.svd',
         'models/pop.PopRank': 'models.basic.models.pop', 'models/itemcf.ItemCF': 'models.basic.models.itemcf',
         'models/usercf.UserCF': 'models.basic.models.usercf',
         'ranking.evaluateCV': 'metrics.ranking', 'ranking.evaluateLOOV': 'metrics.ranking', 'rating.evaluate': 'metrics.rating',
         'IOUtil.loadSparseR': 'utils.IOUtil', 'Util.matBinarize': 'utils.Util'}


def _leading(fn, want, what):
    """``fn`` must start with the reference's parameters (names, order, defaults); whatever follows must be optional."""
    ps = [p for n, p in inspect.signature(fn).parameters.items() if n != 'self']
    assert len(ps) >= len(want), '%s takes fewer parameters than the reference' % what
    for p, (name, has_default, default) in zip(ps, want):
        assert p.name == name, '%s: parameter %r where the reference has %r' % (what, p.name, name)
        assert p.kind in (p.POSITIONAL_OR_KEYWORD, p.POSITIONAL_ONLY), '%s: %s must be positional' % (what, name)
        assert (p.default is not inspect.Parameter.empty) == has_default, '%s: %s default presence' % (what, name)
        if has_default:
            got = list(p.default) if isinstance(p.default, tuple) else p.default
            assert got == default, '%s: %s defaults to %r, the reference to %r' % (what, name, got, default)
    for p in ps[len(want):]:
        assert p.default is not inspect.Parameter.empty or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL), \
            '%s: extra parameter %s has no default' % (what, p.name)


@pytest.mark.parametrize('key', sorted(k for k in SIG if k.startswith('models/')))
def test_model_classes_take_the_reference_arguments(key):
    cls = getattr(importlib.import_module('%s.%s' % (PKG, WHERE[key])), key.split('.')[-1])
    _leading(cls.__init__, SIG[key]['init'], key + '.__init__')
    _leading(cls.train, SIG[key]['train'], key + '.train')
    if SIG[key]['close']:
        assert callable(getattr(cls, 'close', None)), key + ' has no close()'


@pytest.mark.parametrize('key', sorted(k for k in SIG if k.startswith('samplers/')))
def test_samplers_take_the_reference_arguments(key):
    module = key.split('/')[1].split('.')[0]
    cls = importlib.import_module('%s.samplers.%s' % (PKG, module)).Sampler
    _leading(cls.__init__, SIG[key]['init'], key + '.__init__')
    _leading(cls.next_batch, SIG[key]['next_batch'], key + '.next_batch')


@pytest.mark.parametrize('key', sorted(k for k in SIG if '/' not in k))
def test_functions_take_the_reference_arguments(key):
    fn = getattr(importlib.import_module('%s.%s' % (PKG, WHERE[key])), key.split('.')[-1])
    _leading(fn, SIG[key]['call'], key)
