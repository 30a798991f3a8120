import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip = pytest.mark.skip(reason='no CUDA device')
        for it in items:
            if 'gpu' in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope='session', params=['autograd', 'refgraph'])
def step_golden(request):
    """Two sources of expected values for the same inputs (tests/golden/README.md): 'autograd' = the torch-autograd
    restatement of the TF1 graphs (oracle/gen_golden.py); 'refgraph' = the reference's OWN model files, imported
    unmodified and run through their train() on the TF1 stand-in of oracle/tf1_shim (oracle/gen_refgraph_golden.py)."""
    import refgraph_cases
    return refgraph_cases.golden('step', request.param)


@pytest.fixture(scope='session')
def ranking_golden():
    return json.load(open(os.path.join(GOLDEN, 'ranking_golden.json')))


@pytest.fixture(scope='session')
def ml100k():
    """ml-100k fold 1 after the drivers' ``rating > 3`` binarisation (testbprmf.py:21,34), as scipy lil matrices."""
    from scipy.sparse import coo_matrix
    z = np.load(os.path.join(GOLDEN, 'ml100k_fold1.npz'))
    st = json.load(open(os.path.join(GOLDEN, 'ml100k_fold1_stats.json')))
    out = {}
    for part in ('tra', 'tst'):
        u, i, r = z[part + '_u'].astype(np.int64), z[part + '_i'].astype(np.int64), z[part + '_r']
        keep = r > 3
        m = coo_matrix((np.ones(int(keep.sum()), dtype=np.float32), (u[keep], i[keep])),
                       shape=(st['n_users'], st['n_items'])).tolil()
        out[part] = m
        out[part + '_raw'] = (u, i, r)
    out['stats'] = st
    return out
