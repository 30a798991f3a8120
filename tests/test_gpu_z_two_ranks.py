"""N > 1 parity inside the suite the driver runs: spawns ``torch.distributed.run`` over min(2, device_count) GPUs on
tests/dist_check.py (sharded BPR / CML training == one GPU on the global minibatch, sharded top-K + merge == one GPU,
sharded metrics, sharded ALS, replicated GBPR vs the fp64-summed oracle) for every item transport.  On a one-GPU box (the
driver's) the two ranks share GPU 0: NCCL refuses that, so the script runs over gloo with its collectives staged through
the host (tests/dist_check.py::stage_collectives_through_host) -- two processes, CUDA-IPC mappings of each other's
buffers, the same kernels, plan and exchange as on two GPUs.  tests/test_dist_gloo.py covers the host logic at world
size 2 and 3 on CPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('transport', ['nccl', 'peer', 'peer-push', 'fetch', 'replicate', 'auto'])
def test_two_rank_parity_under_torchrun(transport):
    import torch
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env['CF_DIST_BACKEND'] = 'gloo'
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'dist_check.py'), transport]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)      # 7-10 s when healthy
    assert r.returncode == 0 and 'DIST_CHECK OK' in r.stdout, r.stdout[-4000:] + r.stderr[-4000:]
