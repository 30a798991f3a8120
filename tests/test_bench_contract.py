"""bench.py's contract, as far as a box without a GPU can check it: the reference arm (the CPU restatement: the one place
bench.py may execute oracle/) prints ONE JSON line with the keys the driver reads and the same `metric` / `config` objects
the GPU arm prints; under torchrun only rank 0 works; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ['--workload', 'small', '--steps', '2', '--warmup', '1', '--batch', '4096', '--cpu-batch', '4096', '--cpu-budget', '2',
         '--faithful-budget', '1']


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, ROOT)
    import bench
    r = _run(['--impl', 'reference'] + SMALL)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == bench.metric_name(128) and j['unit'] == 'triple updates/s'
    assert j['higher_is_better'] is True and j['n_gpus'] == 1 and j['steps'] == 2 and j['vs_baseline'] is None and j['dtype'] == 'f32'
    assert j['value'] > 0 and abs(j['ms_per_step'] * j['value'] / 1e3 - 4096 * 5) < 1e-6 * 4096 * 5      # triples of one step / its time
    args = type('A', (), dict(batch=4096, optimizer='adagrad', update='sync'))
    assert j['config'] == bench.same_config(bench.WORKLOADS['small'], args)                              # the object the GPU arm prints
    cb = j['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == j['value'] and cb['cores'] >= 1 and 'sample' in cb
    assert j['e2e'] == dict(value=j['value'], unit=j['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    f = j['cpu_baseline_faithful']                                                                       # testbprmf.py:21-30: B = 100, one stream
    assert f['batch_pairs'] == 100 and f['setting'] == 'reference-faithful' and f['value'] > 0


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(['--impl', 'reference', '--gpus', '2'] + SMALL, env=dict(RANK='1', LOCAL_RANK='1', WORLD_SIZE='2'))
    assert r.returncode == 0 and r.stdout.strip() == '', (r.stdout[-500:], r.stderr[-500:])


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip('a GPU is present')
    r = _run(['--workload', 'small', '--steps', '2', '--warmup', '1'])
    assert r.returncode != 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert 'no CPU fallback' in (r.stdout + r.stderr)
