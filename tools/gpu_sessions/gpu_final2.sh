#!/bin/bash
# the driver's round-end sequence on one GPU: smoke, pytest -m gpu, default bench, reference arm
tag=${1:-r2F}
mkdir -p gpurun_out
( timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ); tail -2 gpurun_out/${tag}_smoke.log
( timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" ); tail -3 gpurun_out/${tag}_pytest_gpu.log
( timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" )
( timeout 900 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
r=j['roofline']
print('value %.3f G ms %.3f frac %.3f whole %.3f e2e %.3f G launches %s' % (j['value']/1e9, j['ms_per_step'], r['frac'], r['whole_step_frac'], j['e2e']['value']/1e9, j['gpu_launches']))
print('cpu', (j.get('cpu_baseline') or {}).get('value'), (j.get('cpu_baseline_faithful') or {}).get('value'))
t=j['topk']; print('topk 500k', t['value'], t['frac_of_tensor_peak'], 'c5', t['c5_catalogue']['value'], t['c5_catalogue']['frac_of_tensor_peak'])
for k,v in j['other_configs'].items():
    print('   ', k, v.get('value'), v.get('ms_per_step', v.get('ms_per_half_sweep')), (v.get('roofline') or {}).get('whole_step_frac'), (v.get('roofline') or {}).get('kernel'))
print('clocks', j['clocks'])
jr=json.loads(open('gpurun_out/${tag}_bench_ref.json').read().strip().splitlines()[-1])
print('ref', jr.get('value'), jr.get('metric')==j['metric'], jr.get('cpu_baseline'))
PY
