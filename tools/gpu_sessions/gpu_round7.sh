#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_als.py tests/test_gpu_tuples.py tests/test_gpu_neighbors.py tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py tests/test_gpu_full_size.py tests/test_gpu_e2e.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -25 gpurun_out/${tag}_pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/${tag}_bench_N1.json 2> gpurun_out/${tag}_bench_N1.err; echo "bench N1 rc=$?" )
tail -c 300 gpurun_out/${tag}_bench_N1.err
( CF_TC_NARROW=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/${tag}_bench_N1_narrow.json 2> gpurun_out/${tag}_bench_N1_narrow.err; echo "bench N1 narrow rc=$?" )
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_N1.json','gpurun_out/${tag}_bench_N1_narrow.json'):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); t=j['topk']; c=t['c5_catalogue']
        print(f, '500k: %.0f users/s (%.3f)  10M: %.0f users/s (%.3f burst, %.3f sustained) fb %d/%d' % (t['value'], t['frac_of_tensor_peak'], c['value'], c['frac_of_tensor_peak'], c['frac_of_sustained_tensor_peak'], t['fallback_rows'], c['fallback_rows']))
    except Exception as e: print(f, 'failed', e)
PY
( timeout 300 python - <<'PY' > gpurun_out/${tag}_als.log 2>&1
import json, os, sys, torch
sys.path.insert(0, '.')
import bench
pk = bench.peaks()
dev = torch.device('cuda', 0)
for env in ('', '1'):
    if env: os.environ['CF_ALS_DIRECT'] = env
    r = bench.als_slice(bench.ALS_SLICE, dev, pk)
    print('direct' if env else 'woodbury+direct', r['ms_per_half_sweep'], 'ms per 1M-row half-sweep')
PY
cat gpurun_out/${tag}_als.log | tail -3 )
