#!/bin/bash
# re-rank: two candidates per thread, checked entries skip the mask search: tests + timing
tag=${1:-r2X}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py tests/test_gpu_full_size.py tests/test_gpu_e2e.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -3 gpurun_out/${tag}_pytest.log
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
( timeout 600 python tools/topk_perf.py bpr 200000 500000 128 2 10 >> gpurun_out/${tag}_perf.log 2>&1 )
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 200 >> gpurun_out/${tag}_perf.log 2>&1 )
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 1 1000 >> gpurun_out/${tag}_perf.log 2>&1 )
grep -v fallback gpurun_out/${tag}_perf.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --topk-c5-items 0 > gpurun_out/${tag}_bench_topk.json 2> gpurun_out/${tag}_bench_topk.err; echo "bench rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_topk.json').read().strip().splitlines()[-1])
t=j['topk']
print('bench topk', t['value'], t['ms'], t['frac_of_tensor_peak'], t['candidates_per_row'])
PY
