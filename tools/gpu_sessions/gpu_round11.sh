#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_als.py tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py tests/test_gpu_full_size.py tests/test_gpu_dist.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -6 gpurun_out/${tag}_pytest.log
python tools/als_prof.py > gpurun_out/${tag}_als_plain.log 2>&1; tail -1 gpurun_out/${tag}_als_plain.log
python tools/als_prof.py big > gpurun_out/${tag}_als_big.log 2>&1; tail -1 gpurun_out/${tag}_als_big.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/${tag}_bench_N1.json 2> gpurun_out/${tag}_bench_N1.err; echo "bench N1 rc=$?" )
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_N1.json',):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); t=j['topk']; c=t['c5_catalogue']
        print(f, '500k: %.0f users/s (%.3f)  10M: %.0f users/s (%.3f burst, %.3f sustained) fb %d/%d cand %.0f/%.0f' % (t['value'], t['frac_of_tensor_peak'], c['value'], c['frac_of_tensor_peak'], c['frac_of_sustained_tensor_peak'], t['fallback_rows'], c['fallback_rows'], t['candidates_per_row'], c['candidates_per_row']))
    except Exception as e: print(f, 'failed', e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_topk|k_prep|k_rerank" -c 12 --csv --log-file gpurun_out/${tag}_topk_launches.csv python tools/topk_perf.py cml 37888 500000 128 1 > gpurun_out/${tag}_topk_launch.log 2>&1
tail -3 gpurun_out/${tag}_topk_launch.log
