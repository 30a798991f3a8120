#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -12 gpurun_out/${tag}_pytest.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > gpurun_out/${tag}_bench_N1.json 2> gpurun_out/${tag}_bench_N1.err; echo "bench N1 rc=$?" )
tail -c 400 gpurun_out/${tag}_bench_N1.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/'+__import__('sys').argv[1]+'_bench_N1.json').read().strip().splitlines()[-1]) if False else None
PY
