#!/bin/bash
# One gpurun call's worth of checks: each step under its own timeout, outputs under gpurun_out/<tag>_*.log
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -5 gpurun_out/${tag}_pytest.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" )
tail -2 gpurun_out/${tag}_smoke.log
( timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" )
tail -c 600 gpurun_out/${tag}_bench.err
( timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?" )
