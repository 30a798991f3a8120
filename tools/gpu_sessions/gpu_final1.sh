#!/bin/bash
# round-end sequence on one GPU: full -m gpu suite, smoke, reference arm, default bench, launch list of the bench
tag=${1:-run}
mkdir -p gpurun_out
t0=$(date +%s)
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_all.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - t0 )) s" ); tail -3 gpurun_out/${tag}_pytest_all.log
( timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ); tail -1 gpurun_out/${tag}_smoke.log | cut -c1-200
( timeout 1200 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?" )
t0=$(date +%s)
( timeout 1500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s" ); tail -2 gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs --topk-users 0 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
