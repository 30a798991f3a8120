#!/bin/bash
# the driver's round-end sequence on one GPU: smoke, the reference arm, the default bench
tag=${1:-run}
mkdir -p gpurun_out
( timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ); tail -2 gpurun_out/${tag}_smoke.log
t0=$(date +%s)
( timeout 1200 python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$? $(( $(date +%s) - t0 )) s" )
t0=$(date +%s)
( timeout 1500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s" )
tail -c 600 gpurun_out/${tag}_bench_ref.json; echo; tail -3 gpurun_out/${tag}_bench.err
