#!/bin/bash
# full -m gpu suite + compute-sanitizer memcheck / racecheck of the small all-kernels target
tag=${1:-run}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_all.log 2>&1; echo "pytest rc=$?" )
tail -5 gpurun_out/${tag}_pytest_all.log
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?" )
tail -8 gpurun_out/${tag}_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/${tag}_racecheck.log 2>&1; echo "racecheck rc=$?" )
tail -8 gpurun_out/${tag}_racecheck.log
