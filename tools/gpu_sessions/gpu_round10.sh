#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_als.py -q -v > gpurun_out/${tag}_pytest_als.log 2>&1; echo "pytest rc=$?" )
tail -15 gpurun_out/${tag}_pytest_als.log
python tools/als_prof.py > gpurun_out/${tag}_als_plain.log 2>&1; tail -2 gpurun_out/${tag}_als_plain.log
CF_ALS_DIRECT=1 python tools/als_prof.py > gpurun_out/${tag}_als_plain_direct.log 2>&1; tail -1 gpurun_out/${tag}_als_plain_direct.log
ncu --set full --clock-control none --import-source on -k regex:k_als_ -c 9 -o gpurun_out/${tag}_als python tools/als_prof.py > gpurun_out/${tag}_als_ncu.log 2>&1
python tools/als_prof.py big > gpurun_out/${tag}_als_big.log 2>&1; tail -2 gpurun_out/${tag}_als_big.log
