#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -12 gpurun_out/${tag}_pytest.log
( cd tests/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../collaborativefilteringusingtensorflow_b200/csrc -I ../../include -o mma_micro mma_micro.cu -lcuda && timeout 200 ./mma_micro > ../../gpurun_out/${tag}_mma_micro.log 2>&1; echo "mma_micro rc=$?" )
cat gpurun_out/${tag}_mma_micro.log
