#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" )
tail -8 gpurun_out/${tag}_pytest.log
( cd tests/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../collaborativefilteringusingtensorflow_b200/csrc -I ../../include -o mma_micro mma_micro.cu -lcuda && timeout 120 ./mma_micro > ../../gpurun_out/${tag}_mma_micro.log 2>&1; echo "mma_micro rc=$?" )
cat gpurun_out/${tag}_mma_micro.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --phases > gpurun_out/${tag}_bench_N2.json 2> gpurun_out/${tag}_bench_N2.err; echo "bench N2 rc=$?" )
grep "per-minibatch" gpurun_out/${tag}_bench_N2.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > gpurun_out/${tag}_bench_N1.json 2> gpurun_out/${tag}_bench_N1.err; echo "bench N1 rc=$?" )
tail -c 400 gpurun_out/${tag}_bench_N1.err
