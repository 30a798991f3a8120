#!/bin/bash
# 2-GPU call: world-1 + N=2 parity of every transport, NVLink micro-benchmark, N=2 bench
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -x -q > gpurun_out/${tag}_pytest_dist.log 2>&1; echo "pytest dist rc=$?" )
tail -15 gpurun_out/${tag}_pytest_dist.log
( cd tests/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_micro peer_micro.cu && timeout 120 ./peer_micro > ../../gpurun_out/${tag}_peer_micro.log 2>&1; echo "peer_micro rc=$?" )
cat gpurun_out/${tag}_peer_micro.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --phases > gpurun_out/${tag}_bench_N2.json 2> gpurun_out/${tag}_bench_N2.err; echo "bench N2 rc=$?" )
tail -c 1500 gpurun_out/${tag}_bench_N2.err
( timeout 600 python -m pytest tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest_full.log 2>&1; echo "pytest full rc=$?" )
tail -5 gpurun_out/${tag}_pytest_full.log
