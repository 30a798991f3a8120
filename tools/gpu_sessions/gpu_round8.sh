#!/bin/bash
# N-GPU call (N = $2): multi-GPU parity of every transport + the N-GPU bench line
tag=${1:-run}; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
for tr in auto peer fetch nccl; do
  ( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py $tr > gpurun_out/${tag}_dist_check_N${N}_${tr}.log 2>&1; echo "dist_check $tr rc=$?" )
  grep -E "DIST_CHECK|MISMATCH|False|Error" gpurun_out/${tag}_dist_check_N${N}_${tr}.log | head -5
done
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --phases > gpurun_out/${tag}_bench_N${N}.json 2> gpurun_out/${tag}_bench_N${N}.err; echo "bench N$N rc=$?" )
grep "per-minibatch" gpurun_out/${tag}_bench_N${N}.err
