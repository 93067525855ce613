#!/bin/bash
# Multi-GPU bench only (no reference arm), N = $1
N=${1:-2}
mkdir -p gpurun_out
nproc > gpurun_out/nproc_n$N.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>&1 | tail -4 ) > gpurun_out/bench_n$N.log
cat gpurun_out/nproc_n$N.log; cut -c1-700 gpurun_out/bench_n$N.log
