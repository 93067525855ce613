#!/bin/bash
# Multi-GPU bench (one process per GPU under torchrun), N = $1
N=${1:-2}
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -8 ) > gpurun_out/bench_n$N.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -3 ) > gpurun_out/bench_ref_n$N.log
cat gpurun_out/bench_n$N.log gpurun_out/bench_ref_n$N.log
