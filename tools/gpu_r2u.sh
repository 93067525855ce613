#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_features -s 1 -c 1 -f -o gpurun_out/prof_flat_u $CMD > gpurun_out/ncu_flat.log 2>&1
tail -1 gpurun_out/ncu_flat.log
