#!/bin/bash
# ncu captures: launch list + full sections for the gram and flat kernels (1 GPU).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 16"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_imma -s 1 -c 1 -f -o gpurun_out/prof_gram $CMD > gpurun_out/ncu_gram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_features -s 1 -c 1 -f -o gpurun_out/prof_flat $CMD > gpurun_out/ncu_flat.log 2>&1
tail -3 gpurun_out/ncu_launch.log gpurun_out/ncu_gram.log gpurun_out/ncu_flat.log
ls -la gpurun_out
