#!/bin/bash
# Round 2, second GPU pass: whole GPU suite, bench with the strict pass, launch list, full ncu of gram + flat.
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/bench.log
cat gpurun_out/bench.log
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-strict --frames 20"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|gram_|residual_" -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
for k in gram_imma flat_features; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
