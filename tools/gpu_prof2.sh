#!/bin/bash
# ncu captures of the residual, gram and flat kernels (1 GPU): launch list + full sections.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 20"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|gram_|residual_" -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
for k in ${KERNELS:-gram_imma residual}; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out
