#!/bin/bash
# Round 2 evidence: launch list (durations + DRAM bytes) of one-stream batches, full ncu of every hot kernel, bench lines.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"flat_|gram_|residual_|latest_" -s 14 -c 14 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
for k in gram_imma flat_features residual gram_plan latest; do
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 1 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
tail -1 gpurun_out/ncu_$k.log
done
G1S_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_reforder -s 0 -c 1 -f -o gpurun_out/prof_strict python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-stats --frames 20 --strict-steps 1 > gpurun_out/ncu_strict.log 2>&1
tail -1 gpurun_out/ncu_strict.log
( timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_4k10.json
( timeout 900 python bench.py --steps 20 --warmup 3 --workload 1080p8 --repeat 16 --no-strict 2>&1 | tail -1 ) > gpurun_out/bench_1080p8.json
( timeout 900 python bench.py --steps 10 --warmup 3 --workload 8k10 --frames 16 --repeat 8 --no-strict --no-stats --no-e2e-variants 2>&1 | tail -1 ) > gpurun_out/bench_8k10.json
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/bench_reference.json
ls -la gpurun_out | head -40
