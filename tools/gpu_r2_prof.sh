#!/bin/bash
# Round 2 evidence: launch list (durations + DRAM bytes) of one-stream batches, full ncu of every hot kernel, bench lines.
mkdir -p gpurun_out
F=41  # frames per launch: the engine's default batch at 4K 10-bit
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames $F"
G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"flat_|gram_|residual_" -s 12 -c 12 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"latest_" -s 1 -c 2 --csv --log-file gpurun_out/launches_latest.csv $CMD > gpurun_out/ncu_launch_latest.log 2>&1
for k in gram_imma flat_features residual gram_plan; do
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s 1 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
tail -1 gpurun_out/ncu_$k.log
done
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:latest_kernel -s 1 -c 1 -f -o gpurun_out/prof_latest $CMD > gpurun_out/ncu_latest.log 2>&1
tail -1 gpurun_out/ncu_latest.log
G1S_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_reforder -s 0 -c 1 -f -o gpurun_out/prof_strict python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-stats --frames 20 --strict-steps 1 > gpurun_out/ncu_strict.log 2>&1
tail -1 gpurun_out/ncu_strict.log
./tools/chain_probe > gpurun_out/chain_probe.txt 2>&1
cp variants/libg1s_prof.so /tmp/libg1s_prof.so 2>/dev/null && cp grav1synth_b200/libg1s.so /tmp/libg1s_keep.so && cp /tmp/libg1s_prof.so grav1synth_b200/libg1s.so && \
  ( G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 $CMD 2>&1 | grep "^latest" | tail -24 > gpurun_out/latest_phases.log ); cp /tmp/libg1s_keep.so grav1synth_b200/libg1s.so
( timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_4k10.json
( G1S_DEVICE_MODEL=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-strict 2>&1 | tail -1 ) > gpurun_out/bench_4k10_device_model.json
( timeout 900 python bench.py --steps 20 --warmup 3 --workload 1080p8 --repeat 16 --no-strict 2>&1 | tail -1 ) > gpurun_out/bench_1080p8.json
( timeout 900 python bench.py --steps 10 --warmup 3 --workload 8k10 --frames 16 --repeat 8 --no-strict --no-stats --no-e2e-variants 2>&1 | tail -1 ) > gpurun_out/bench_8k10.json
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/bench_reference.json
ls -la gpurun_out | head -50
