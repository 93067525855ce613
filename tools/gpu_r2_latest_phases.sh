#!/bin/bash
# Round 2: phase clocks of latest_kernel (instrumented build, not the product build).
mkdir -p gpurun_out
cp grav1synth_b200/libg1s.so /tmp/libg1s_keep.so
cp variants/libg1s_prof.so grav1synth_b200/libg1s.so
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20 2>&1 | grep "^latest" | tail -60 > gpurun_out/latest_phases.log
cp /tmp/libg1s_keep.so grav1synth_b200/libg1s.so
tail -30 gpurun_out/latest_phases.log
