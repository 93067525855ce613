#!/bin/bash
# One gpurun round: parity tests, smoke, bench, then ncu captures.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 ) > gpurun_out/bench.log
if [ "$1" == "all" ]; then
( timeout 600 python bench.py --workload 1080p8 --steps 20 --warmup 3 --frames 165 2>&1 | tail -1 ) > gpurun_out/bench_1080p8.log
( timeout 900 python bench.py --workload 8k10 --steps 10 --warmup 3 --frames 15 --cpu-frames 1 2>&1 | tail -1 ) > gpurun_out/bench_8k10.log
cat gpurun_out/bench_1080p8.log gpurun_out/bench_8k10.log
fi
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log
if [ "$1" == "prof" ]; then
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 20"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|gram_" -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_imma -s 1 -c 1 -f -o gpurun_out/prof_gram $CMD > gpurun_out/ncu_gram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_features -s 1 -c 1 -f -o gpurun_out/prof_flat $CMD > gpurun_out/ncu_flat.log 2>&1
fi
