#!/bin/bash
# End-of-round evidence on one B200: parity tests, smoke, the three bench workloads, the ncu launch list.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/bench_4k10.json
( timeout 600 python bench.py --workload 1080p8 --steps 20 --warmup 3 --frames 165 2>&1 | tail -1 ) > gpurun_out/bench_1080p8.json
( timeout 900 python bench.py --workload 8k10 --steps 10 --warmup 3 --frames 15 --cpu-frames 1 2>&1 | tail -1 ) > gpurun_out/bench_8k10.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/bench_reference.json
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --frames 20"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|gram_|residual_" -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log
for f in gpurun_out/bench_*.json; do echo "== $f"; cut -c1-1500 $f; done
