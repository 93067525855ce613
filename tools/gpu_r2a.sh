#!/bin/bash
# Round 2, first GPU pass: smoke, the new strict-mode tests, the parity suite with the row-pair Gram kernel, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 ) > gpurun_out/smoke.log
cat gpurun_out/smoke.log
( timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_parity.log
cat gpurun_out/pytest_parity.log
( timeout 1200 python -m pytest tests/test_gpu_strict.py -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_strict.log
cat gpurun_out/pytest_strict.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -5 ) > gpurun_out/bench.log
cat gpurun_out/bench.log
