#!/bin/bash
# Quick gpurun round for a new kernel: smoke first (bounded), then parity tests, then a short bench.
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 ) > gpurun_out/smoke.log
cat gpurun_out/smoke.log
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 ) > gpurun_out/bench.log
cat gpurun_out/bench.log
