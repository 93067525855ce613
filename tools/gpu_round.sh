#!/bin/bash
# One gpurun round: parity tests, smoke, bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 ) > gpurun_out/bench.log
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log
