#!/bin/bash
# Round 2: device-side per-frame model + source filters: tests, bench.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_latest_gpu.py tests/test_filters_gpu.py -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_new.log
cat gpurun_out/pytest_new.log
( timeout 1500 python -m pytest tests -x -q -m gpu  2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for hm in 0 1; do
( [ $hm = 1 ] && export G1S_HOST_MODEL=1; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict --no-stats --no-e2e-variants 2>&1 | tail -1 ) > gpurun_out/bench_hm$hm.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_hm$hm.log").read())
    print("host_model=$hm value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v*1000,1) for k,v in d["kernels"]["ms_per_frame_one_stream"].items() if k!="frames" and k!="frames_per_launch"})
except Exception as e: print("failed", e, open("gpurun_out/bench_hm$hm.log").read()[-400:])
PY
done
