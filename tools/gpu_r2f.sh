#!/bin/bash
# Round 2: plan kernel from shared memory, multi-device handle tests, strict timing.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strict.py tests/test_multi_device_gpu.py tests/test_zz_workflow_gpu.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for s in 1 3; do
( G1S_STREAMS=$s timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $( [ $s = 1 ] && echo --no-strict ) 2>&1 | tail -1 ) > gpurun_out/bench_streams$s.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_streams$s.log").read())
    print("streams=$s value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["config"].items() if k.startswith("device_ms")}, "strict", d.get("value_strict"))
except Exception as e: print("bench failed", e, open("gpurun_out/bench_streams$s.log").read()[-400:])
PY
done
