#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict --no-e2e 2>&1 | tail -1 ) > gpurun_out/z.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/z.log").read().strip().splitlines()[-1])
print("value", round(d["value"]), "sparse", d.get("sparse_input"))
PY
( G1S_DEVICE_MODEL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict --no-e2e 2>&1 | tail -1 ) > gpurun_out/z2.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/z2.log").read().strip().splitlines()[-1])
print("device model: value", round(d["value"]), "sparse", d.get("sparse_input"))
PY
