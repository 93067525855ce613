#!/bin/bash
# Round 2: full GPU suite + default bench after the latest / flat / e2e changes.
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
start=$(date +%s)
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/bench_default.log
echo "bench wall seconds: $(( $(date +%s) - start ))"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","roofline","kernels","value_strict","e2e","e2e_pageable","e2e_host_narrow","e2e_pinned_direct","gpu_launches","clocks"):
    print(k, ":", json.dumps(d.get(k))[:420])
PY
