#!/bin/bash
# Round 2: the bench lines that go under profiles/ (after the last engine change).
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_4k10.json
( G1S_DEVICE_MODEL=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-strict 2>&1 | tail -1 ) > gpurun_out/bench_4k10_device_model.json
( timeout 900 python bench.py --steps 20 --warmup 3 --workload 1080p8 --repeat 16 --no-strict 2>&1 | tail -1 ) > gpurun_out/bench_1080p8.json
python - <<'PY'
import json
for f in ("4k10","4k10_device_model","1080p8"):
    d=json.loads(open(f"gpurun_out/bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "traffic", d["roofline"].get("traffic"), "sparse", (d.get("sparse_input") or {}).get("value"))
PY
