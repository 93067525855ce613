#!/bin/bash
# Round 2: full default bench after the staging-allocation fix; flat-kernel prefetch variants.
mkdir -p gpurun_out
for pf in 0 1 2; do
  G1S_EXTRA_NVCC="-DG1S_FLAT_PF=$pf" python -m grav1synth_b200.build --force > /dev/null 2>&1
  ( G1S_STREAMS=1 timeout 600 python bench.py --steps 4 --warmup 2 --repeat 4 --no-cpu-baseline --no-e2e --no-strict --no-stats 2>&1 | tail -1 ) > gpurun_out/pf_$pf.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pf_$pf.log").read())
    print("flat pf $pf: value", round(d["value"]), {k:round(v*1000,1) for k,v in d["kernels"]["ms_per_frame_one_stream"].items() if k!="frames" and k!="frames_per_launch"})
except Exception as e: print("variant failed", e, open("gpurun_out/pf_$pf.log").read()[-300:])
PY
done
python -m grav1synth_b200.build --force > /dev/null 2>&1
start=$(date +%s)
( timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_full.log
echo "bench wall seconds: $(( $(date +%s) - start ))"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.log").read())
for k in ("value","ms_per_step","roofline","kernels","value_strict","e2e","e2e_pageable","e2e_host_narrow","sparse_input","workload"):
    print(k, ":", json.dumps(d.get(k))[:500])
PY
