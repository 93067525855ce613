#!/bin/bash
# Round 2: stream count, and batch size, after the kernel changes.
mkdir -p gpurun_out
show() {
python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", round(d["value"]))
except Exception as e: print(sys.argv[2], "failed", e, open(sys.argv[1]).read()[-300:])
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict --no-stats --no-e2e"
for st in 2 3 4; do
  ( G1S_STREAMS=$st timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/y_s$st.log; show gpurun_out/y_s$st.log "streams=$st"
done
for bt in 10 30 40; do
  ( timeout 600 $B --batch $bt 2>&1 | tail -1 ) > gpurun_out/y_b$bt.log; show gpurun_out/y_b$bt.log "batch=$bt"
done
( G1S_DEVICE_MODEL=1 timeout 600 $B --batch 40 2>&1 | tail -1 ) > gpurun_out/y_d40.log; show gpurun_out/y_d40.log "device model batch=40"
( G1S_DEVICE_MODEL=1 timeout 600 $B --batch 30 2>&1 | tail -1 ) > gpurun_out/y_d30.log; show gpurun_out/y_d30.log "device model batch=30"
