#!/bin/bash
# Round 2: 8-GPU scaling with the per-frame model half on the device (ShardedDiff's placement rule), and forced host.
mkdir -p gpurun_out
nproc
for mode in auto host; do
  ( [ $mode = host ] && export G1S_HOST_MODEL=1; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e-variants 2>&1 | tail -1 ) > gpurun_out/x_n8_$mode.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/x_n8_$mode.log").read().strip().splitlines()[-1])
    print("N=8 $mode:", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_checked"), d["config"].get("parallelism","")[:80])
except Exception as e: print("failed", e, open("gpurun_out/x_n8_$mode.log").read()[-600:])
PY
done
