#!/bin/bash
# Round 2: 8-GPU scaling preview (bench only, short).
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
for n in 8 4; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-e2e-variants 2>&1 | tail -3 ) > gpurun_out/bench_n$n.log
python - <<PY
import json
txt=open("gpurun_out/bench_n$n.log").read()
try:
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("N=$n value", round(d["value"]), "per gpu", round(d["value"]/$n), "parity", d.get("parity_checked"), "e2e", round(d["e2e"]["value"]))
except Exception as e: print("failed", e, txt[-1500:])
PY
done
