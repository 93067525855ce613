#!/bin/bash
# Round 2: bench.py --gpus N under torchrun (the driver's scaling launch), N from the command line.
N=$1
mkdir -p gpurun_out
if [ "$N" = 2 ]; then ( timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_multi_device_gpu.py -x -q -m gpu 2>&1 | tail -3 ) > gpurun_out/pytest_2gpu.log; cat gpurun_out/pytest_2gpu.log; fi
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity_checked", d.get("parity_checked"), "placement", d["config"].get("model_placement"))
PY
