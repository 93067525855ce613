#!/bin/bash
# Round 2: two-GPU checks: NCCL sharded parity test, torchrun bench with parity_checked, multi-device handle on distinct GPUs.
mkdir -p gpurun_out
nvidia-smi -L
( timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_multi_device_gpu.py -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/pytest_2gpu.log
cat gpurun_out/pytest_2gpu.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench_n2.log
python - <<'PY'
import json
txt=open("gpurun_out/bench_n2.log").read()
try:
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    for k in ("value","n_gpus","ms_per_step","parity_checked","parity_segments","roofline","e2e","config"):
        print(k, ":", json.dumps(d.get(k))[:400])
except Exception as e: print("failed", e, txt[-1500:])
PY
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-strict --no-stats --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_n1.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.log').read()); print('N=1 value', d['value'])"
