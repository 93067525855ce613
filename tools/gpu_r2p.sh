#!/bin/bash
# Round 2: latest kernel v3b (row factors in the scan warp, shared-space chain loads, padded lists).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_latest_gpu.py tests/test_multi_device_gpu.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/pytest_p.log
cat gpurun_out/pytest_p.log
show() {
python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", round(d["value"]), {k:round(v*1000,1) for k,v in d["kernels"]["ms_per_frame_one_stream"].items() if k!="frames" and k!="frames_per_launch"})
except Exception as e: print(sys.argv[2], "failed", e, open(sys.argv[1]).read()[-400:])
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict --no-stats --no-e2e"
( G1S_HOST_MODEL=1 timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/p_host.log; show gpurun_out/p_host.log "host-model"
( G1S_DEVICE_MODEL=1 timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/p_dev3.log; show gpurun_out/p_dev3.log "device-model streams=3"
( G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/p_dev1.log; show gpurun_out/p_dev1.log "device-model streams=1"
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"latest_" -s 1 -c 1 --csv --log-file gpurun_out/launches_p.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/launches_p.csv
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:latest_kernel -s 1 -c 1 -f -o gpurun_out/prof_latest $CMD > gpurun_out/ncu_latest.log 2>&1
tail -1 gpurun_out/ncu_latest.log
