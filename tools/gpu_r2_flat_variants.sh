#!/bin/bash
# Round 2: flat_features_kernel: table copies x CTA size.
mkdir -p gpurun_out
cp grav1synth_b200/libg1s.so /tmp/libg1s_keep.so
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
for v in rowahead; do
  cp variants/libg1s_$v.so grav1synth_b200/libg1s.so
  ( timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/t_$v.log; show gpurun_out/t_$v.log "flat $v"
done
cp /tmp/libg1s_keep.so grav1synth_b200/libg1s.so
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 )
