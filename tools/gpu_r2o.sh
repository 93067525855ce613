#!/bin/bash
# Round 2: seamless chunk producer (chunk size variants), latest kernel v3 (side-by-side eliminations, pipelined chains).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_latest_gpu.py tests/test_gpu_parity.py tests/test_multi_device_gpu.py -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/pytest_o.log
cat gpurun_out/pytest_o.log
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
for c in 8 16 32; do
  cp variants/libg1s_c$c.so grav1synth_b200/libg1s.so
  ( G1S_HOST_MODEL=1 timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/o_c${c}_host.log; show gpurun_out/o_c${c}_host.log "chunk=$c host-model"
done
cp variants/libg1s_c16.so grav1synth_b200/libg1s.so
( G1S_DEVICE_MODEL=1 timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/o_dev3.log; show gpurun_out/o_dev3.log "chunk=16 device-model streams=3"
( G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 $B 2>&1 | tail -1 ) > gpurun_out/o_dev1.log; show gpurun_out/o_dev1.log "chunk=16 device-model streams=1"
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|gram_|residual_|latest_" -s 8 -c 8 --csv --log-file gpurun_out/launches_o.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_o.csv")) if len(r)>10]
hdr=rows[0]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value")
for r in rows[1:]: print(r[k].split("(")[0][-30:], r[v])
PY
G1S_DEVICE_MODEL=1 G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:latest_kernel -s 1 -c 1 -f -o gpurun_out/prof_latest $CMD > gpurun_out/ncu_latest.log 2>&1
tail -1 gpurun_out/ncu_latest.log
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_imma -s 1 -c 1 -f -o gpurun_out/prof_gram_o $CMD > gpurun_out/ncu_gram.log 2>&1
tail -1 gpurun_out/ncu_gram.log
