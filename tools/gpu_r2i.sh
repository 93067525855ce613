#!/bin/bash
# Round 2: y8 flat path + residual-first order, warp-aggregated plan kernel: parity, launch list, variants.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strict.py tests/test_multi_device_gpu.py tests/test_zz_workflow_gpu.py tests/test_zz_host_narrow.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"flat_|gram_|residual_" -s 6 -c 6 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>10]
hdr=rows[0]; k=hdr.index("Kernel Name"); m=hdr.index("Metric Name"); v=hdr.index("Metric Value"); i=hdr.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((r[i], r[k].split("(")[0][-28:]), {})[r[m]]=r[v]
for key,val in out.items(): print(key, val)
PY
for v in "8 4" "8 3" "6 3" "10 5"; do
  set -- $v
  G1S_EXTRA_NVCC="-DG1S_GRAM_WARPS=$1 -DG1S_LUMA_WARPS=$2" python -m grav1synth_b200.build --force > /dev/null 2>&1
  for st in 1 3; do
  ( G1S_STREAMS=$st timeout 600 python bench.py --steps 5 --warmup 2 --repeat 4 --no-cpu-baseline --no-e2e --no-strict --no-stats 2>&1 | tail -1 ) > gpurun_out/v_$1_$2_$st.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$1_$2_$st.log").read())
    print("warps $1 luma $2 streams $st: value", round(d["value"]), {k:round(v*1000,1) for k,v in d["kernels"]["ms_per_frame_one_stream"].items() if k!="frames" and k!="frames_per_launch"})
except Exception as e: print("variant $v failed", e, open("gpurun_out/v_$1_$2_$st.log").read()[-300:])
PY
  done
done
python -m grav1synth_b200.build --force > /dev/null 2>&1
