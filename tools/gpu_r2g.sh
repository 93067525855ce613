#!/bin/bash
# Round 2: parallel plan kernel: parity, launch list, full default bench (timed by the shell).
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strict.py tests/test_multi_device_gpu.py tests/test_zz_workflow_gpu.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"flat_|gram_|residual_" -s 12 -c 12 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>10]
hdr=rows[0]; k=hdr.index("Kernel Name"); m=hdr.index("Metric Name"); v=hdr.index("Metric Value"); i=hdr.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((r[i], r[k].split("(")[0][-28:]), {})[r[m]]=r[v]
for key,val in out.items(): print(key, val)
PY
start=$(date +%s)
( timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_full.log
echo "bench wall seconds: $(( $(date +%s) - start ))"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_full.log").read())
for k,v in d.items(): print(k, ":", json.dumps(v)[:600])
PY
