#!/bin/bash
# Round 2: dynamic chunk scheduling + faster latest kernel: tests, bench in both model modes, stream variants.
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for hm in 1 0; do
for st in 1 3; do
( [ $hm = 0 ] && export G1S_DEVICE_MODEL=1; G1S_STREAMS=$st timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-strict --no-stats --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_hm${hm}_$st.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_hm${hm}_$st.log").read())
    print("host_model=$hm streams=$st value", round(d["value"]), {k:round(v*1000,1) for k,v in d["kernels"]["ms_per_frame_one_stream"].items() if k!="frames" and k!="frames_per_launch"})
except Exception as e: print("failed", e, open("gpurun_out/bench_hm${hm}_$st.log").read()[-400:])
PY
done
done
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"flat_|gram_|residual_|latest_" -s 7 -c 7 --csv --log-file gpurun_out/launches_n.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_n.csv")) if len(r)>10]
hdr=rows[0]; k=hdr.index("Kernel Name"); v=hdr.index("Metric Value")
for r in rows[1:]: print(r[k].split("(")[0][-30:], r[v])
PY
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:latest_kernel -s 1 -c 1 -f -o gpurun_out/prof_latest $CMD > gpurun_out/ncu_latest.log 2>&1
tail -1 gpurun_out/ncu_latest.log
