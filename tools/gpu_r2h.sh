#!/bin/bash
# Round 2: ncu of the plan kernel; luma/chroma warp split experiment.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --repeat 1 --no-e2e --no-cpu-baseline --no-strict --no-stats --frames 20"
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_plan -s 1 -c 1 -f -o gpurun_out/prof_plan $CMD > gpurun_out/ncu_plan.log 2>&1
tail -2 gpurun_out/ncu_plan.log
for v in "8 5" "8 6" "8 4" "12 8" "12 7"; do
  set -- $v
  G1S_EXTRA_NVCC="-DG1S_GRAM_WARPS=$1 -DG1S_LUMA_WARPS=$2" python -m grav1synth_b200.build --force > /dev/null 2>&1
  ( G1S_STREAMS=1 timeout 600 python bench.py --steps 5 --warmup 2 --repeat 4 --no-cpu-baseline --no-e2e --no-strict --no-stats 2>&1 | tail -1 ) > gpurun_out/v_$1_$2.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_$1_$2.log").read())
    print("warps $1 luma $2: value", round(d["value"]), {k:round(v*1000,1) for k,v in d["kernels"]["ms_per_frame_one_stream"].items() if k!="frames" and k!="frames_per_launch"})
except Exception as e: print("variant $v failed", e, open("gpurun_out/v_$1_$2.log").read()[-300:])
PY
done
python -m grav1synth_b200.build --force > /dev/null 2>&1
