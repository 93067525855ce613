#!/bin/bash
# Round 2: plan kernel + strip-ordered Gram kernel: parity, bench, ncu of gram_imma and of the strict kernel.
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strict.py tests/test_zz_workflow_gpu.py -x -q -m gpu 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for s in 1 3; do
( G1S_STREAMS=$s timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-strict 2>&1 | tail -1 ) > gpurun_out/bench_streams$s.log
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_streams$s.log").read())
    print("streams=$s value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["config"].items() if k.startswith("device_ms")})
except Exception as e: print("bench failed", e, open("gpurun_out/bench_streams$s.log").read()[-400:])
PY
done
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-strict --frames 20"
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_imma -s 1 -c 1 -f -o gpurun_out/prof_gram_imma $CMD > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log
G1S_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_reforder -s 0 -c 1 -f -o gpurun_out/prof_strict python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --frames 20 --strict-steps 1 > gpurun_out/ncu_strict.log 2>&1
tail -2 gpurun_out/ncu_strict.log
