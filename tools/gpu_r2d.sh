#!/bin/bash
# Round 2: stream / occupancy / batch matrix + strict kernel timing.
mkdir -p gpurun_out
run() { # name, env..., args
  name=$1; shift
  ( env "$@" timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --no-strict $EXTRA 2>&1 | tail -1 ) > gpurun_out/m_$name.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/m_$name.log").read())
    print("$name", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k[10:]:round(v,3) for k,v in d["config"].items() if k.startswith("device_ms")})
except Exception as e: print("$name failed", e, open("gpurun_out/m_$name.log").read()[-300:])
PY
}
run s1 G1S_STREAMS=1
run s2 G1S_STREAMS=2
run s3 G1S_STREAMS=3
run s2_occ1 G1S_STREAMS=2 G1S_GRAM_OCC=1
run s3_occ1 G1S_STREAMS=3 G1S_GRAM_OCC=1
EXTRA="--batch 10" run s3_b10 G1S_STREAMS=3
EXTRA="--batch 10" run s4_b10_occ1 G1S_STREAMS=4 G1S_GRAM_OCC=1
EXTRA="--batch 30" run s2_b30 G1S_STREAMS=2
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_strict.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_strict.log').read()); print('strict', d.get('value_strict'), d.get('strict'))"
( timeout 900 python -m pytest tests/test_gpu_strict.py -x -q -m gpu 2>&1 | tail -4 )
