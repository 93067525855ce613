#!/bin/bash
# First gpurun of the next round: everything that was written after the round-1 GPU budget ran out.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh'
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
cat gpurun_out/smoke.log
( timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log     # no -x: see every failure
cat gpurun_out/pytest_gpu.log
# e2e with and without the host-side 10 -> 8 bit reduction
( timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_4k10.json
( timeout 900 python bench.py --steps 10 --warmup 3 --host-narrow --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_4k10_host_narrow.json
# the same with pageable (staged) planes, where the narrowing copy replaces a memcpy rather than a direct DMA
( G1S_NO_DIRECT_H2D=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_4k10_staged.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/bench_reference.json
for f in gpurun_out/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("host_narrow"),
          "roofline", d.get("roofline", {}).get("frac"), "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("unreadable:", e)
PY
done
