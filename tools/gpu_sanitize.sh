#!/bin/bash
# compute-sanitizer passes over a small end-to-end run (all kernels, TMA path and generic fallback)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from grav1synth_b200 import diff as D
from grav1synth_b200.synth import SynthSpec, make_pair_numpy
for spec in (SynthSpec(203, 117, 8, textured=0.3, sigma0=1.2, seed=3), SynthSpec(320, 176, 10, textured=0.3, seed=7)):
    g = D.DiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, spec.width, spec.height, batch_frames=2)
    for k in range(3):
        s, d = make_pair_numpy(spec, k)
        if k == 1:
            s[0][40, 50] = (1 << spec.bit_depth) - 1   # forces the int8-overflow fallback for one unit
            d[0][40, 50] = 0
        g.diff_frame(s, d)
    print(len(g.finish()), g.counters())
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|Error|error|hazard|Invalid|Uninit" | head -8
done > gpurun_out/sanitizer.log 2>&1
cat gpurun_out/sanitizer.log
