#!/bin/bash
# Measurement aid: the Gram kernel with its k-loops or its TMA traffic removed (results are wrong by design).
mkdir -p gpurun_out
for m in 0 1 2; do
  G1S_GRAM_MODE=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.read()); c = l['config']
print('mode $m', 'value', round(l['value']), 'flat', round(c['device_ms_flat_kernel'], 4), 'res', round(c['device_ms_residual_kernel'], 4), 'gram', round(c['device_ms_gram_kernel'], 4))"
done | tee gpurun_out/modes.log
