#!/bin/bash
# A/B of libg1s variants (tools only): copies each libg1s_<tag>.so over libg1s.so and runs the short bench.
mkdir -p gpurun_out
cp grav1synth_b200/libg1s.so /tmp/libg1s_orig.so
for f in grav1synth_b200/libg1s_*.so; do
  cp $f grav1synth_b200/libg1s.so
  timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.read()); c = l['config']
print('$f', 'value', round(l['value']), 'flat', round(c['device_ms_flat_kernel'], 4), 'res', round(c['device_ms_residual_kernel'], 4), 'gram', round(c['device_ms_gram_kernel'], 4))"
done | tee gpurun_out/ab.log
cp /tmp/libg1s_orig.so grav1synth_b200/libg1s.so
