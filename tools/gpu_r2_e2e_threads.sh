#!/bin/bash
# Round 2: e2e host_narrow: host threads x streaming stores.
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo; grep -c avx512bw /proc/cpuinfo
for th in 8 12 16; do
for ns in 0 1; do
  ( [ $ns = 1 ] && export G1S_NO_STREAM_STORES=1; G1S_HOST_THREADS=$th timeout 600 python bench.py --steps 3 --warmup 3 --repeat 2 --no-cpu-baseline --no-strict --no-stats 2>&1 | tail -1 ) > gpurun_out/v_${th}_$ns.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v_${th}_$ns.log").read().strip().splitlines()[-1])
    print("threads $th no_stream=$ns:", "pinned", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]), "narrow", round(d["e2e_host_narrow"]["value"]), "value", round(d["value"]))
except Exception as e: print("failed", e)
PY
done
done
