#!/usr/bin/env python
"""Digest of one .ncu-rep: headline metrics, opcode mix, heaviest SASS regions, most-sampled instructions.

    python tools/ncu_digest.py gpurun_out/prof_gram_imma.ncu-rep [--regions]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
for i, h in enumerate(hdr):
    if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.2):
        print(f"{h:90s} {units[i]:12s} {vals[i]}")

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
ia, isrc, iex, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), r[isrc], int(r[iex]), int(r[ism])))
    except (ValueError, IndexError):
        pass
base = data[0][0]
tot = sum(d[2] for d in data)
ts = sum(d[3] for d in data)
print(f"instructions {tot}  samples {ts}")
op = collections.Counter()
for a, s, e, sm in data:
    o = s.split()[1] if s.startswith("@") else s.split()[0]
    op[o.split(".")[0]] += e
print("opcode mix: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in op.most_common(14)))
if "--regions" in sys.argv:
    cur, regions = None, []
    for a, s, e, sm in data:
        if cur and cur["e"] == e:
            cur["n"] += 1; cur["end"] = a; cur["sm"] += sm; cur["imma"] += "IMMA" in s
        else:
            cur = dict(start=a, end=a, e=e, n=1, sm=sm, imma=int("IMMA" in s)); regions.append(cur)
    for r in regions:
        w = r["e"] * r["n"]
        if w / tot > 0.005 or r["sm"] / ts > 0.01:
            print(f"{r['start'] - base:6x}-{r['end'] - base:6x} n={r['n']:4d} exec={r['e']:9d} weight={100 * w / tot:5.1f}% imma={r['imma']:3d} samples={100 * r['sm'] / ts:5.1f}%")
print("most sampled:")
for a, s, e, sm in sorted(data, key=lambda d: -d[3])[:12]:
    print(f"  {a - base:6x} {100 * sm / ts:5.1f}%  exec={e:9d}  {s[:80]}")
