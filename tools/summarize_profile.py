#!/usr/bin/env python
"""Turns the ncu captures in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_profile.py r01
"""
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
src_dir = os.path.join(root, "gpurun_out")

# ---- launch list (gpu__time_duration per launch; cold-cache and serialised: compare SHARES)
rows = [r for r in csv.reader(open(os.path.join(src_dir, "launches.csv"))) if len(r) > 5]
hdr, agg = None, {}
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").strip()
    agg.setdefault(name, []).append(float(r[hdr.index("Metric Value")]))
total = sum(sum(v) for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'flat_|gram_|residual_' python bench.py ...\n")
    f.write("# kernel, launches, avg us, total us, share of the step\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k:50s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / 1e3:10.1f} {100 * sum(v) / total:6.1f}%\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
traffic = {}
for kern in ("gram_imma", "residual", "flat_features"):
    rep = os.path.join(src_dir, f"prof_{kern}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, vals = rr[0], rr[1], rr[2]
    got = {}
    with open(os.path.join(out_dir, f"{tag}_{kern}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kern} (one launch, 20 frame pairs), key metrics\n")
        f.write(f"# kernel: {vals[h.index('Kernel Name')]}\n")
        for name, u, v in zip(h, units, vals):
            if name in WANT:
                f.write(f"{name:90s} {u:12s} {v}\n")
                got[name] = (u, v)

    def to_bytes(key):
        u, v = got[key]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic[kern] = {"dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                     "ncu_duration": " ".join(reversed(got["gpu__time_duration.sum"]))}
if "gram_imma" in traffic and "residual" in traffic:
    # the roofline's `traffic`: DRAM bytes of one residual launch + one gram launch (the two launches that do
    # the residual + autocorrelation work of a 20-frame batch)
    tot = sum(traffic[k]["dram_bytes_read"] + traffic[k]["dram_bytes_write"] for k in ("gram_imma", "residual"))
    json.dump({"dram_bytes_per_launch": tot, "per_kernel": traffic,
               "source": f"profiles/{tag}_gram_imma_ncu.txt + profiles/{tag}_residual_ncu.txt"},
              open(os.path.join(out_dir, "gram_traffic.json"), "w"), indent=1)
print("wrote summaries for", tag)
