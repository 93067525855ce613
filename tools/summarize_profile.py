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

# ---- launch list (gpu__time_duration + DRAM bytes per launch; cold-cache and serialised: compare SHARES)
rows = [r for r in csv.reader(open(os.path.join(src_dir, "launches.csv"))) if len(r) > 5]
hdr, agg = None, {}
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").strip().split("::")[-1]
    metric = r[hdr.index("Metric Name")]
    agg.setdefault(name, {}).setdefault(metric, []).append(float(r[hdr.index("Metric Value")].replace(",", "")))
dur = {k: v.get("gpu__time_duration.sum", []) for k, v in agg.items()}
total = sum(sum(v) for v in dur.values())
frames = int(os.environ.get("G1S_PROFILE_FRAMES", "41"))
launches = max(len(v) for v in dur.values())
dram_total = 0.0
with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
            f"-k regex:'flat_|gram_|residual_' python bench.py ... (G1S_STREAMS=1, {frames} frames per launch)\n")
    f.write("# kernel, launches, avg us, share of the step, avg DRAM read MB, avg DRAM write MB\n")
    for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
        rd = agg[k].get("dram__bytes_read.sum", [0]); wr = agg[k].get("dram__bytes_write.sum", [0])
        dram_total += (sum(rd) + sum(wr)) / max(1, len(v)) * (len(v) / launches)
        f.write(f"{k:34s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {100 * sum(v) / total:6.1f}% "
                f"{sum(rd) / len(rd) / 1e6:10.1f} {sum(wr) / len(wr) / 1e6:10.1f}\n")
    f.write(f"# whole step: {total / launches / 1e3:.1f} us per {frames}-frame batch, DRAM traffic {dram_total / 1e6:.1f} MB "
            f"= {dram_total / frames / 1e6:.2f} MB per frame pair (algorithmic: 49.77 MB)\n")
json.dump({"dram_bytes_per_frame": dram_total / frames, "frames_per_launch": frames,
           "source": f"profiles/{tag}_launches.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum of every kernel of one batch)"},
          open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1)

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
for kern in ("gram_imma", "residual", "flat_features", "gram_plan", "latest", "strict"):
    rep = os.path.join(src_dir, f"prof_{kern}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, vals = rr[0], rr[1], rr[2]
    got = {}
    with open(os.path.join(out_dir, f"{tag}_{kern}_ncu.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kern} (one launch of the batch the launch list shows; strict: 20 frame pairs), key metrics\n")
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
# ---- the per-frame model kernel (device placement only): its own launch list, the phase clocks, the chain probe
ll = os.path.join(src_dir, "launches_latest.csv")
if os.path.exists(ll):
    vals = {}
    hdr = None
    for r in csv.reader(open(ll)):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
        elif hdr and len(r) > 5:
            vals.setdefault(r[hdr.index("Metric Name")], []).append(float(r[hdr.index("Metric Value")].replace(",", "")))
    with open(os.path.join(out_dir, f"{tag}_latest_launches.txt"), "w") as f:
        f.write(f"# G1S_DEVICE_MODEL=1: latest_kernel, one CTA per frame ({frames} frames per launch), ncu per-launch metrics\n")
        for k, v in vals.items():
            f.write(f"{k:34s} launches {len(v)}  avg {sum(v) / len(v):.1f}\n")
for name in ("latest_phases.log", "chain_probe.txt"):
    p = os.path.join(src_dir, name)
    if os.path.exists(p):
        with open(os.path.join(out_dir, f"{tag}_{name.replace('.log', '.txt')}"), "w") as f:
            if name == "latest_phases.log":
                f.write("# clock64() phase marks of CTA 0 of latest_kernel (build with -DG1S_LATEST_PROF; not the product build)\n")
            else:
                f.write("# tools/chain_probe.cu: one thread's order-dependent f64 sum from shared memory, clocks per element\n")
            f.write(open(p).read())
print("wrote summaries for", tag)
