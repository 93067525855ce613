#!/usr/bin/env python
"""Benchmark of the `diff` hot path (BASELINE.json metric: diff frames/s at 4K 10-bit).

    python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W   # the CPU restatement of the reference

A step is one pass of the hot path over one batch of `--frames` synthetic 3840x2160 10-bit 4:2:0
source/denoised pairs (BASELINE configs[2]).  `value` is frames/s with the frames already resident
in HBM; `e2e` is the same metric through the C-ABI push_frame call with HOST buffers (host->device
copies and the record read-back inside the timed region).  Under torchrun every rank processes
its own `--frames` frames per step (weak scaling); the per-frame records are all-gathered over
NCCL and folded into the model on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "4k10": dict(width=3840, height=2160, bit_depth=10, label="diff 3840x2160 10-bit YUV420, ar_coeff_lag=3 + chroma"),
    "1080p8": dict(width=1920, height=1080, bit_depth=8, label="diff 1920x1080 8-bit YUV420"),
    "8k10": dict(width=7680, height=4320, bit_depth=10, label="diff 7680x4320 10-bit YUV420 + chroma"),
}
METRIC = "diff_frames_per_sec_4k_10bit"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="4k10", choices=list(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0,
                    help="resident frame pairs per GPU, one pass = one sharded super-batch (0: two engine batches, 82 at 4K 10-bit)")
    ap.add_argument("--batch", type=int, default=0, help="frame pairs per kernel launch (engine batch)")
    ap.add_argument("--repeat", type=int, default=40,
                    help="passes over the resident frames per step (a step is repeat x frames frame pairs per GPU, so that "
                         "the default 20..40 steps time seconds, not milliseconds)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-variants", action="store_true", help="skip the pageable / host_narrow e2e legs")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-stats", action="store_true", help="skip the workload statistics and the ~10 %-flat input pass")
    ap.add_argument("--host-narrow", action="store_true",
                    help="e2e leg: reduce >8-bit host samples to 8 bits while staging (cfg.host_narrow), half the H2D bytes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strict", action="store_true", help="skip the strict-mode (reference-order Gram) pass")
    ap.add_argument("--strict-steps", type=int, default=2)
    ap.add_argument("--strict-batch", type=int, default=0, help="frames per launch in the strict pass (0: min(frames, 60))")
    ap.add_argument("--cpu-frames", type=int, default=2, help="frames of the CPU-baseline sample")
    return ap.parse_args()


def synth_spec(wl):
    from grav1synth_b200.synth import SynthSpec
    return SynthSpec(wl["width"], wl["height"], wl["bit_depth"], textured=0.1, sigma0=1.0, sigma1=1.5, seed=20260917)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_cpu_sample(wl, nframes, cores=1):
    """The CPU restatement (oracle, reference accumulation order, libm exp) timed on nframes pairs."""
    from grav1synth_b200.synth import make_pair_numpy
    from oracle import oracle as O
    spec = synth_spec(wl)
    frames = [make_pair_numpy(spec, k) for k in range(nframes)]
    o = O.OracleDiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, O.GRAM_REF_ORDER, O.EXP_LIBM)
    t0 = time.perf_counter()
    for s, d in frames:
        o.diff_frame(s, d)
    o.finish()
    dt = time.perf_counter() - t0
    return nframes / dt, dt


def run_libaom_sample(wl, nframes=1):
    """Corroboration of the CPU baseline: the same frames (reduced to 8 bits, outside the timed region) through
    libaom 3.13.1's own compiled noise_model.c (oracle/aom_pin.py), the code av1-grain's `diff` ports.  Returns
    None where the bundled libaom is absent."""
    try:
        from grav1synth_b200.synth import make_pair_numpy
        from oracle import aom_pin as P
        ok, where = P.available()
        if not ok:
            return None
        spec = synth_spec(wl)
        frames = [make_pair_numpy(spec, k) for k in range(nframes)]
        frames = [([P.to_u8(p, spec.bit_depth) for p in s], [P.to_u8(p, spec.bit_depth) for p in d])
                  for s, d in frames]
        a = P.AomNoiseModel()
        t0 = time.perf_counter()
        for s, d in frames:
            a.update(s, d)
        a.finish()
        dt = time.perf_counter() - t0
        a.close()
        return {"value": nframes / dt, "unit": "frames/s", "cores": 1,
                "sample": f"{nframes} frame pair(s), {dt:.1f} s, {os.path.basename(where)}: aom_flat_block_finder_run + "
                          "aom_noise_model_update + get_grain_parameters on the 8-bit-reduced planes"}
    except Exception as e:  # the corroboration must never break the bench line
        return {"unavailable": repr(e)[:200]}


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference binary cannot be
    built here (Rust; hot path in the un-vendored crate av1-grain 0.4.2), so this times the C oracle in
    reference operation order.  The reference's diff loop is single-threaded (src/main.rs:432-521), and
    its model update is sequential across frames, so it can use one host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    from grav1synth_b200.synth import make_pair_numpy
    from oracle import oracle as O
    spec = synth_spec(wl)
    per_step = 1  # bounded sample: one frame pair of the workload per step
    frames = [make_pair_numpy(spec, k) for k in range(2)]
    o = O.OracleDiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, O.GRAM_REF_ORDER, O.EXP_LIBM)
    for i in range(args.warmup):
        o.diff_frame(*frames[i % 2])
    t0 = time.perf_counter()
    for i in range(args.steps):
        o.diff_frame(*frames[i % 2])
    dt = time.perf_counter() - t0
    o.finish()
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "frames_per_step": per_step,
                   "note": "C restatement of av1-grain 0.4.2 diff in reference operation order, not the Rust binary"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                         "sample": f"{per_step} frame pair per step x {args.steps} steps of the same workload"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    up = run_libaom_sample(wl, 1)
    if up is not None:
        line["cpu_baseline"]["upstream_libaom"] = up
    # Courtesy number, NOT what the reference does: its diff loop is single-threaded and its model update is sequential
    # across frames.  One independent frame per host thread (ctypes releases the GIL) bounds what a frame-parallel
    # rewrite of the reference could reach on this box.
    try:
        from concurrent.futures import ThreadPoolExecutor
        workers = max(1, min(os.cpu_count() or 1, 256))

        def one(i):
            h = O.OracleDiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, O.GRAM_REF_ORDER, O.EXP_LIBM)
            h.diff_frame(*frames[i % 2])
            h.close()

        t1 = time.perf_counter()
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(one, range(workers)))
        dtp = time.perf_counter() - t1
        line["cpu_baseline"]["frame_parallel_courtesy"] = {
            "value": workers / dtp, "unit": "frames/s", "cores": workers,
            "sample": f"{workers} independent frame pairs, one per host thread, {dtp:.1f} s; per-frame work only "
                      "(no sequential model combine) -- an upper bound for a hypothetical multi-threaded reference"}
    except Exception as e:
        line["cpu_baseline"]["frame_parallel_courtesy"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line), flush=True)


def kernel_pass(D, abi, dev_args, dims, local_rank, batch, passes):
    """Per-kernel device times with ONE kernel stream (nothing overlaps, so every launch's CUDA-event duration is its
    own): a few passes over the resident frames on a fresh handle.  Returns ms per frame of each kernel."""
    W, H, bd = dims
    old = os.environ.get("G1S_STREAMS")
    os.environ["G1S_STREAMS"] = "1"
    try:
        g = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=batch)
    finally:
        if old is None:
            del os.environ["G1S_STREAMS"]
        else:
            os.environ["G1S_STREAMS"] = old

    def one():
        for (sp, ss), (dp, ds) in dev_args:
            g.diff_frame_device(sp, ss, dp, ds)

    one()
    g.flush()
    c0 = g.counters()
    g.mark(0)
    for _ in range(passes):
        one()
    g.flush()
    g.mark(1)
    ms = g.marks_elapsed_ms()
    c1 = g.counters()
    g.close()
    n = max(1.0, c1["frames_done"] - c0["frames_done"])
    return {"flat_features": (c1["flat_ms"] - c0["flat_ms"]) / n, "residual": (c1["residual_ms"] - c0["residual_ms"]) / n,
            "gram_plan+gram_imma": (c1["gram_ms"] - c0["gram_ms"]) / n, "whole_step_one_stream": ms / n,
            "frames": int(n), "frames_per_launch": n / max(1.0, c1["gram_launches"] - c0["gram_launches"])}


def workload_stats(D, frames_np, dims):
    """Flat-block fraction and observation counts of the synthetic input (what the Gram work is proportional to)."""
    W, H, bd = dims
    g = D.DiffGenerator(24, 1, bd, bd, W, H)
    recs = []
    g.set_record_tap(lambda i, r: recs.append(r))
    for s, d in frames_np:
        g.diff_frame(s, d)
    g.finish()
    rl = D.RecordLayout(g.num_blocks)
    u = [rl.unpack(r) for r in recs]
    g.close()
    nobs = [float(sum(int(x["nobs"][c]) for x in u)) / len(u) for c in range(3)]
    flat = float(sum(x["num_flat"] for x in u)) / len(u) / g.num_blocks
    macs = nobs[0] * 324 + (nobs[1] + nobs[2]) * 350  # symmetric half of the 24/25-tap outer product + the b vector
    return {"flat_block_fraction": flat, "observations_per_frame": nobs, "algorithmic_int_mac_per_frame": macs}


def main():
    args = parse_args()
    # A run takes a minute or two.  Should it ever stall (a device fault, a lost rank), say where and leave instead of
    # holding the GPUs until somebody's timeout: every Python thread's stack goes to stderr, then the process exits.
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("G1S_BENCH_WATCHDOG_S", "1200")), exit=True)
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from grav1synth_b200 import abi
    from grav1synth_b200 import diff as D
    from grav1synth_b200.sharded import ShardedDiff
    from grav1synth_b200.synth import SynthSpec, frame_pair_bytes, make_pair, to_numpy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this engine has no CPU fallback")
    if world > 1 and "G1S_HOST_THREADS" not in os.environ:
        # the per-frame half of the host model runs on every rank: share the host cores between the ranks
        os.environ["G1S_HOST_THREADS"] = str(max(2, min(8, (os.cpu_count() or 16) // world)))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = WORKLOADS[args.workload]
    spec = synth_spec(wl)
    W, H, bd = spec.width, spec.height, spec.bit_depth
    F, R = args.frames, args.repeat
    pair_bytes = frame_pair_bytes(W, H, 1, 1, bd, bd)
    if F <= 0:
        # whole engine batches per pass: a sharded super-batch (one pass per rank) then ends on a launch boundary
        probe = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=args.batch)
        F = 2 * probe.batch_frames
        probe.close()

    # ---- multi-GPU parity, before anything is timed: a two-scene stream through the NCCL-sharded path must give the
    # single-GPU table (segment cut and short tail super-batch included); a mismatch fails the run
    parity = None
    if world > 1:
        from grav1synth_b200.sharded import parity_check
        pframes, pseg = parity_check(local_rank)
        parity = {"parity_checked": True, "parity_frames": pframes, "parity_segments": pseg}

    # ---- synthetic frames, generated directly in HBM (distinct per rank), outside the timed region
    dev = f"cuda:{local_rank}"
    frames = [make_pair(spec, rank * F + k, dev) for k in range(F)]
    torch.cuda.synchronize()

    def ptrs(planes):
        return [p.data_ptr() for p in planes], [p.stride(0) * p.element_size() for p in planes]

    dev_args = [(ptrs(s), ptrs(d)) for s, d in frames]

    if world == 1:
        # single GPU: the drop-in handle itself (kernels + host model), frames pushed as device pointers
        sd = None
        eng = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=args.batch)

        prepared = [(eng.device_frame(sp, ss), eng.device_frame(dp, ds)) for (sp, ss), (dp, ds) in dev_args]

        def step():
            # R passes over the F resident frame pairs (3.0 GB at 4K: every pass streams from HBM); the engine pipelines
            # launches and folds results asynchronously, the drain happens once at the end of the timed region
            push = eng.diff_frames_prepared
            for _ in range(R):
                for sf, df in prepared:
                    push(sf, df)
    else:
        sd = ShardedDiff(24, 1, bd, bd, W, H, 1, 1, frames_per_rank=F, device=local_rank, batch_frames=args.batch)
        eng = sd.producer

        def step():
            for _ in range(R):
                for sp, dp in dev_args:
                    sd.push_local(sp, dp, device_resident=True)
                sd.exchange()

    def barrier():
        if world == 1:
            eng.flush()  # every pushed frame processed by the device AND folded into the host model
        else:
            sd.exchange(final=True)  # drain kernels, gather the outstanding digests, fold them on rank 0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    c0 = eng.counters()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    # timed region: CUDA events on the engine's own kernel streams (torch events only see torch's stream),
    # cross-checked against the host clock; the larger of the two is reported
    eng.mark(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    eng.mark(1)
    ev_ms = eng.marks_elapsed_ms()
    t1 = time.perf_counter()
    clk = clocks.stop() if clocks else None
    c1 = eng.counters()
    elapsed = torch.tensor([max(t1 - t0, ev_ms * 1e-3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed = float(elapsed.item())
    total_frames = world * F * R * args.steps
    value = total_frames / elapsed
    frames_per_launch = (c1["frames_done"] - c0["frames_done"]) / max(1.0, c1["gram_launches"] - c0["gram_launches"])

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, peak_src = float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    # the path's roofline: algorithmic bytes (one read of every sample of both frames) over the WHOLE step
    # -- flat-block finder, threshold select, residual, plan and Gram launches, as they overlap on the kernel streams --
    # per GPU, measured over the timed region itself
    step_gbs = (total_frames / world) * pair_bytes / elapsed / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath) and args.workload == "4k10":
        # ncu --set full captures of every kernel of one 20-frame batch of exactly this workload
        tj = json.load(open(tpath))
        traffic = tj.get("dram_bytes_per_frame", 0) * frames_per_launch or None

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": wl["label"], "frames_per_step_per_gpu": F * R, "resident_frames": F, "passes_per_step": R,
                   "frames_per_launch": frames_per_launch, "bytes_per_frame_pair": pair_bytes,
                   "l2": f"inputs larger than L2 ({F * pair_bytes / 1e6:.0f} MB resident per GPU, streamed once per pass)",
                   "timed_region_s": elapsed,
                   "parallelism": f"frame-sharded x{world}, NCCL all-gather of per-frame model digests ({D.digest_bytes()} B/frame)",
                   "kernel_streams": int(os.environ.get("G1S_STREAMS", "3")),
                   "model_placement": ("device (latest_kernel: per-frame model half on the GPU, digests cross PCIe)"
                                       if eng.model_on_device else "host (per-frame model half on host threads, records cross PCIe)")},
        "gpu_launches": int(c1["kernels_launched"] - c0["kernels_launched"]),
        "roofline": {"bound": "hbm",
                     "kernel": "whole step: flat_features + flat_select + residual + gram_plan + gram_imma (+ gram_generic "
                               "for flagged blocks), overlapped on the engine's kernel streams; algorithmic bytes over "
                               "the timed region",
                     "achieved": step_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": step_gbs / peak,
                     "traffic": traffic},
    }
    if clk is not None:
        line["clocks"] = clk
    if parity is not None:
        line.update(parity)
    table = eng.finish() if world == 1 else sd.finish()
    if rank == 0:
        line["config"]["segments"] = len(table)

    if world == 1:
        # ---- per-kernel times, one stream (no overlap), and the figures the survey asked for beside the roofline
        kp = kernel_pass(D, abi, dev_args, (W, H, bd), local_rank, args.batch, 2)
        rg = kp["residual"] + kp["gram_plan+gram_imma"]
        line["kernels"] = {"ms_per_frame_one_stream": kp,
                           "residual+gram": {"achieved": pair_bytes / (rg * 1e-3) / 1e9, "unit": "GB/s",
                                             "frac": pair_bytes / (rg * 1e-3) / 1e9 / peak},
                           "whole_step_one_stream": {"achieved": pair_bytes / (kp["whole_step_one_stream"] * 1e-3) / 1e9,
                                                     "frac": pair_bytes / (kp["whole_step_one_stream"] * 1e-3) / 1e9 / peak}}
        if rank == 0 and not args.no_stats:
            ws = workload_stats(D, [(to_numpy(s), to_numpy(d)) for s, d in frames[:2]], (W, H, bd))
            ws["int_mac_per_s"] = ws["algorithmic_int_mac_per_frame"] * value
            line["workload"] = ws
            # the same path on an input where only ~10 % of the blocks are flat (the Gram work shrinks with it)
            sp2 = SynthSpec(W, H, bd, textured=0.93, sigma0=spec.sigma0, sigma1=spec.sigma1, seed=spec.seed)
            fr2 = [make_pair(sp2, k, dev) for k in range(min(F, 20))]
            da2 = [(ptrs(s), ptrs(d)) for s, d in fr2]
            g2 = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=args.batch)
            for _ in range(3):
                for (sp, ss), (dp, ds) in da2:
                    g2.diff_frame_device(sp, ss, dp, ds)
            g2.flush()
            g2.mark(0)
            n2 = 0
            for _ in range(100):  # 2000 frames: dozens of batches, so that filling and draining the pipeline do not count
                for (sp, ss), (dp, ds) in da2:
                    g2.diff_frame_device(sp, ss, dp, ds)
                    n2 += 1
            g2.flush()
            g2.mark(1)
            v2 = n2 / (g2.marks_elapsed_ms() * 1e-3)
            g2.close()
            ws2 = workload_stats(D, [(to_numpy(s), to_numpy(d)) for s, d in fr2[:1]], (W, H, bd))
            line["sparse_input"] = {"value": v2, "unit": "frames/s", "flat_block_fraction": ws2["flat_block_fraction"],
                                    "roofline_frac": v2 * pair_bytes / 1e9 / peak}
            del fr2, da2

    # ---- strict mode (gram_order = REF_ORDER): the reference's per-term f64 accumulation reproduced bit for bit on
    # the device, so every table integer equals the reference's; same device-resident frames, its own timed pass
    if world == 1 and not args.no_strict:
        sb = args.strict_batch or min(F, 60)
        gs = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=sb, gram_order=abi.GRAM_REF_ORDER)

        def strict_step():
            for (sp, ss), (dp, ds) in dev_args:
                gs.diff_frame_device(sp, ss, dp, ds)

        strict_step()
        gs.flush()
        gs.mark(0)
        ts0 = time.perf_counter()
        for _ in range(args.strict_steps):
            strict_step()
        gs.flush()
        gs.mark(1)
        s_ms = max(gs.marks_elapsed_ms(), (time.perf_counter() - ts0) * 1e3)
        strict_table = gs.finish()
        line["value_strict"] = F * args.strict_steps / (s_ms * 1e-3)
        line["strict"] = {"unit": "frames/s", "gram_order": "REF_ORDER (per-term f64 chains, reference pixel order)",
                          "frames": F * args.strict_steps, "frames_per_launch": sb, "segments": len(strict_table),
                          "bound": "fp64 pipe / chain latency (5 f64 ops per term, 3.6 G terms per frame)"}

    # ---- e2e: host buffers through the C ABI (push_frame), copies inside the timed region
    if not args.no_e2e:
        nh = min(F, 8)
        host = []
        for s, d in frames[:nh]:
            hs = [t.cpu().pin_memory() for t in s]
            hd = [t.cpu().pin_memory() for t in d]
            host.append((hs, hd))

        def as_np(ts):
            return [t.numpy().view(np.uint16) if bd > 8 else t.numpy() for t in ts]

        pinned_frames = [(as_np(hs), as_np(hd)) for hs, hd in host]

        def run_e2e(np_frames, narrow, steps):
            if world == 1:
                g = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=args.batch, host_narrow=narrow)
                sd2 = None

                # the g1s_frame structs over the caller's buffers are built once (a caller that reuses its buffers does
                # the same); the timed loop is the C ABI call itself
                prep = [(g.host_frame(s), g.host_frame(d)) for s, d in np_frames]

                def e2e_step():
                    push = g.diff_frames_prepared_host
                    for k in range(F):
                        (sf, _), (df, _) = prep[k % nh]
                        push(sf, df)

                def e2e_barrier():
                    g.flush()
                    torch.cuda.synchronize()
            else:
                sd2 = ShardedDiff(24, 1, bd, bd, W, H, 1, 1, frames_per_rank=F, device=local_rank, batch_frames=args.batch)
                g = sd2.producer

                def e2e_step():
                    for k in range(F):
                        s, d = np_frames[k % nh]
                        sd2.push_local(s, d)
                    sd2.exchange()

                def e2e_barrier():
                    sd2.exchange(final=True)
                    torch.cuda.synchronize()
                    dist.barrier()

            e2e_step()
            e2e_barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step()
            e2e_barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            rb = g.record_bytes
            g.close()
            return world * F * steps / float(dt.item()), rb

        n_e2e = max(1, min(args.steps, args.e2e_steps))
        # Three ways a caller's host frames reach the device, all through g1s_diff_push_frame:
        #   pinned_direct  page-locked planes, one 2-D DMA per plane (PCIe-bound: 49.8 MB per 4K 10-bit frame pair)
        #   pageable       ordinary planes (the Rust caller's v_frame), staged through the engine's pinned ring
        #   host_narrow    ordinary planes, samples reduced to 8 bits while staged (cfg.host_narrow; the kernels' first
        #                  step anyway, identical results): half the bytes on PCIe -- what `python -m grav1synth_b200 diff`
        #                  does for sources deeper than 8 bits, and the headline when it applies
        api = "g1s_diff_push_frame (C ABI) with host planes + g1s_diff_flush"
        v_pin, rb = run_e2e(pinned_frames, False, n_e2e)
        e_pin = {"value": v_pin, "unit": "frames/s", "h2d_bytes_per_step": F * pair_bytes, "d2h_bytes_per_step": F * rb,
                 "records_bytes_per_frame": rb, "frames_per_step": F, "steps": n_e2e,
                 "host_memory": "pinned (direct 2-D DMA per plane)", "api": api}
        line["e2e"] = e_pin
        if world == 1 and not args.no_e2e_variants:
            pageable = [([np.array(p) for p in s], [np.array(p) for p in d]) for s, d in pinned_frames]
            v_pg, _ = run_e2e(pageable, False, n_e2e)
            line["e2e_pageable"] = {"value": v_pg, "unit": "frames/s", "h2d_bytes_per_step": F * pair_bytes,
                                    "host_memory": "pageable (staged through the engine's pinned ring by its host threads)"}
            if bd > 8:
                v_nr, _ = run_e2e(pageable, True, n_e2e)
                e_nr = {"value": v_nr, "unit": "frames/s", "h2d_bytes_per_step": F * pair_bytes // 2,
                        "d2h_bytes_per_step": F * rb, "records_bytes_per_frame": rb, "frames_per_step": F, "steps": n_e2e,
                        "host_memory": "pageable planes, reduced to 8 bits into the engine's pinned ring by its host threads "
                                       "(cfg.host_narrow), DMA from there", "api": api}
                try:
                    e_nr["narrow_isa"] = {0: "compiler loop", 1: "avx2 + streaming stores", 2: "avx512bw + streaming stores"}[
                        int(D.lib().g1s_narrow_isa())]
                    e_nr["host_threads"] = min(16, os.cpu_count() or 1)
                except Exception:
                    pass
                line["e2e_host_narrow"] = e_nr
                line["e2e_pinned_direct"] = e_pin
                if v_nr > v_pin:
                    line["e2e"] = e_nr

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, dt = run_cpu_sample(wl, args.cpu_frames)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": 1, "kind": "port",
                                "sample": f"{args.cpu_frames} frame pairs of the same workload, {dt:.1f} s; "
                                          "C restatement of av1-grain 0.4.2 diff (reference op order), host has "
                                          f"{os.cpu_count()} cores, reference diff loop is single-threaded"}
        up = run_libaom_sample(wl, 1)
        if up is not None:
            line["cpu_baseline"]["upstream_libaom"] = up
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
