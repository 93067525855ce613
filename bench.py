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
    ap.add_argument("--frames", type=int, default=60, help="frame pairs per step per GPU (3 engine batches at 4K)")
    ap.add_argument("--batch", type=int, default=0, help="frame pairs per kernel launch (engine batch)")
    ap.add_argument("--no-e2e", action="store_true")
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


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from grav1synth_b200 import abi
    from grav1synth_b200 import diff as D
    from grav1synth_b200.sharded import ShardedDiff
    from grav1synth_b200.synth import frame_pair_bytes, make_pair, to_numpy

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
    F = args.frames
    pair_bytes = frame_pair_bytes(W, H, 1, 1, bd, bd)

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

        def step():
            # one pass over the batch of F frame pairs; the engine pipelines launches and folds results
            # asynchronously, the drain happens once at the end of the timed region (barrier())
            for (sp, ss), (dp, ds) in dev_args:
                eng.diff_frame_device(sp, ss, dp, ds)
    else:
        sd = ShardedDiff(24, 1, bd, bd, W, H, 1, 1, frames_per_rank=F, device=local_rank, batch_frames=args.batch)
        eng = sd.producer

        def step():
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
    # timed region: CUDA events on the engine's own kernel stream (torch events only see torch's stream),
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
    total_frames = world * F * args.steps
    value = total_frames / elapsed

    # dominant work: residual + Gram accumulation (two back-to-back launches per batch: the streaming
    # residual kernel and the TMA-fed tensor-core Gram kernel), CUDA events on the engine's stream
    gram_ms = (c1["gram_ms"] - c0["gram_ms"]) / max(1.0, c1["gram_launches"] - c0["gram_launches"])
    res_ms = (c1["residual_ms"] - c0["residual_ms"]) / max(1.0, c1["gram_launches"] - c0["gram_launches"])
    flat_ms = (c1["flat_ms"] - c0["flat_ms"]) / max(1.0, c1["flat_launches"] - c0["flat_launches"])
    frames_per_launch = (c1["frames_done"] - c0["frames_done"]) / max(1.0, c1["gram_launches"] - c0["gram_launches"])
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    path_ms = gram_ms + res_ms
    achieved = frames_per_launch * pair_bytes / (path_ms * 1e-3) / 1e9 if path_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(tpath) and args.workload == "4k10" and abs(frames_per_launch - 20.0) < 1e-9:
        # ncu --set full capture of one residual + one gram launch of exactly this configuration (20 frame pairs)
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": wl["label"], "frames_per_step_per_gpu": F, "frames_per_launch": frames_per_launch, "bytes_per_frame_pair": pair_bytes,
                   "l2": f"inputs larger than L2 ({F * pair_bytes / 1e6:.0f} MB per step per GPU)",
                   "parallelism": f"frame-sharded x{world}, NCCL all-gather of per-frame model digests ({D.digest_bytes()} B/frame)",
                   "device_ms_flat_kernel": flat_ms, "device_ms_residual_kernel": res_ms,
                   "device_ms_gram_kernel": gram_ms},
        "gpu_launches": int(c1["kernels_launched"] - c0["kernels_launched"]),
        "roofline": {"bound": "hbm", "kernel": "residual_kernel + gram_imma_kernel (residual + autocorrelation; "
                                                  "algorithmic bytes over the SUM of both launch durations)",
                     "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic},
    }
    if clk is not None:
        line["clocks"] = clk
    if parity is not None:
        line.update(parity)
    table = eng.finish() if world == 1 else sd.finish()
    if rank == 0:
        line["config"]["segments"] = len(table)

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
        cs0 = gs.counters()
        gs.mark(0)
        ts0 = time.perf_counter()
        for _ in range(args.strict_steps):
            strict_step()
        gs.flush()
        gs.mark(1)
        s_ms = max(gs.marks_elapsed_ms(), (time.perf_counter() - ts0) * 1e3)
        cs1 = gs.counters()
        strict_table = gs.finish()
        line["value_strict"] = F * args.strict_steps / (s_ms * 1e-3)
        line["strict"] = {"unit": "frames/s", "gram_order": "REF_ORDER (per-term f64 chains, reference pixel order)",
                          "frames": F * args.strict_steps, "frames_per_launch": sb,
                          "device_ms_strict_kernel_per_frame": (cs1["strict_ms"] - cs0["strict_ms"]) / (F * args.strict_steps),
                          "segments": len(strict_table), "bound": "fp64 pipe / chain latency (5 f64 ops per term)"}

    # ---- e2e: host buffers through the C ABI (push_frame), copies inside the timed region
    if not args.no_e2e:
        nh = min(F, 8)
        host = []
        for s, d in frames[:nh]:
            hs = [t.cpu().pin_memory() for t in s]
            hd = [t.cpu().pin_memory() for t in d]
            host.append((hs, hd))
        np_frames = [([t.numpy().view(np.uint16) if bd > 8 else t.numpy() for t in hs],
                      [t.numpy().view(np.uint16) if bd > 8 else t.numpy() for t in hd]) for hs, hd in host]
        if world == 1:
            g = D.DiffGenerator(24, 1, bd, bd, W, H, device=local_rank, batch_frames=args.batch,
                                host_narrow=args.host_narrow)
            sd2 = None

            def e2e_step():
                for k in range(F):
                    s, d = np_frames[k % nh]
                    g.diff_frame(s, d)
        else:
            sd2 = ShardedDiff(24, 1, bd, bd, W, H, 1, 1, frames_per_rank=F, device=local_rank, batch_frames=args.batch)
            g = sd2.producer

            def e2e_step():
                for k in range(F):
                    s, d = np_frames[k % nh]
                    sd2.push_local(s, d)
                sd2.exchange()

        def e2e_barrier():
            if world == 1:
                g.flush()
            else:
                sd2.exchange(final=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()

        e2e_step()
        e2e_barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        e2e_barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        line["e2e"] = {"value": world * F * n_e2e / float(dt.item()), "unit": "frames/s",
                       "h2d_bytes_per_step": F * (pair_bytes // 2 if (args.host_narrow and bd > 8 and world == 1) else pair_bytes),
                       "d2h_bytes_per_step": F * g.record_bytes, "records_bytes_per_frame": g.record_bytes,
                       "host_narrow": bool(args.host_narrow and bd > 8 and world == 1),
                       "steps": n_e2e, "api": "g1s_diff_push_frame (C ABI) with host planes + g1s_diff_flush"}
        g.close()

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
