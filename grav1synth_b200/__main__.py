"""`python -m grav1synth_b200 diff SOURCE DENOISED -o OUT [-y]` — the reference's `diff` sub-command
(/root/reference/src/main.rs:347-533, arguments :846-871) over the B200 engine, reading .y4m clips
instead of going through FFmpeg.  Same checks and messages as the reference's Diff arm; the per-pixel
work goes through the C ABI (g1s_diff_create / push_frame / finish) and the table through
g1s_write_grain_table.

`python -m grav1synth_b200 inspect INPUT -o OUT [-y] [--fps N/D]` -- the reference's `inspect` (src/main.rs:145-196):
AV1 OBU headers of an .ivf / .obu file -> grain table, CPU only (C++ parser behind g1s_inspect_*).
`apply IN.ivf -g TABLE -o OUT.ivf` / `remove IN.ivf -o OUT.ivf` -- src/main.rs:197-246, 309-346: film grain headers
rewritten by the C++ parser (g1s_rewrite_*), IVF in and out.  `generate IN.ivf --iso N [--chroma] -o OUT.ivf`
(src/main.rs:247-308) applies a photon-noise segment (g1s_generate_photon_noise).
"""
from __future__ import annotations

import argparse
import logging
import os
import sys

log = logging.getLogger("grav1synth")


def parse_devices(spec):
    """'0-3' / '0,2,5' / None -> list of CUDA ordinals (None: the single --device)."""
    if not spec:
        return None
    out = []
    for part in spec.split(","):
        if "-" in part:
            a, b = part.split("-", 1)
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    if not 1 <= len(out) <= 8:
        raise SystemExit("--devices takes 1 to 8 CUDA ordinals")
    return out


def main(argv=None) -> int:
    logging.basicConfig(level=os.environ.get("G1S_LOG", "INFO"), format=" %(levelname)-5s %(name)s > %(message)s")
    ap = argparse.ArgumentParser(prog="grav1synth_b200", description="Grain Synth analyzer: B200 `diff` path")
    sub = ap.add_subparsers(dest="command", required=True)
    d = sub.add_parser("diff", help="Compares a source video to a denoised video and generates a film grain table")
    d.add_argument("source", help="The untouched source file to inspect (.y4m).")
    d.add_argument("denoised", help="The denoised file to inspect (.y4m).")
    d.add_argument("-o", "--output", required=True, help="The path to the output film grain table.")
    d.add_argument("-y", "--overwrite", action="store_true", help="Overwrite the output file without prompting.")
    d.add_argument("-f", "--filters", default=None,
                   help='A semicolon-separated list of filters to apply to the source before running the diff, e.g. '
                        '"crop:top=42,left=64" or "resize:width=1920,height=1080,alg=lanczos".')
    d.add_argument("--device", type=int, default=0)
    d.add_argument("--devices", default=None,
                   help="Several GPUs behind one handle, e.g. 0-7 or 0,2,3: batches of frames are dealt round-robin "
                        "(g1s_diff_config.n_devices / device_ids); the table is the single-GPU table.")
    d.add_argument("--strict", action="store_true",
                   help="Accumulate the AR normal equations in the reference's order and rounding (gram_order = "
                        "REF_ORDER): every table integer equals the reference's, at ~1/50 of the fast mode's speed.")
    i = sub.add_parser("inspect", help="Outputs a film grain table corresponding to a given AV1 video, or reports "
                                       "if there is no film grain information.")
    i.add_argument("input", help="The AV1 file to inspect (.ivf or low-overhead .obu).")
    i.add_argument("-o", "--output", required=True, help="The path to the output film grain table.")
    i.add_argument("-y", "--overwrite", action="store_true", help="Overwrite the output file without prompting.")
    i.add_argument("--fps", default=None, help="Frame rate N/D (default: the IVF header's rate/scale).")
    a = sub.add_parser("apply", help="Applies film grain from a table file to a given AV1 video, and outputs it at a "
                                     "given `output` path. Overwrites any existing grain.")
    a.add_argument("input", help="The AV1 file to apply grain to (.ivf).")
    a.add_argument("-o", "--output", required=True, help="The path to write the grain-synthed AV1 file.")
    a.add_argument("-y", "--overwrite", action="store_true", help="Overwrite the output file without prompting.")
    a.add_argument("-g", "--grain", required=True, help="The path to the input film grain table.")
    r = sub.add_parser("remove", help="Removes all film grain from a given AV1 video, and outputs it at a given "
                                      "`output` path.")
    r.add_argument("input", help="The AV1 file to remove grain from (.ivf).")
    r.add_argument("-o", "--output", required=True, help="The path to write the non-grain-synthed AV1 file.")
    r.add_argument("-y", "--overwrite", action="store_true", help="Overwrite the output file without prompting.")
    g = sub.add_parser("generate", help="Generates photon-noise-based film grain for a given AV1 video at a given ISO "
                                        "strength, and outputs it at a given `output` path. Overwrites any existing grain.")
    g.add_argument("input", help="The AV1 file to apply grain to (.ivf).")
    g.add_argument("-o", "--output", required=True, help="The path to write the grain-synthed AV1 file.")
    g.add_argument("-y", "--overwrite", action="store_true", help="Overwrite the output file without prompting.")
    g.add_argument("--iso", type=int, required=True, help="The ISO strength of the generated grain (>= 1).")
    g.add_argument("--chroma", action="store_true", help="Whether to apply grain to the chroma planes as well.")
    args = ap.parse_args(argv)
    if args.command == "inspect":
        return inspect_main(args)
    if args.command in ("apply", "remove", "generate"):
        return rewrite_main(args)

    # src/main.rs:354-368
    if os.path.abspath(args.source) == os.path.abspath(args.output) or \
            os.path.abspath(args.denoised) == os.path.abspath(args.output):
        log.error("Input and output paths are the same. This is probably a typo, because this would overwrite "
                  "your input. Exiting.")
        return 0
    if os.path.abspath(args.source) == os.path.abspath(args.denoised):
        log.error("Source and denoised paths are the same. This is probably a typo, because this would always "
                  "compute an empty diff. Exiting.")
        return 0
    from .filters import FilterChain, FilterError
    try:  # src/main.rs:370-380
        chain = FilterChain(args.filters) if args.filters else None
    except FilterError as e:
        log.error("Invalid filter chain: %s", e)
        return 0
    # src/main.rs:382-393
    if os.path.exists(args.output) and not args.overwrite:
        if not sys.stdin.isatty() or input(f"File {args.output} exists. Overwrite? [y/n] ").strip().lower() != "y":
            log.warning("Not overwriting existing file. Exiting.")
            return 0

    from .diff import DiffGenerator, write_grain_table
    from .y4m import Y4MReader
    src, den = Y4MReader(args.source), Y4MReader(args.denoised)
    sd, dd = src.get_video_details(), den.get_video_details()
    for bd in (sd.bit_depth, dd.bit_depth):
        if not 8 <= bd <= 16:
            raise SystemExit("Bit depths not between 8-16 are not currently supported")  # src/main.rs:516
    # the engine is sized for the frames it will see: the denoised clip's size (the source is filtered to it)
    # y4m planes are ordinary host memory: samples wider than 8 bits are reduced while they are staged (the kernels'
    # first step anyway, util.rs::frame_into_u8) so that half the bytes cross PCIe; the filter chain needs the full samples
    narrow = max(sd.bit_depth, dd.bit_depth) > 8 and not (chain is not None and chain.filters)
    differ = DiffGenerator(sd.fps_num, sd.fps_den, sd.bit_depth, dd.bit_depth, dd.width, dd.height, sd.ss_x, sd.ss_y,
                           monochrome=sd.monochrome, device=args.device, devices=parse_devices(args.devices),
                           gram_order=1 if args.strict else 0, host_narrow=narrow)
    if chain is not None and chain.filters:
        # source only, src/main.rs:621-624: the chain runs on the device between the upload and the kernels
        try:
            differ.set_source_filters(chain.filters, sd.width, sd.height)
        except ValueError as e:
            log.error("Invalid filter chain: %s", e)
            return 1
    frames = 0
    while True:  # src/main.rs:432-521
        s, d_ = src.get_frame(), den.get_frame()
        if s is None and d_ is None:
            break
        if s is None or d_ is None:
            log.warning("Videos did not have equal frame counts. Resulting grain table may not be as expected.")
            break
        differ.diff_frame(s, d_)  # raises ValueError on a dimension mismatch, like `?` on diff_frame
        frames += 1
    write_grain_table(differ.finish(), args.output)
    log.info("Computed diff for %d frames", frames)
    log.info("Done, wrote output file to %s", args.output)
    return 0


def inspect_main(args) -> int:
    """src/main.rs:145-196."""
    if os.path.abspath(args.input) == os.path.abspath(args.output):
        log.error("Input and output paths are the same. This is probably a typo, because this would overwrite "
                  "your input. Exiting.")
        return 0
    if os.path.exists(args.output) and not args.overwrite:
        if not sys.stdin.isatty() or input(f"File {args.output} exists. Overwrite? [y/n] ").strip().lower() != "y":
            log.warning("Not overwriting existing file. Exiting.")
            return 0
    from .diff import write_grain_table
    from .inspect import UPDATE_GRAIN, BitstreamParser
    fps = (0, 0)
    if args.fps:
        n, _, d = args.fps.partition("/")
        fps = (int(n), int(d or 1))
    parser = BitstreamParser()
    fps = parser.push_file(args.input, fps)
    if fps[0] <= 0 or fps[1] <= 0:
        raise SystemExit("the container carries no frame rate: pass --fps N/D")
    headers = parser.get_grain_headers()
    if not any(h.kind == UPDATE_GRAIN for h in headers):
        log.info("No film grain headers found--this video does not use grain synthesis")
        return 0
    write_grain_table(parser.aggregate_grain_headers(*fps), args.output)
    log.info("Done, wrote grain table to %s", args.output)
    return 0


def rewrite_main(args) -> int:
    """`apply` (src/main.rs:197-246) and `remove` (:309-346): IVF in, IVF out."""
    if os.path.abspath(args.input) == os.path.abspath(args.output):
        log.error("Input and output paths are the same. This is probably a typo, because this would overwrite "
                  "your input. Exiting.")
        return 0
    if os.path.exists(args.output) and not args.overwrite:
        if not sys.stdin.isatty() or input(f"File {args.output} exists. Overwrite? [y/n] ").strip().lower() != "y":
            log.warning("Not overwriting existing file. Exiting.")
            return 0
    from .grain_table import parse_grain_table
    from .inspect import GrainRewriter, rewrite_ivf
    table = None
    if args.command == "apply":
        with open(args.grain) as f:
            table = parse_grain_table(f.read())
    with open(args.input, "rb") as f:
        data = f.read()
    if args.command == "generate":  # src/main.rs:247-308: one segment over the whole stream from the stream's parameters
        if args.iso < 1:
            raise SystemExit("--iso must be at least 1")
        from .inspect import TRANSFER_BT1886, TRANSFER_SMPTE2084, BitstreamParser, generate_photon_noise_params
        probe = BitstreamParser()
        probe.push_file(args.input)
        info = probe.stream_info()
        trc = TRANSFER_SMPTE2084 if info["transfer_characteristics"] == 16 else TRANSFER_BT1886
        table = [generate_photon_noise_params(0, 2 ** 64 - 1, args.iso, info["max_frame_width"],
                                              info["max_frame_height"], trc, args.chroma,
                                              full_range=info["color_range"] == 1)]  # src/main.rs:299
    out = rewrite_ivf(data, GrainRewriter(table))
    with open(args.output, "wb") as f:
        f.write(out)
    log.info("Done, wrote output file to %s", args.output)
    return 0


if __name__ == "__main__":
    sys.exit(main())
