"""Reader for the aom `filmgrn1` grain-table text (the format src/main.rs:631-696 writes and the reference's
`parse_grain_table`, used by its `apply` command at src/main.rs:223, reads back).  Host-side plumbing: lets the
tests and the CLI round-trip tables; the writer is g1s_format_grain_table in the C ABI."""
from __future__ import annotations

from typing import List

from .abi import CSegment, GrainTableSegment


def parse_grain_table(text: str) -> List[GrainTableSegment]:
    lines = text.splitlines()
    if not lines or lines[0].strip() != "filmgrn1":
        raise ValueError("not a filmgrn1 grain table")
    segs: List[GrainTableSegment] = []
    i = 1
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        e = lines[i].split()
        if e[0] != "E" or len(e) < 6:
            raise ValueError(f"line {i + 1}: expected an E line")
        s = CSegment()
        s.start_time, s.end_time, s.random_seed = int(e[1]), int(e[2]), int(e[4])
        body = {}
        for ln in lines[i + 1:i + 8]:
            t = ln.split()
            body[t[0]] = [int(x) for x in t[1:]]
        p = body["p"]
        (s.ar_coeff_lag, s.ar_coeff_shift, s.grain_scale_shift, s.scaling_shift, s.chroma_scaling_from_luma,
         s.overlap_flag, s.cb_mult, s.cb_luma_mult, s.cb_offset, s.cr_mult, s.cr_luma_mult, s.cr_offset) = p
        for key, dst, cnt in (("sY", s.scaling_points_y, "num_y_points"), ("sCb", s.scaling_points_cb, "num_cb_points"),
                              ("sCr", s.scaling_points_cr, "num_cr_points")):
            v = body[key]
            setattr(s, cnt, v[0])
            for k in range(v[0]):
                dst[k][0], dst[k][1] = v[1 + 2 * k], v[2 + 2 * k]
        for k, (key, dst) in enumerate((("cY", s.ar_coeffs_y), ("cCb", s.ar_coeffs_cb), ("cCr", s.ar_coeffs_cr))):
            v = body[key]
            for j, c in enumerate(v):
                dst[j] = c
            s.num_ar_coeffs_plus1[k] = len(v) + 1
        segs.append(GrainTableSegment.from_c(s))
        i += 8
    return segs
