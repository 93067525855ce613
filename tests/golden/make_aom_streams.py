"""Writes tests/golden/aom/stream_*.ivf + .json: short AV1 streams produced by libaom 3.13.1's encoder
(oracle/aom_encode.py) with film grain signalled, and what `inspect` must report for them.  The expected
parameters come from libaom's side (its test-vector table / the input grain table), not from our parser.

    python tests/golden/make_aom_streams.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import av1_writer as W  # noqa: E402
from grav1synth_b200 import inspect as I  # noqa: E402
from grav1synth_b200.diff import format_grain_table  # noqa: E402
from grav1synth_b200.grain_table import parse_grain_table  # noqa: E402
from oracle import aom_encode as E  # noqa: E402
from test_inspect_libaom import header_view, vector_view  # noqa: E402

OUT = os.path.join(HERE, "aom")
JOBS = {
    "stream_fgtest1_altref": dict(opts={"film-grain-test": "1", "auto-alt-ref": "1"}, lag=19, frames=12, fps=(24, 1)),
    "stream_fgtest9_lowdelay": dict(opts={"film-grain-test": "9"}, lag=0, frames=8, fps=(30000, 1001)),
    "stream_table_c2": dict(opts={"film-grain-table": os.path.join(HERE, "c2_small_8bit.tbl")}, lag=0, frames=6,
                            fps=(25, 1)),
}
for name, j in JOBS.items():
    w, h = 176, 144
    packets = E.encode(E.synthetic_frames(j["frames"], w, h, seed=len(name)), w, h, j["opts"], lag_in_frames=j["lag"])
    path = os.path.join(OUT, name + ".ivf")
    with open(path, "wb") as f:
        f.write(W.ivf(packets, w, h, j["fps"][0], j["fps"][1]))
    if "film-grain-test" in j["opts"]:
        params = vector_view(E.test_vector(int(j["opts"]["film-grain-test"])))
    else:
        seg = parse_grain_table(open(j["opts"]["film-grain-table"]).read())[0]
        params = header_view(type("H", (), {"params": seg})())
    p = I.BitstreamParser()
    fps = p.push_file(path)
    hs = p.get_grain_headers()
    for hdr in hs:
        if hdr.kind == I.UPDATE_GRAIN:
            assert json.loads(json.dumps(header_view(hdr))) == json.loads(json.dumps(params)), name
    out = dict(fps=list(fps), kinds=[hdr.kind for hdr in hs], params=params,
               table=format_grain_table(p.aggregate_grain_headers(*fps)))
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print(name, os.path.getsize(path), "bytes", out["kinds"])
