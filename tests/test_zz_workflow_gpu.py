"""The whole grav1synth workflow on a real codec pair, product code only on the GPU side (`-m gpu`, needs the bundled
libaom for the codec): source frames + libaom's decode of its own grainless encode of them (the "denoised" clip)
-> `diff` (CLI, .y4m in, CUDA engine) -> grain table -> `apply` (CLI) to the encoded stream -> libaom decodes the result
and synthesises grain again.  Checks: the table is the one libaom's own noise model returns for these frames, `inspect`
reads it back from the rewritten stream, and the grain libaom re-synthesises has about the strength of what the codec
removed."""
import json
import os
import struct

import pytest

from helpers import ROOT
from oracle import aom_pin

pytestmark = pytest.mark.gpu


def ivf_packets(data):
    off, packets = 32, []
    while off + 12 <= len(data):
        sz = struct.unpack_from("<I", data, off)[0]
        packets.append(data[off + 12: off + 12 + sz])
        off += 12 + sz
    return packets


@pytest.mark.skipif(not aom_pin.available()[0], reason="libaom pin unavailable")
def test_denoise_diff_apply_resynthesise(tmp_path):
    from aom_cases import codec_source
    from grav1synth_b200.__main__ import main
    from grav1synth_b200.grain_table import parse_grain_table
    from grav1synth_b200.y4m import write_y4m
    from oracle import aom_encode as E
    from test_aom_pin import seg_view

    ivf = os.path.join(ROOT, "tests", "golden", "aom", "codec_pair_cq28.ivf")
    source = codec_source()
    denoised = [list(f) for f in E.decode(ivf_packets(open(ivf, "rb").read()))]
    src_y4m, den_y4m, table, applied = (str(tmp_path / n) for n in ("src.y4m", "den.y4m", "grain.tbl", "applied.ivf"))
    write_y4m(src_y4m, source, 8, (24, 1), (1, 1))
    write_y4m(den_y4m, denoised, 8, (24, 1), (1, 1))

    assert main(["diff", src_y4m, den_y4m, "-o", table, "-y"]) == 0
    segs = parse_grain_table(open(table).read())
    with open(os.path.join(ROOT, "tests", "golden", "aom", "codec_pair_cq28.json")) as f:
        upstream = json.load(f)["segments"]
    assert [seg_view(s) for s in segs] == upstream          # libaom's noise model on the same frames

    assert main(["apply", ivf, "-o", applied, "-g", table, "-y"]) == 0
    back = str(tmp_path / "back.tbl")
    assert main(["inspect", applied, "-o", back, "-y"]) == 0
    strip = lambda text: [ln for ln in text.splitlines()[1:] if not ln.startswith("E ")]
    assert strip(open(back).read()) == strip(open(table).read())

    grainy = E.decode(ivf_packets(open(applied, "rb").read()))
    for k in range(len(source)):
        removed = (source[k][0].astype(float) - denoised[k][0]).std()
        resynth = (grainy[k][0].astype(float) - denoised[k][0]).std()
        assert 0.5 * removed < resynth < 2.5 * removed, (k, removed, resynth)
