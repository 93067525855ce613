"""The `diff` command line (reference: src/main.rs:347-533) and its .y4m reader."""
import os

import numpy as np
import pytest

from helpers import ROOT, corpus_frames
from grav1synth_b200.y4m import Y4MReader, write_y4m


@pytest.mark.parametrize("name", ["c2_small_8bit", "c3_small_10bit", "yuv422_10bit", "odd_size_8bit"])
def test_y4m_round_trip(name, tmp_path):
    spec, fps, frames = corpus_frames(name)
    p = str(tmp_path / "a.y4m")
    write_y4m(p, [s for s, _ in frames], spec.bit_depth, fps, (spec.ss_x, spec.ss_y))
    r = Y4MReader(p)
    d = r.get_video_details()
    assert (d.width, d.height, d.bit_depth, d.ss_x, d.ss_y) == (spec.width, spec.height, spec.bit_depth, spec.ss_x, spec.ss_y)
    assert (d.fps_num, d.fps_den) == fps
    got = list(r)
    assert len(got) == len(frames)
    for planes, (s, _) in zip(got, frames):
        for a, b in zip(planes, s):
            assert a.dtype == b.dtype and np.array_equal(a, b)
    assert r.get_frame() is None


def test_cli_refuses_same_paths(tmp_path, caplog):
    from grav1synth_b200.__main__ import main
    p = str(tmp_path / "a.y4m")
    assert main(["diff", p, p, "-o", str(tmp_path / "o.tbl")]) == 0
    assert not os.path.exists(tmp_path / "o.tbl")
    assert main(["diff", p, str(tmp_path / "b.y4m"), "-o", p]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c2_small_8bit", "c3_small_10bit"])
def test_cli_diff_matches_golden(name, tmp_path):
    from grav1synth_b200.__main__ import main
    spec, fps, frames = corpus_frames(name)
    a, b, o = str(tmp_path / "src.y4m"), str(tmp_path / "den.y4m"), str(tmp_path / "out.tbl")
    write_y4m(a, [s for s, _ in frames], spec.bit_depth, fps)
    write_y4m(b, [d for _, d in frames], spec.bit_depth, fps)
    assert main(["diff", a, b, "-o", o, "-y"]) == 0
    with open(os.path.join(ROOT, "tests", "golden", name + ".tbl"), "rb") as f:
        assert open(o, "rb").read() == f.read()
