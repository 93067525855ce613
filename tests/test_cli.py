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


# ---- --filters: the reference's FilterChain::new tests (src/filters.rs:184-364), same inputs and messages
def _err(filters):
    from grav1synth_b200.filters import FilterChain, FilterError
    with pytest.raises(FilterError) as e:
        FilterChain(filters)
    return str(e.value)


def test_filter_chain_parsing_matches_reference():
    from grav1synth_b200.filters import Crop, FilterChain, Resize
    assert FilterChain("").filters == []
    assert FilterChain("crop:top=1,bottom=2,left=3,right=4").filters == [Crop(1, 2, 3, 4)]
    assert FilterChain("resize:width=1920,height=1080").filters == [Resize(1920, 1080, "catmullrom")]
    for alg in ("hermite", "catmullrom", "mitchell", "lanczos", "spline36"):
        assert FilterChain(f"resize:width=640,height=360,alg={alg}").filters == [Resize(640, 360, alg)]
    assert FilterChain("crop:top=4;resize:width=320,height=240,alg=lanczos").filters == [Crop(4, 0, 0, 0),
                                                                                          Resize(320, 240, "lanczos")]
    assert 'Invalid filter syntax in "crop"' in _err("crop")
    assert 'Unrecognized filter "rotate"' in _err("rotate:degrees=90")
    assert 'Invalid filter syntax in "top"' in _err("crop:top")
    assert 'Unrecognized crop arg "width"' in _err("crop:width=12")
    assert "invalid digit found in string" in _err("crop:top=abc")
    assert 'Invalid filter syntax in "width"' in _err("resize:width")
    assert 'Unrecognized resize arg "depth"' in _err("resize:width=1,height=1,depth=3")
    assert 'Unrecognized resize algorithm "nearest"' in _err("resize:width=1,height=1,alg=nearest")
    assert "Both width and height must be provided to resize filter" in _err("resize:width=10")
    assert "Both width and height must be provided to resize filter" in _err("resize:height=10")
    assert "invalid digit found in string" in _err("resize:width=abc,height=10")


def test_crop_is_plane_slicing():
    from grav1synth_b200.filters import FilterChain
    y = np.arange(64 * 48, dtype=np.uint8).reshape(48, 64)
    c = np.arange(32 * 24, dtype=np.uint8).reshape(24, 32)
    out = FilterChain("crop:top=4,bottom=8,left=2,right=6").apply([y, c, c], 1, 1)
    assert out[0].shape == (36, 56) and out[1].shape == (18, 28)
    assert np.array_equal(out[0], y[4:40, 2:58]) and np.array_equal(out[2], c[2:20, 1:29])


@pytest.mark.gpu
def test_cli_diff_with_crop_filter(tmp_path):
    """A padded source cropped back by --filters must give the same table as the unpadded source."""
    from grav1synth_b200.__main__ import main
    spec, fps, frames = corpus_frames("c2_small_8bit")
    pad = [[np.pad(s[0], ((4, 8), (2, 6)), mode="edge"), np.pad(s[1], ((2, 4), (1, 3)), mode="edge"),
            np.pad(s[2], ((2, 4), (1, 3)), mode="edge")] for s, _ in frames]
    a, b, o = str(tmp_path / "src.y4m"), str(tmp_path / "den.y4m"), str(tmp_path / "out.tbl")
    write_y4m(a, pad, 8, fps)
    write_y4m(b, [d for _, d in frames], 8, fps)
    assert main(["diff", a, b, "-o", o, "-y", "--filters", "crop:top=4,bottom=8,left=2,right=6"]) == 0
    with open(os.path.join(ROOT, "tests", "golden", "c2_small_8bit.tbl"), "rb") as f:
        assert open(o, "rb").read() == f.read()
