"""The reference's integration tests (/root/reference/tests/sanity_tests.rs: `inspect` :768-786, `apply` :1548-1592,
`remove` :2354-2393), on the IVF streams libaom produced for this repository instead of dav1d-test-data (an empty
submodule here): every clip must inspect; after `apply` of the reference's example table `inspect` must write a table;
after `remove` it must report that there is no film grain.  Same command lines, same log messages."""
import glob
import logging
import os

import pytest

from helpers import ROOT
from grav1synth_b200.__main__ import main
from test_oracle import EXAMPLE_TABLE

CLIPS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "aom", "*.ivf")))
assert CLIPS


@pytest.mark.parametrize("clip", CLIPS, ids=[os.path.basename(c) for c in CLIPS])
def test_inspect(clip, tmp_path, caplog):
    caplog.set_level(logging.INFO, logger="grav1synth")
    assert main(["inspect", clip, "-o", str(tmp_path / "t.tbl"), "-y"]) == 0
    assert "Done, wrote grain table" in caplog.text or "No film grain headers found" in caplog.text


@pytest.mark.parametrize("clip", CLIPS, ids=[os.path.basename(c) for c in CLIPS])
def test_apply(clip, tmp_path, caplog):
    caplog.set_level(logging.INFO, logger="grav1synth")
    grain_file, output = tmp_path / "example-table.tbl", tmp_path / "out.ivf"
    grain_file.write_text(EXAMPLE_TABLE)
    assert main(["apply", clip, "-o", str(output), "-g", str(grain_file), "-y"]) == 0
    caplog.clear()
    table = tmp_path / "t.tbl"
    assert main(["inspect", str(output), "-o", str(table), "-y"]) == 0
    assert "Done, wrote grain table" in caplog.text
    # the applied table comes back: same parameter lines (the E line carries the stream's own times and seeds)
    got = table.read_text().splitlines()
    assert got[0] == "filmgrn1" and got[2:9] == EXAMPLE_TABLE.splitlines()[2:9]


@pytest.mark.parametrize("clip", CLIPS, ids=[os.path.basename(c) for c in CLIPS])
def test_remove(clip, tmp_path, caplog):
    caplog.set_level(logging.INFO, logger="grav1synth")
    output = tmp_path / "out.ivf"
    assert main(["remove", clip, "-o", str(output), "-y"]) == 0
    caplog.clear()
    assert main(["inspect", str(output), "-o", str(tmp_path / "t.tbl"), "-y"]) == 0
    assert "No film grain headers found" in caplog.text
    assert not (tmp_path / "t.tbl").exists()
