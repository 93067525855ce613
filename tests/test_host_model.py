"""CPU tests of the product's HOST side (no GPU): the C-ABI library loads and exports every
symbol include/g1s.h declares, its writer reproduces the reference fixture, and the host noise
model (CONSUMER mode: records in, grain table out) agrees with the oracle when fed records built
independently in numpy."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import CORPUS, ROOT, corpus_frames, gram_to_pairs, numpy_record
from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from oracle import oracle as O
from test_oracle import EXAMPLE_TABLE, example_segment


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "g1s.h")).read()
    declared = set(re.findall(r"\b(g1s_[a-z_0-9]+)\s*\(", hdr)) - {"g1s_record_fn"}
    from grav1synth_b200 import inspect as I
    assert declared == set(D.EXPORTS) | set(I.EXPORTS)
    L = D.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.g1s_abi_version() == 2


def test_struct_sizes_match_header():
    # compile-time truth from a tiny C program would need gcc at test time; instead pin the layout we bind
    assert C.sizeof(abi.CSegment) == 8 + 8 + 16 + 8 + 28 + 20 + 20 + 24 + 25 + 25 + 2  # = 184, 8-byte aligned
    assert C.sizeof(abi.CFrame) == 3 * 8 + 3 * 8 + 8
    assert C.sizeof(abi.CDiffConfig) == 16 + 4 * 27 + 4  # 27 int32 fields (incl. reserved), padded to 8
    from grav1synth_b200.inspect import CStreamInfo
    assert C.sizeof(CStreamInfo) == 16 * 4 + 2 * 8 + 4 * 4   # g1s_stream_info


def test_writer_matches_reference_fixture(tmp_path):
    seg = example_segment()
    assert D.format_grain_table([seg]) == EXAMPLE_TABLE
    p = tmp_path / "x.tbl"
    D.write_grain_table([seg], str(p))
    assert p.read_bytes().decode() == EXAMPLE_TABLE


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(D.G1SError) as e:
        D.DiffGenerator(24, 1, 8, 8, 64, 64)
    assert e.value.code == abi.G1S_E_CUDA


def test_bad_config_rejected():
    with pytest.raises(D.G1SError) as e:
        D.DiffGenerator(24, 1, 7, 8, 64, 64, mode=abi.MODE_CONSUMER)
    assert e.value.code == abi.G1S_E_ARG


def test_div255_sequence():
    # the kernel's divide-free pix/255.0: q = p*inv; r = fma(-q, 255, p); q + r*inv — exact for all 256 inputs
    from fractions import Fraction
    inv = 1.0 / 255.0

    def fma(a, b, c):
        return float(Fraction(a) * Fraction(b) + Fraction(c))

    for p in range(256):
        q = float(p) * inv
        r = fma(-q, 255.0, float(p))
        assert fma(r, inv, q) == p / 255.0


def records_from_oracle(name):
    """Per-frame records built with numpy from the frames and the oracle's flat mask."""
    spec, fps, frames = corpus_frames(name)
    g = O.OracleDiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, O.GRAM_EXACT_INT, O.EXP_FIXED,
                              spec.ss_x, spec.ss_y)
    nb = ((spec.width + 31) // 32) * ((spec.height + 31) // 32)
    rl = D.RecordLayout(nb)
    recs = []
    for s, d in frames:
        g.diff_frame(s, d)
        flat, scores, _ = g.last_flat()
        r = numpy_record(s, d, spec.bit_depth, spec.bit_depth, spec.ss_x, spec.ss_y, flat)
        pairs = np.stack([gram_to_pairs(r["gram"][c]) for c in range(3)])
        recs.append(rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat))
    return spec, fps, recs, g.finish()


@pytest.mark.parametrize("name", list(CORPUS))
def test_host_model_matches_oracle(name):
    spec, fps, recs, want = records_from_oracle(name)
    h = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x,
                        spec.ss_y, mode=abi.MODE_CONSUMER)
    for r in recs:
        h.consume_record(r)
    got = h.finish()
    assert got == want
    assert D.format_grain_table(got) == open(os.path.join(ROOT, "tests", "golden", name + ".tbl")).read()


def test_consumer_rejects_frames_and_wrong_record_size():
    h = D.DiffGenerator(24, 1, 8, 8, 64, 64, mode=abi.MODE_CONSUMER)
    z = np.zeros((64, 64), np.uint8)
    c = np.zeros((32, 32), np.uint8)
    with pytest.raises(D.G1SError):
        h.diff_frame([z, c, c], [z, c, c])
    with pytest.raises(D.G1SError):
        h.consume_record(np.zeros(10, np.uint8))


def test_segment_cut_matches_oracle():
    from grav1synth_b200.synth import SynthSpec, make_pair_numpy
    a = SynthSpec(256, 192, 8, textured=0.0, sigma0=1.0, sigma1=0.5, ar_strength=0.0, seed=1)
    b = SynthSpec(256, 192, 8, textured=0.0, sigma0=2.2, sigma1=0.5, ar_strength=0.6, seed=2)
    frames = [make_pair_numpy(a, k) for k in range(3)] + [make_pair_numpy(b, k) for k in range(3)]
    g = O.OracleDiffGenerator(30000, 1001, 8, 8)
    h = D.DiffGenerator(30000, 1001, 8, 8, 256, 192, mode=abi.MODE_CONSUMER)
    rl = D.RecordLayout(8 * 6)
    for s, d in frames:
        g.diff_frame(s, d)
        flat, scores, _ = g.last_flat()
        r = numpy_record(s, d, 8, 8, 1, 1, flat)
        pairs = np.stack([gram_to_pairs(r["gram"][c]) for c in range(3)])
        h.consume_record(rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat))
    want, got = g.finish(), h.finish()
    assert len(want) >= 2 and got == want


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under grav1synth_b200/ may import, load or link it (and libg1s.so must not
    carry its symbols)."""
    import glob
    import subprocess
    for path in glob.glob(os.path.join(ROOT, "grav1synth_b200", "*.py")):
        src = open(path).read()
        assert "import oracle" not in src and "from oracle" not in src and "libg1s_oracle" not in src, path
    syms = subprocess.run(["nm", "-D", "--defined-only", D.LIB_PATH], capture_output=True, text=True).stdout
    assert "g1s_oracle" not in syms and "aom_noise_model" not in syms
