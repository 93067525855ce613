"""Strict mode on the device (`-m gpu`): gram_reforder_kernel reproduces the reference's per-term f64 accumulation
of the AR normal equations (A[i][j] += buf[i]*buf[j] / 255^2 behind /root/reference/src/main.rs:442) BIT FOR BIT,
so the engine's tables equal libaom's / the reference's on every stream -- including the fit_piecewise tie cases
the exact-integer fast path flips.  Everything goes through the C ABI; nothing reads /root/reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from aom_cases import CASES, EXACT_INT_TIE_FLIPS
from helpers import ROOT
from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from grav1synth_b200.synth import SynthSpec
from oracle import oracle as O
from test_aom_pin import digest, load_or_skip, seg_view
from test_strict import gramf_from_oracle

pytestmark = pytest.mark.gpu


def strict_run(frames, bd, ss, fps, batch=0):
    h, w = frames[0][0][0].shape
    g = D.DiffGenerator(fps[0], fps[1], bd, bd, w, h, ss[0], ss[1], batch_frames=batch, gram_order=abi.GRAM_REF_ORDER)
    raw = []
    g.set_record_tap(lambda i, r: raw.append(r))
    for s, d in frames:
        g.diff_frame(s, d)
    segs = g.finish()
    rl = D.RecordLayout(g.num_blocks)
    digests = [g.digest_from_record(r) for r in raw]
    return segs, [rl.unpack(r) for r in raw], digests, g


def latest_from_digest(dg, c):
    """(x, ar_gain, strength x, nobs) of channel c out of a LatestFrame digest (g1s_model.cpp to_digest)."""
    p = 4 + c * (375 + 2 + 80 + 2)
    n = 24 if c == 0 else 25
    x = dg[p + 350:p + 350 + n]
    gain, nobs = dg[p + 375], int(dg[p + 376])
    sx = dg[p + 377 + 60:p + 377 + 80]
    return x, gain, sx, nobs


@pytest.mark.parametrize("name", [n for n in CASES])
def test_strict_engine_is_bit_identical_to_libaom(name):
    """Per frame: flat map, observation counts and the f64 BIT PATTERNS of the latest AR solution / gain / strength
    solution; per segment every integer -- against what libaom 3.13.1's noise_model.c returned (tests/golden/aom)."""
    with open(os.path.join(ROOT, "tests", "golden", "aom", name + ".json")) as f:
        want = json.load(f)
    frames, bd, ss, fps = load_or_skip(name)
    segs, recs, digests, _ = strict_run(frames, bd, ss, fps)
    for k, (r, dg, wf) in enumerate(zip(recs, digests, want["frames"])):
        assert hashlib.sha256(np.ascontiguousarray(r["flat"]).tobytes()).hexdigest()[:16] == wf["flat_sha"], k
        if wf["status"] not in (0, 3):
            continue
        for c in range(3):
            x, gain, sx, nobs = latest_from_digest(dg, c)
            ws = wf["state"][f"latest{c}"]
            assert nobs == ws["nobs"], (k, c)
            assert float(gain).hex() == ws["ar_gain"], (k, c)
            assert digest(x, sx) == ws["digest"], (k, c)
    assert [seg_view(s) for s in segs] == want["segments"]


@pytest.mark.parametrize("name", ["c2_small_8bit", "yuv444_8bit", "yuv422_10bit", "heavy_grain_12bit",
                                  "saturated_residual", "segment_cut", "odd_size_8bit"])
def test_strict_sums_equal_the_reference_order_oracle(name):
    """Every one of the ~1000 chains per frame against the oracle's reference-order A and b: identical doubles."""
    frames, bd, ss, fps = load_or_skip(name)
    _, recs, _, _ = strict_run(frames, bd, ss, fps, batch=2)
    o = O.OracleDiffGenerator(fps[0], fps[1], bd, bd, O.GRAM_REF_ORDER, O.EXP_FIXED, ss[0], ss[1])
    iu = np.triu_indices(26)
    for k, ((s, d), r) in enumerate(zip(frames, recs)):
        o.diff_frame(s, d)
        if o.last_num_flat <= 1:
            continue
        want = gramf_from_oracle(o, ss)
        for c in range(3):
            live = np.ones((26, 26), bool)
            live[25, 25] = False
            if c == 0:
                live[24, :] = live[:, 24] = False
            m = live[iu]
            assert np.array_equal(r["gramf"][c][m].view(np.uint64), want[c][m].view(np.uint64)), (k, c)


def test_fast_mode_still_flips_the_tie_and_strict_mode_does_not():
    """The documented difference between the two modes, on the device: EXACT_INT picks another scaling-point list on
    the tie case, REF_ORDER gives libaom's."""
    name = sorted(EXACT_INT_TIE_FLIPS)[0]
    with open(os.path.join(ROOT, "tests", "golden", "aom", name + ".json")) as f:
        want = json.load(f)["segments"]
    frames, bd, ss, fps = load_or_skip(name)
    h, w = frames[0][0][0].shape
    fast = D.DiffGenerator(fps[0], fps[1], bd, bd, w, h, ss[0], ss[1])
    for s, d in frames:
        fast.diff_frame(s, d)
    got_fast = [seg_view(s) for s in fast.finish()]
    got_strict = [seg_view(s) for s in strict_run(frames, bd, ss, fps)[0]]
    assert got_strict == want
    strip = lambda s: {k: v for k, v in s.items() if not k.startswith("scaling_points")}
    assert [strip(s) for s in got_fast] == [strip(s) for s in want]
