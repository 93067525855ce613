"""Strict mode (g1s_diff_config.gram_order = G1S_GRAM_REF_ORDER) -- CPU half, `-m "not gpu"`.

The reference accumulates A[i][j] += buf[i]*buf[j] / 255^2 term by term in f64 behind differ.diff_frame
(/root/reference/src/main.rs:442); in strict mode the engine's record carries those sums (gramf, written on the
device by gram_reforder_kernel) and the host model loads them instead of the exact-integer Gram.  Here the
product's C++ host model (CONSUMER handle of libg1s.so) is fed records whose gramf comes from the oracle in
reference-order mode and must give libaom's tables on EVERY pinned case -- including the ones where the
exact-integer mode lands on the other side of a fit_piecewise tie (aom_cases.EXACT_INT_TIE_FLIPS).
The device half (the kernel reproduces those sums bit for bit) is tests/test_gpu_strict.py."""
import ctypes as C
import math

import numpy as np
import pytest

from aom_cases import CASES, EXACT_INT_TIE_FLIPS
from helpers import gram_to_pairs, numpy_record
from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from oracle import oracle as O
from test_aom_pin import golden, load_or_skip, seg_view

SLOW = {"hd_1080p_frame", "long_12_frames", "uhd_4k_10bit_frame"}


def gramf_from_oracle(o, ss):
    """[3][351] reference-order sums in the record's convention (tap 24 unscaled, tap 25 = centre sample)."""
    out = np.zeros((3, 351))
    iu = np.triu_indices(26)
    for c in range(3):
        A, b = o.last_eqns(c)
        n = A.shape[0]
        nss = float(1 << (ss[0] + ss[1])) if c else 1.0
        scale = np.ones(26)
        scale[24] = nss
        M = np.zeros((26, 26))
        M[:n, :n] = A * np.outer(scale[:n], scale[:n])
        M[:n, 25] = b * scale[:n]
        out[c] = M[iu]
    return out


def test_div65025_sequence():
    """q0 = p*y; r = fma(-q0, 65025, p); q = fma(r, y, q0) with y = RN(1/65025) is p / 65025.0 correctly rounded for
    every integer product a chain can see (|p| <= 1020^2: luma tap squared).  Checked here on a dense sample with
    exact rational arithmetic; the exhaustive loop is oracle/div65025_check.c."""
    from fractions import Fraction
    y = 1.0 / 65025.0
    rng = np.random.default_rng(0)
    ps = list(range(-2000, 2001)) + [int(v) for v in rng.integers(-1040400, 1040401, 4000)] + [65025, -65025, 1040400, -1040400]
    for p in ps:
        pd = float(p)
        q0 = pd * y
        r = math.fma(-q0, 65025.0, pd) if hasattr(math, "fma") else float(Fraction(pd) - Fraction(q0) * 65025)
        q = math.fma(r, y, q0) if hasattr(math, "fma") else float(Fraction(r) * Fraction(y) + Fraction(q0))
        assert q == pd / 65025.0, p


@pytest.mark.parametrize("name", [n for n in CASES if n not in SLOW])
def test_strict_host_model_gives_libaom_tables(name):
    want = golden(name)
    frames, bd, ss, fps = load_or_skip(name)
    h, w = frames[0][0][0].shape
    o = O.OracleDiffGenerator(fps[0], fps[1], bd, bd, O.GRAM_REF_ORDER, O.EXP_FIXED, ss[0], ss[1])
    c = D.DiffGenerator(fps[0], fps[1], bd, bd, w, h, ss[0], ss[1], mode=abi.MODE_CONSUMER, gram_order=abi.GRAM_REF_ORDER)
    rl = D.RecordLayout(((w + 31) // 32) * ((h + 31) // 32))
    for s, d in frames:
        o.diff_frame(s, d)
        flat, scores, _ = o.last_flat()
        r = numpy_record(s, d, bd, bd, ss[0], ss[1], flat)
        pairs = np.stack([gram_to_pairs(r["gram"][k]) for k in range(3)])
        c.consume_record(rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat,
                                 gramf=gramf_from_oracle(o, ss)))
    assert [seg_view(s) for s in c.finish()] == want["segments"]


def test_tie_flip_cases_are_still_covered():
    assert EXACT_INT_TIE_FLIPS and EXACT_INT_TIE_FLIPS <= set(CASES)


def test_strict_mode_rejects_unknown_order():
    with pytest.raises(D.G1SError) as e:
        D.DiffGenerator(24, 1, 8, 8, 64, 64, mode=abi.MODE_CONSUMER, gram_order=7)
    assert e.value.code == abi.G1S_E_ARG
