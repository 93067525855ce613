"""Source filters (`diff --filters`) -- CPU half: the resize weights the device uses
(g1s_resize_table -> csrc/g1s_filters.cu::build_resize_table) against the numpy statement of the published kernels
(oracle/resize_oracle.py), and properties of those weights.  Parity with the reference's video-resize 0.2.0 is UNPINNED
(crate absent); the device half (tests/test_filters_gpu.py) pins the CUDA chain to the oracle bit for bit."""
import numpy as np
import pytest

from grav1synth_b200 import diff as D
from oracle import resize_oracle as R

CASES = [(64, 32), (64, 48), (48, 64), (1920, 1280), (101, 37), (37, 101), (16, 16), (5, 40), (40, 5)]


@pytest.mark.parametrize("alg", R.ALGS)
@pytest.mark.parametrize("src,dst", CASES)
def test_device_tables_follow_the_published_kernels(alg, src, dst):
    left, coef = D.resize_table(alg, src, dst)
    lo, co = R.table(alg, src, dst)
    assert np.array_equal(left, lo)
    assert coef.shape == co.shape
    # identical construction in C++ and Python; sin() may differ in the last place between the two libms
    assert np.allclose(coef, co, rtol=0, atol=2e-7)
    assert np.allclose(coef.astype(np.float64).sum(axis=1), 1.0, atol=1e-5)       # weights are normalised
    assert (left >= 0).all() and (left + coef.shape[1] <= src).all()              # every tap reads inside the plane


@pytest.mark.parametrize("alg", R.ALGS)
def test_identity_and_constant_planes(alg):
    left, coef = D.resize_table(alg, 40, 40)   # same size: an interpolating kernel puts everything on the centre tap
    for i in range(40):
        row = np.zeros(40)
        row[left[i]:left[i] + coef.shape[1]] = coef[i]
        assert abs(row.sum() - 1.0) < 1e-5
        if alg != "mitchell":  # B = 1/3 smooths: 8/9 on the centre, 1/18 either side
            assert abs(row[i] - 1.0) < 1e-6
        elif 2 <= i < 38:  # (at the frame edge the mirrored tap folds onto the centre)
            assert abs(row[i] - 8.0 / 9.0) < 1e-6
    flat = [np.full((48, 64), 517, np.uint16), np.full((24, 32), 300, np.uint16), np.full((24, 32), 700, np.uint16)]
    out = R.resize_planes(flat, 96, 40, alg, 10, (1, 1))
    assert [o.shape for o in out] == [(40, 96), (20, 48), (20, 48)]
    assert (out[0] == 517).all() and (out[1] == 300).all() and (out[2] == 700).all()


def test_resize_stays_in_range_and_crop_is_slicing():
    rng = np.random.default_rng(0)
    planes = [rng.integers(0, 1024, (48, 64), dtype=np.uint16), rng.integers(0, 1024, (24, 32), dtype=np.uint16),
              rng.integers(0, 1024, (24, 32), dtype=np.uint16)]
    planes[0][:8, :8] = 1023
    planes[0][8:16, :8] = 0
    for alg in R.ALGS:
        out = R.resize_planes(planes, 96, 72, alg, 10, (1, 1))
        assert all(int(o.max()) <= 1023 for o in out)
    c = R.crop_planes(planes, 4, 8, 2, 6, (1, 1))
    assert np.array_equal(c[0], planes[0][4:40, 2:58]) and np.array_equal(c[1], planes[1][2:20, 1:29])
