"""cfg.host_narrow: host frames wider than 8 bits are reduced to 8 bits by the staging threads (frame_into_u8's
truncating shift, which is all the path reads), so half the bytes cross PCIe.  The CPU test pins the reduction
routine; the GPU test (last file of the suite on purpose: the option is new and off by default) checks that records and
tables are identical with and without it."""
import ctypes as C

import numpy as np
import pytest

from helpers import corpus_frames
from grav1synth_b200 import diff as D


def test_narrow_row_is_the_truncating_shift():
    L = D.lib()
    L.g1s_narrow_row.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.g1s_narrow_row.restype = None
    L.g1s_narrow_row_with.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.g1s_narrow_row_with.restype = None
    L.g1s_narrow_isa.restype = C.c_int
    rng = np.random.default_rng(0)
    # every code path this host has (plain loop, AVX2 and AVX-512BW with streaming stores), every destination alignment
    for isa in range(-1, L.g1s_narrow_isa() + 1):
        for n in (1, 7, 31, 32, 33, 64, 95, 1000, 3840):
            for shift in (0, 2, 4, 8):
                for off in (0, 1, 5, 31):
                    src = rng.integers(0, 1 << 16, n, dtype=np.uint16)
                    buf = np.full(n + 72, 0xAB, np.uint8)
                    dst = buf[off:]
                    if isa < 0:
                        L.g1s_narrow_row(dst.ctypes.data, src.ctypes.data, n, shift)
                    else:
                        L.g1s_narrow_row_with(dst.ctypes.data, src.ctypes.data, n, shift, isa)
                    assert np.array_equal(dst[:n], (src >> shift).astype(np.uint8))   # `as u8`: wraps, like the kernels' cast
                    assert np.all(dst[n:] == 0xAB) and np.all(buf[:off] == 0xAB)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c3_small_10bit", "heavy_grain_12bit"])
def test_host_narrow_gives_identical_records_and_tables(name):
    spec, fps, frames = corpus_frames(name)

    def run(**kw):
        g = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y, **kw)
        rl = D.RecordLayout(g.num_blocks)
        recs = []
        g.set_record_tap(lambda i, r: recs.append({k: np.array(v) for k, v in rl.unpack(np.array(r)).items()}))
        for s, d in frames:
            g.diff_frame(s, d)
        return g.finish(), recs

    plain, narrow = run(), run(host_narrow=True)
    assert narrow[0] == plain[0]
    assert len(narrow[1]) == len(plain[1]) == len(frames)
    for a, b in zip(plain[1], narrow[1]):
        m = a["flat"] != 0
        assert np.array_equal(a["flat"], b["flat"]) and a["num_flat"] == b["num_flat"]
        assert np.array_equal(a["score"].view(np.uint32), b["score"].view(np.uint32))
        assert np.array_equal(a["gram"], b["gram"]) and np.array_equal(a["nobs"], b["nobs"])
        assert np.array_equal(a["luma_sum"][m], b["luma_sum"][m])
        assert np.array_equal(a["rsum"][:, m], b["rsum"][:, m]) and np.array_equal(a["rsq"][:, m], b["rsq"][:, m])


@pytest.mark.gpu
def test_host_narrow_mixed_depths_and_device_frames_refused():
    import torch
    spec, fps, frames = corpus_frames("c3_small_10bit")
    a = D.DiffGenerator(24, 1, 10, 8, spec.width, spec.height)
    b = D.DiffGenerator(24, 1, 10, 8, spec.width, spec.height, host_narrow=True)
    for s, d in frames:
        d8 = [(p >> 2).astype(np.uint8) for p in d]
        a.diff_frame(s, d8)
        b.diff_frame(s, d8)
    assert a.finish() == b.finish()
    c = D.DiffGenerator(24, 1, 10, 10, spec.width, spec.height, host_narrow=True)
    t = [torch.zeros(p.shape, dtype=torch.int16, device="cuda") for p in frames[0][0]]
    with pytest.raises(D.G1SError):
        c.diff_frame_device([x.data_ptr() for x in t], [x.stride(0) * 2 for x in t],
                            [x.data_ptr() for x in t], [x.stride(0) * 2 for x in t])
