"""`generate`: photon-noise grain segments (g1s_generate_photon_noise) -- CPU tests.

The reference's only fixture, /root/reference/tests/example-table.tbl (restated as EXAMPLE_TABLE in test_oracle.py),
is an output of this generator (libaom's photon_noise_table tool: seed 7391, gamma 2.8): reproduced byte for byte."""
import pytest

import av1_writer as W
from av1_writer import Frame, Grain, Seq
from grav1synth_b200 import inspect as I
from grav1synth_b200.diff import G1SError, format_grain_table
from oracle import aom_pin
from test_oracle import EXAMPLE_TABLE


@pytest.mark.parametrize("iso,w,h", [(750, 1920, 1080), (1700, 1280, 720), (1000, 1920, 816), (3800, 720, 576)])
def test_reference_fixture_is_reproduced(iso, w, h):
    seg = I.generate_photon_noise_params(0, 26460000000, iso, w, h, I.TRANSFER_BT470BG, False, 7391)
    assert format_grain_table([seg]) == EXAMPLE_TABLE


def test_crate_defaults_and_monotonic_behaviour():
    a = I.generate_photon_noise_params(0, 2 ** 64 - 1, 400, 1920, 1080)
    assert a.random_seed == 10956 and a.ar_coeff_lag == 0 and a.ar_coeff_shift == 6 and a.scaling_shift == 8
    assert a.ar_coeffs_y == [] and a.ar_coeffs_cb == [0] and a.ar_coeffs_cr == [0] and a.overlap_flag
    assert [p[0] for p in a.scaling_points_y] == [round(255 * i / 13) for i in range(14)]
    assert a.scaling_points_cb == [] and not a.chroma_scaling_from_luma
    assert I.generate_photon_noise_params(0, 1, 400, 1920, 1080, chroma_grain=True).chroma_scaling_from_luma
    # more sensitivity, or smaller photosites, means more noise at every luma level
    strength = lambda s: [p[1] for p in s.scaling_points_y]
    lo, hi = strength(a), strength(I.generate_photon_noise_params(0, 1, 6400, 1920, 1080))
    assert all(y >= x for x, y in zip(lo, hi)) and sum(hi) > sum(lo)
    uhd = strength(I.generate_photon_noise_params(0, 1, 400, 3840, 2160))
    assert all(y >= x for x, y in zip(lo, uhd)) and sum(uhd) > sum(lo)
    pq = strength(I.generate_photon_noise_params(0, 1, 400, 3840, 2160, I.TRANSFER_SMPTE2084))
    assert pq[0] > pq[-1] > 0           # PQ spends its code values on the shadows
    with pytest.raises(G1SError):
        I.generate_photon_noise_params(0, 1, 0, 1920, 1080)


def test_generate_command(tmp_path):
    seq = Seq(color_description=(9, 16, 9), bit_depth=10, profile=0)   # BT.2020 / PQ -> the SMPTE2084 curve
    packets = [W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, grain=Grain(kind="disable")).frame_obu(seq),
               W.temporal_delimiter() + Frame(frame_type=1, order_hint=1, grain=Grain(kind="disable")).frame_obu(seq)]
    src, out = tmp_path / "in.ivf", tmp_path / "out.ivf"
    src.write_bytes(W.ivf(packets, seq.width, seq.height, 24, 1))
    from grav1synth_b200.__main__ import main
    assert main(["generate", str(src), "-o", str(out), "--iso", "800", "--chroma"]) == 0
    p = I.BitstreamParser()
    p.push_file(str(out))
    hs = p.get_grain_headers()
    # the stream signals limited range, so the command passes full_range = false (src/main.rs:299: range == JPEG)
    assert p.stream_info()["color_range"] == 0
    want = I.generate_photon_noise_params(0, 1, 800, seq.width, seq.height, I.TRANSFER_SMPTE2084, True, full_range=False)
    full = I.generate_photon_noise_params(0, 1, 800, seq.width, seq.height, I.TRANSFER_SMPTE2084, True, full_range=True)
    assert want.scaling_points_y[0][0] == 16 and want.scaling_points_y[-1][0] == 235
    assert full.scaling_points_y[0][0] == 0 and full.scaling_points_y[-1][0] == 255
    assert [h.kind for h in hs] == [I.UPDATE_GRAIN] * 2
    for k, h in enumerate(hs):
        assert h.params.scaling_points_y == want.scaling_points_y and h.params.chroma_scaling_from_luma
        assert h.params.random_seed == (10956 * (k + 2)) % 65536 and h.params.ar_coeff_lag == 0
        # chroma_scaling_from_luma codes one coefficient per chroma plane at lag 0
        assert h.params.ar_coeffs_y == [] and h.params.ar_coeffs_cb == [0] and h.params.ar_coeffs_cr == [0]


@pytest.mark.skipif(not aom_pin.available()[0], reason="libaom pin unavailable")
def test_generated_grain_on_a_libaom_stream_decodes(tmp_path):
    import numpy as np
    from oracle import aom_encode as E
    fr = E.synthetic_frames(5, 176, 144, seed=2)
    plain = E.encode(fr, 176, 144, {}, lag_in_frames=0)
    seg = I.generate_photon_noise_params(0, 2 ** 64 - 1, 3200, 176, 144)
    rw = I.GrainRewriter([seg])
    out = [rw.rewrite_packet(p, k * 416667) for k, p in enumerate(plain)]
    a, b = E.decode(plain), E.decode(out)
    assert len(a) == len(b) == 5
    # luma-only grain: every luma plane changes, chroma planes do not
    assert all(not np.array_equal(x[0], y[0]) for x, y in zip(a, b))
    assert all(np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) for x, y in zip(a, b))
