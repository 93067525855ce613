"""`inspect` path (BASELINE configs[0]) -- CPU tests of csrc/g1s_obu.cpp through the C ABI.

Two kinds of evidence:
 * the reference's OWN unit-test vectors for the syntax groups (the bit patterns and expected values of the
   #[test]s in /root/reference/src/parser/grain.rs:366-706 and frame.rs:2181-3860), replayed through the
   g1s_obu_probe test hook -- these pin the parser to the reference;
 * whole streams built by tests/av1_writer.py (an independent header ENCODER), pushed as packets / IVF / .obu
   files, checked header by header and against a Python statement of aggregate_grain_headers
   (/root/reference/src/main.rs:713-772).
"""
import math

import pytest

import av1_writer as W
from av1_writer import BitBuilder, Frame, Grain, Seq
from grav1synth_b200 import abi
from grav1synth_b200 import inspect as I
from grav1synth_b200.diff import G1SError, format_grain_table


def probe(what, bits: BitBuilder, args=()):
    """Runs one syntax group on `bits` + trailer; asserts that exactly the pushed bits were consumed."""
    data, n = bits.with_trailer()
    used, out, seg = I.probe(what, data, list(args))
    assert used == n, f"{what}: consumed {used} bits, the vector holds {n}"
    return out, seg


def B():
    return BitBuilder()


# ---------------------------------------------------------------- film_grain_params: grain.rs:366-706
def grain_of(seg):
    g = abi.GrainTableSegment.from_c(seg)
    return g, bool(seg.clip_to_restricted_range)


def test_film_grain_not_allowed_consumes_nothing():
    out, _ = probe("film_grain_params", B(), [0, 0, 0, 1, 1])
    assert out[0] == I.DISABLE


def test_film_grain_apply_grain_false():
    out, _ = probe("film_grain_params", B().push_bool(False), [1, 0, 0, 1, 1])
    assert out[0] == I.DISABLE


def test_film_grain_inter_copy_ref_frame():
    b = B().push_bool(True).push_bits(0x1234, 16).push_bool(False).push_bits(0b101, 3)
    out, _ = probe("film_grain_params", b, [1, 1, 0, 0, 0])
    assert out[0] == I.COPY_REF_FRAME


def test_film_grain_key_frame_420_zero_luma_shortcut():
    b = B().push_bool(True).push_bits(0xBEEF, 16).push_bits(0, 4).push_bool(False).push_bits(0b10, 2)
    b.push_bits(0, 2).push_bits(0b01, 2).push_bits(0b10, 2).push_bool(True).push_bool(False)
    out, seg = probe("film_grain_params", b, [1, 0, 0, 1, 1])
    g, clip = grain_of(seg)
    assert out[0] == I.UPDATE_GRAIN and g.random_seed == 0xBEEF
    assert g.scaling_points_y == [] and g.scaling_points_cb == [] and g.scaling_points_cr == []
    assert not g.chroma_scaling_from_luma and g.ar_coeff_lag == 0
    assert g.ar_coeffs_y == [] and g.ar_coeffs_cb == [0] and g.ar_coeffs_cr == [0]
    # grain_scaling_minus_8 = 0b10 -> 10 by the reference's code (grain.rs:276) and the AV1 spec; the reference's
    # own assertion for this vector says 8 (grain.rs:462), which its code cannot produce
    assert (g.scaling_shift, g.ar_coeff_shift, g.grain_scale_shift) == (10, 7, 2)
    assert (g.cb_mult, g.cb_luma_mult, g.cb_offset, g.cr_mult, g.cr_luma_mult, g.cr_offset) == (0,) * 6
    assert g.overlap_flag and not clip


def test_film_grain_chroma_scaling_from_luma():
    b = B().push_bool(True).push_bits(0x3456, 16).push_bool(True).push_bits(1, 4).push_bits(7, 8).push_bits(9, 8)
    b.push_bool(True).push_bits(0, 2).push_bits(1, 2)
    for c in [1, 2, 3, 4] + [5, 6, 7, 8, 9] + [-1, -2, -3, -4, -5]:
        b.push_bits(c + 128, 8)
    b.push_bits(0b11, 2).push_bits(0b01, 2).push_bool(True).push_bool(False)
    out, seg = probe("film_grain_params", b, [1, 1, 0, 0, 0])
    g, clip = grain_of(seg)
    assert g.scaling_points_y == [(7, 9)] and g.chroma_scaling_from_luma
    assert g.scaling_points_cb == [] and g.scaling_points_cr == [] and g.ar_coeff_lag == 1
    assert g.ar_coeffs_y == [1, 2, 3, 4] and g.ar_coeffs_cb == [5, 6, 7, 8, 9] and g.ar_coeffs_cr == [-1, -2, -3, -4, -5]
    assert (g.ar_coeff_shift, g.grain_scale_shift) == (9, 1) and g.overlap_flag and not clip


def test_film_grain_chroma_points_and_multipliers():
    b = B().push_bool(True).push_bits(0x4567, 16).push_bool(True).push_bits(1, 4).push_bits(1, 8).push_bits(2, 8)
    b.push_bool(False).push_bits(1, 4).push_bits(30, 8).push_bits(40, 8).push_bits(1, 4).push_bits(50, 8).push_bits(60, 8)
    b.push_bits(0b01, 2).push_bits(0, 2).push_bits(3 + 128, 8).push_bits(-3 + 128, 8).push_bits(0b10, 2).push_bits(0b10, 2)
    b.push_bits(11, 8).push_bits(22, 8).push_bits(0x1AB, 9).push_bits(33, 8).push_bits(44, 8).push_bits(0x055, 9)
    b.push_bool(True).push_bool(True)
    out, seg = probe("film_grain_params", b, [1, 1, 0, 0, 0])
    g, clip = grain_of(seg)
    assert g.scaling_points_y == [(1, 2)] and g.scaling_points_cb == [(30, 40)] and g.scaling_points_cr == [(50, 60)]
    assert g.ar_coeff_lag == 0 and g.ar_coeffs_y == [] and g.ar_coeffs_cb == [3] and g.ar_coeffs_cr == [-3]
    assert (g.scaling_shift, g.ar_coeff_shift, g.grain_scale_shift) == (9, 8, 2)
    assert (g.cb_mult, g.cb_luma_mult, g.cb_offset) == (11, 22, 0x1AB)
    assert (g.cr_mult, g.cr_luma_mult, g.cr_offset) == (33, 44, 0x055) and g.overlap_flag and clip


def test_film_grain_num_pos_luma_for_chroma_when_no_luma_points():
    b = B().push_bool(True).push_bits(0x5678, 16).push_bool(True).push_bits(0, 4).push_bool(False)
    b.push_bits(1, 4).push_bits(70, 8).push_bits(80, 8).push_bits(1, 4).push_bits(90, 8).push_bits(100, 8)
    b.push_bits(0, 2).push_bits(1, 2)
    for c in [10, 11, 12, 13] + [-10, -11, -12, -13]:
        b.push_bits(c + 128, 8)
    b.push_bits(1, 2).push_bits(0, 2)
    b.push_bits(1, 8).push_bits(2, 8).push_bits(3, 9).push_bits(4, 8).push_bits(5, 8).push_bits(6, 9)
    b.push_bool(False).push_bool(False)
    out, seg = probe("film_grain_params", b, [1, 1, 0, 0, 0])
    g, clip = grain_of(seg)
    assert g.scaling_points_y == [] and g.scaling_points_cb == [(70, 80)] and g.scaling_points_cr == [(90, 100)]
    assert g.ar_coeff_lag == 1 and g.ar_coeffs_y == []
    assert g.ar_coeffs_cb == [10, 11, 12, 13] and g.ar_coeffs_cr == [-10, -11, -12, -13]
    assert (g.ar_coeff_shift, g.grain_scale_shift) == (7, 0)
    assert (g.cb_mult, g.cb_luma_mult, g.cb_offset, g.cr_mult, g.cr_luma_mult, g.cr_offset) == (1, 2, 3, 4, 5, 6)
    assert not g.overlap_flag and not clip


def test_film_grain_monochrome_skips_chroma_fields():
    b = B().push_bool(True).push_bits(0x2345, 16).push_bits(2, 4)
    b.push_bits(10, 8).push_bits(20, 8).push_bits(30, 8).push_bits(40, 8)
    b.push_bits(0b11, 2).push_bits(1, 2)
    for c in [-2, -1, 0, 1]:
        b.push_bits(c + 128, 8)
    b.push_bits(0, 2).push_bits(3, 2).push_bool(False).push_bool(True)
    out, seg = probe("film_grain_params", b, [1, 0, 1, 1, 1])
    g, clip = grain_of(seg)
    assert g.scaling_points_y == [(10, 20), (30, 40)] and not g.chroma_scaling_from_luma
    assert g.scaling_points_cb == [] and g.ar_coeffs_y == [-2, -1, 0, 1] and g.ar_coeffs_cb == [0] and g.ar_coeffs_cr == [0]
    assert (g.scaling_shift, g.ar_coeff_shift, g.grain_scale_shift) == (11, 6, 3) and not g.overlap_flag and clip


# ---------------------------------------------------------------- frame.rs helper vectors (frame.rs:2181-3860)
def test_tile_info_vectors():
    out, _ = probe("tile_info", B().push_bool(True), [0, 8, 8])                      # uniform_single_tile_64
    assert out[:4] == [1, 1, 0, 0]
    b = B().push_bool(True).push_bool(True).push_bool(False).push_bool(False).push_bits(0, 1).push_bits(0, 2)
    out, _ = probe("tile_info", b, [0, 1024, 544])                                   # uniform_multi_tile
    assert out[:4] == [2, 1, 1, 0]
    out, _ = probe("tile_info", B().push_bool(False), [0, 8, 8])                     # non_uniform_small_frame
    assert out[:4] == [1, 1, 0, 0]
    out, _ = probe("tile_info", B().push_bool(True).push_bool(False).push_bool(False), [1, 480, 272])  # 128x128 SB
    assert out[:4] == [1, 1, 0, 0]
    # not from the reference: 7 superblock columns split at log2 = 1 are tiles of 4 + 3, i.e. TWO tiles (spec 5.9.15
    # counts tile starts; the reference's floor division, frame.rs:1106, would say one)
    b = B().push_bool(True).push_bool(True).push_bool(False).push_bool(False).push_bits(0, 1).push_bits(0, 2)
    out, _ = probe("tile_info", b, [0, 100, 56])
    assert out[:4] == [2, 1, 1, 0]


def test_quantization_params_vectors():
    out, _ = probe("quantization_params", B().push_bits(128, 8).push_bool(False).push_bool(False), [1, 0])
    assert out[:6] == [128, 0, 0, 0, 0, 0]
    b = B().push_bits(100, 8).push_bool(False).push_bool(False).push_bool(False).push_bool(False)
    out, _ = probe("quantization_params", b, [3, 0])
    assert out[0] == 100 and out[4] == out[2] and out[5] == out[3]
    b = B().push_bits(50, 8).push_bool(False).push_bool(True).push_bool(True).push_su(10, 7).push_bool(False)
    b.push_bool(True).push_su(-5, 7).push_bool(False).push_bool(False)
    out, _ = probe("quantization_params", b, [3, 1])
    assert (out[2], out[4], out[3], out[5]) == (10, -5, 0, 0)
    b = B().push_bits(50, 8).push_bool(False).push_bool(False).push_bool(True).push_su(7, 7).push_bool(False).push_bool(False)
    out, _ = probe("quantization_params", b, [3, 1])
    assert (out[2], out[4]) == (7, 7)
    b = B().push_bits(128, 8).push_bool(False).push_bool(False).push_bool(False).push_bool(False).push_bool(True)
    b.push_bits(5, 4).push_bits(3, 4).push_bits(7, 4)
    probe("quantization_params", b, [3, 1])                                          # qmatrix_separate_uv
    b = B().push_bits(128, 8).push_bool(False).push_bool(False).push_bool(False).push_bool(True).push_bits(5, 4).push_bits(3, 4)
    probe("quantization_params", b, [3, 0])                                          # qmatrix_shared_uv


def test_segmentation_params_vectors():
    out, _ = probe("segmentation_params", B().push_bool(False), [7])
    assert out[0] == 0
    b = B().push_bool(True)
    for _ in range(64):
        b.push_bool(False)
    out, _ = probe("segmentation_params", b, [7, 0, 0])
    assert out[:2] == [1, 0]
    b = B().push_bool(True)
    for _ in range(5):
        b.push_bool(False)
    b.push_bool(True).push_bits(5, 3).push_bool(False).push_bool(False)
    for _ in range(56):
        b.push_bool(False)
    out, _ = probe("segmentation_params", b, [7, 0, 5])
    assert out[:3] == [1, 1, 5]
    b = B().push_bool(True).push_bool(True).push_su(-50, 9)
    for _ in range(7 + 56):
        b.push_bool(False)
    out, _ = probe("segmentation_params", b, [7, 0, 0])
    assert out[:3] == [1, 1, -50]
    b = B().push_bool(True).push_bool(True).push_bool(True).push_bool(True)
    for _ in range(64):
        b.push_bool(False)
    assert probe("segmentation_params", b, [0])[0][0] == 1                           # update_map_and_data
    assert probe("segmentation_params", B().push_bool(True).push_bool(False).push_bool(False), [0])[0][0] == 1
    b = B().push_bool(True).push_bool(False).push_bool(True)
    for _ in range(64):
        b.push_bool(False)
    assert probe("segmentation_params", b, [0])[0][0] == 1                           # no_map_but_data


def test_delta_q_and_delta_lf_vectors():
    assert probe("delta_q_params", B(), [0])[0][0] == 0
    assert probe("delta_q_params", B().push_bool(True).push_bits(2, 2), [100])[0][0] == 1
    assert probe("delta_q_params", B().push_bool(False), [100])[0][0] == 0
    probe("delta_lf_params", B(), [0, 0])
    probe("delta_lf_params", B(), [1, 1])
    probe("delta_lf_params", B().push_bool(True).push_bits(1, 2).push_bool(False), [1, 0])
    probe("delta_lf_params", B().push_bool(False), [1, 0])


def test_loop_filter_cdef_lr_vectors():
    probe("loop_filter_params", B(), [1, 0, 3])
    probe("loop_filter_params", B(), [0, 1, 3])
    probe("loop_filter_params", B().push_bits(0, 6).push_bits(0, 6).push_bits(2, 3).push_bool(False), [0, 0, 1])
    b = B().push_bits(10, 6).push_bits(5, 6).push_bits(3, 6).push_bits(7, 6).push_bits(4, 3).push_bool(False)
    probe("loop_filter_params", b, [0, 0, 3])
    b = B().push_bits(0, 6).push_bits(5, 6).push_bits(1, 6).push_bits(2, 6).push_bits(0, 3).push_bool(False)
    probe("loop_filter_params", b, [0, 0, 3])
    b = B().push_bits(10, 6).push_bits(0, 6).push_bits(3, 3).push_bool(True).push_bool(True).push_bool(True).push_su(5, 7)
    for _ in range(7 + 2):
        b.push_bool(False)
    probe("loop_filter_params", b, [0, 0, 1])
    probe("loop_filter_params", B().push_bits(10, 6).push_bits(0, 6).push_bits(3, 3).push_bool(True).push_bool(False), [0, 0, 1])
    probe("cdef_params", B(), [1, 0, 1, 3])
    probe("cdef_params", B(), [0, 1, 1, 3])
    probe("cdef_params", B(), [0, 0, 0, 3])
    probe("cdef_params", B().push_bits(1, 2).push_bits(0, 2).push_bits(5, 4).push_bits(1, 2), [0, 0, 1, 1])
    b = B().push_bits(2, 2).push_bits(1, 2)
    for _ in range(2):
        b.push_bits(3, 4).push_bits(1, 2).push_bits(2, 4).push_bits(0, 2)
    probe("cdef_params", b, [0, 0, 1, 3])
    b = B().push_bits(0, 2).push_bits(3, 2)
    for _ in range(8):
        b.push_bits(0, 12)
    probe("cdef_params", b, [0, 0, 1, 3])
    probe("lr_params", B(), [1, 0, 1, 0, 3, 1, 1])
    probe("lr_params", B(), [0, 1, 1, 0, 3, 1, 1])
    probe("lr_params", B(), [0, 0, 0, 0, 3, 1, 1])
    probe("lr_params", B().push_bits(0, 6), [0, 0, 1, 0, 3, 1, 1])
    probe("lr_params", B().push_bits(1, 2).push_bits(0, 4).push_bool(True).push_bool(False), [0, 0, 1, 0, 3, 1, 1])
    probe("lr_params", B().push_bits(0, 2).push_bits(1, 2).push_bits(0, 2).push_bool(False).push_bool(True), [0, 0, 1, 1, 3, 1, 1])
    probe("lr_params", B().push_bits(0, 2).push_bits(1, 2).push_bits(0, 2).push_bool(True), [0, 0, 1, 1, 3, 0, 0])


def test_skip_mode_vectors():
    def args(intra, refsel, ohb, oh, hints, idx):
        return [intra, refsel, ohb, oh] + list(hints) + list(idx)
    z8, z7 = [0] * 8, [0] * 7
    assert probe("skip_mode_params", B(), args(1, 1, 4, 10, z8, z7))[0][0] == 0
    assert probe("skip_mode_params", B(), args(0, 0, 4, 10, z8, z7))[0][0] == 0
    assert probe("skip_mode_params", B(), args(0, 1, 0, 10, z8, z7))[0][0] == 0
    h = [5, 12, 0, 0, 0, 0, 0, 0]
    assert probe("skip_mode_params", B().push_bool(False), args(0, 1, 4, 10, h, [0, 1, 0, 0, 0, 0, 0]))[0][0] == 1
    h = [5, 3, 0, 0, 0, 0, 0, 0]
    assert probe("skip_mode_params", B().push_bool(True), args(0, 1, 4, 10, h, [0, 1, 0, 0, 0, 0, 0]))[0][0] == 1
    h = [5, 0, 0, 0, 0, 0, 0, 0]
    assert probe("skip_mode_params", B(), args(0, 1, 4, 10, h, z7))[0][0] == 0


def test_relative_dist_recenter_subexp_vectors():
    for (a, b_, n), want in [((5, 3, 0), 0), ((5, 3, 4), 2), ((3, 5, 4), -2), ((15, 1, 4), -2), ((1, 15, 4), 2), ((5, 5, 4), 0)]:
        assert I.probe("get_relative_dist", b"", [a, b_, n])[1][0] == want
    for (r, v), want in [((3, 7), 7), ((10, 5), 7), ((10, 4), 12), ((5, 0), 5), ((5, 1), 4), ((0, 3), 3)]:
        assert I.probe("inverse_recenter", b"", [r, v])[1][0] == want
    assert probe("decode_subexp", B().push_ns(5, 10), [10])[0][0] == 5
    assert probe("decode_subexp", B().push_bool(False).push_bits(5, 3), [100])[0][0] == 5
    assert probe("decode_subexp", B().push_bool(True).push_ns(4, 22), [30])[0][0] == 12
    assert probe("decode_signed_subexp_with_ref", B().push_ns(0, 21), [-10, 11, 0])[0][0] == 0
    assert probe("decode_unsigned_subexp_with_ref", B().push_ns(0, 20), [20, 3])[0][0] == 3
    assert probe("decode_unsigned_subexp_with_ref", B().push_ns(0, 20), [20, 15])[0][0] == 15
    for n in [2, 3, 5, 8, 10, 16, 100]:                                              # push_ns_roundtrips_through_ns_parser
        for v in range(n):
            assert probe("ns", B().push_ns(v, n), [n])[0][0] == v
    for width, vals in [(7, [0, 1, -1, 63, -64]), (9, [0, 255, -256, -50])]:         # push_su_roundtrips
        for v in vals:
            assert probe("su", B().push_su(v, width), [width])[0][0] == v


def test_global_motion_vectors():
    probe("global_motion_params", B(), [1, 1])
    probe("global_motion_params", B().push_bits(0, 7), [0, 1])
    b = B().push_bool(True).push_bool(False).push_bool(True)
    for _ in range(2):
        b.push_bool(False).push_bits(0, 3)
    probe("global_motion_params", b.push_bits(0, 6), [0, 1])                         # single_translation
    b = B().push_bool(True).push_bool(True)
    for _ in range(4):
        b.push_bool(False).push_bits(0, 3)
    probe("global_motion_params", b.push_bits(0, 6), [0, 1])                         # rotzoom
    b = B().push_bool(True).push_bool(False).push_bool(False)
    for _ in range(6):
        b.push_bool(False).push_bits(0, 3)
    probe("global_motion_params", b.push_bits(0, 6), [0, 1])                         # affine


def test_leb128_and_uvlc():
    assert I.probe("leb128", bytes([0xE5, 0x8E, 0x26]), [])[:2][1][0] == 624485
    assert I.probe("leb128", bytes([0xE5, 0x8E, 0x26]), [])[0] == 24
    assert I.probe("leb128", bytes([0x80]), [])[0] == abi.G1S_E_STREAM
    for v in [0, 1, 2, 5, 100, 65535]:
        assert probe("uvlc", B().push_uvlc(v), [])[0][0] == v


# ---------------------------------------------------------------- sequence header
@pytest.mark.parametrize("seq,want", [
    (Seq(), dict(profile=0, bd=8, planes=3, ss=(1, 1), fg=1, ohb=7, w=352, h=288)),
    (Seq(profile=1, bit_depth=10, width=1920, height=1080, order_hint_bits=0, film_grain_params_present=False),
     dict(profile=1, bd=10, planes=3, ss=(0, 0), fg=0, ohb=0, w=1920, h=1080)),
    (Seq(profile=2, bit_depth=12, ss=(1, 0), width=3840, height=2160, timing_info=True, decoder_model=True,
         operating_point_idc=(0x101, 0x301), frame_id_numbers=True, use_128=True),
     dict(profile=2, bd=12, planes=3, ss=(1, 0), fg=1, ohb=7, w=3840, h=2160)),
    (Seq(profile=2, bit_depth=10, width=640, height=360), dict(profile=2, bd=10, planes=3, ss=(1, 0), fg=1, ohb=7, w=640, h=360)),
    (Seq(monochrome=True, choose_screen_content_tools=False), dict(profile=0, bd=8, planes=1, ss=(1, 1), fg=1, ohb=7, w=352, h=288)),
    (Seq(color_description=(1, 13, 0), profile=1), dict(profile=1, bd=8, planes=3, ss=(0, 0), fg=1, ohb=7, w=352, h=288)),
    (Seq(reduced_still_picture_header=True, width=64, height=64),
     dict(profile=0, bd=8, planes=3, ss=(1, 1), fg=1, ohb=0, w=64, h=64)),
])
def test_sequence_header_variants(seq, want):
    payload = seq.payload()
    used, out, _ = I.probe("sequence_header", payload, [])
    assert used > 0 and used <= len(payload) * 8 and len(payload) * 8 - used <= 8   # trailing bit + padding only
    assert out[0] == want["profile"] and out[1] == want["bd"] and out[2] == want["planes"]
    assert (out[3], out[4]) == want["ss"] and out[5] == want["fg"] and out[6] == want["ohb"]
    assert (out[7], out[8]) == (want["w"], want["h"])
    assert out[9] == int(seq.reduced_still_picture_header)
    assert out[10] == len(seq.operating_point_idc) - 1 and out[11] == seq.operating_point_idc[0]
    assert out[12] == int(seq.timing_info and seq.decoder_model) and out[13] == int(seq.frame_id_numbers)


# ---------------------------------------------------------------- whole streams
def py_aggregate(headers, fps):
    """aggregate_grain_headers (src/main.rs:713-772) restated; headers = [(kind, params_dict_or_None)]."""
    t = fps[1] / fps[0] * 10_000_000.0
    start, end_f = 0, t
    end = math.ceil(end_f)
    acc = []
    for kind, p in headers:
        has = bool(acc) and acc[-1]["end"] == start
        if has:
            if kind == I.COPY_REF_FRAME:
                acc[-1]["end"] = end
            elif kind == I.UPDATE_GRAIN:
                same = {k: v for k, v in p.items() if k != "random_seed"} == \
                       {k: v for k, v in acc[-1]["p"].items() if k != "random_seed"}
                if same:
                    acc[-1]["end"] = end
                else:
                    acc.append(dict(start=start, end=end, p=p))
        elif kind == I.UPDATE_GRAIN:
            acc.append(dict(start=start, end=end, p=p))
        start = end
        end_f += t
        end = math.ceil(end_f)
    return acc


def header_view(h):
    if h.kind != I.UPDATE_GRAIN:
        return (h.kind, None)
    g = h.params
    d = {k: getattr(g, k) for k in (
        "random_seed", "scaling_points_y", "scaling_points_cb", "scaling_points_cr", "chroma_scaling_from_luma",
        "scaling_shift", "ar_coeff_lag", "ar_coeffs_y", "ar_coeffs_cb", "ar_coeffs_cr", "ar_coeff_shift",
        "grain_scale_shift", "cb_mult", "cb_luma_mult", "cb_offset", "cr_mult", "cr_luma_mult", "cr_offset",
        "overlap_flag")}
    return (h.kind, d)


def expected_view(kind, grain: Grain, seq: Seq):
    if kind != I.UPDATE_GRAIN:
        return (kind, None)
    return (kind, grain.expected(seq.monochrome, seq.subsampling()))


def check_against_aggregate(parser, views, fps):
    segs = parser.aggregate_grain_headers(*fps)
    want = py_aggregate(views, fps)
    assert len(segs) == len(want)
    for s, w in zip(segs, want):
        assert (s.start_time, s.end_time) == (w["start"], w["end"])
        for k, v in w["p"].items():
            assert getattr(s, k) == v, k
    return segs


GA = Grain(seed=111)
GA2 = Grain(seed=999)                      # same parameters, another seed: must extend the segment
GB = Grain(seed=222, points_y=((0, 30), (255, 90)), ar_coeff_lag=3, scaling_shift=11, ar_coeff_shift=9,
           clip_to_restricted_range=True)


def basic_stream(seq: Seq):
    """key(A) copy A' B disable B -- one Frame OBU per packet, temporal delimiters, a padding and a metadata OBU."""
    frames = [
        (Frame(frame_type=0, grain=GA), I.UPDATE_GRAIN, GA),
        (Frame(frame_type=1, order_hint=1, grain=Grain(kind="copy", seed=5, ref_idx=3)), I.COPY_REF_FRAME, None),
        (Frame(frame_type=1, order_hint=2, grain=GA2, tile_cols_log2=1), I.UPDATE_GRAIN, GA2),
        (Frame(frame_type=1, order_hint=3, grain=GB, segmentation=True, primary_ref_frame=2, gm_translation_on_last=True),
         I.UPDATE_GRAIN, GB),
        (Frame(frame_type=1, order_hint=4, grain=Grain(kind="disable")), I.DISABLE, None),
        (Frame(frame_type=2, order_hint=5, grain=GB, refresh_frame_flags=0x0F), I.UPDATE_GRAIN, GB),
        (Frame(frame_type=1, order_hint=6, error_resilient=True, grain=GB), I.UPDATE_GRAIN, GB),
    ]
    packets = []
    for i, (f, _, _) in enumerate(frames):
        p = W.temporal_delimiter()
        if i == 0:
            p += seq.obu() + W.obu(W.OBU_METADATA, b"\x01\x02\x03")
        p += f.frame_obu(seq)
        if i == 2:
            p += W.obu(W.OBU_PADDING, b"\x00" * 5)
        packets.append(p)
    return packets, [expected_view(k, g, seq) for _, k, g in frames]


@pytest.mark.parametrize("seq", [
    Seq(), Seq(monochrome=True), Seq(profile=1, bit_depth=10, width=1920, height=1080),
    Seq(profile=2, bit_depth=12, ss=(1, 0), width=3840, height=2160, timing_info=True, decoder_model=True, use_128=True,
        frame_id_numbers=True, enable_warped_motion=True, enable_ref_frame_mvs=True, separate_uv_delta_q=True),
    Seq(order_hint_bits=0, enable_cdef=False, enable_restoration=False, choose_screen_content_tools=False, width=64, height=48),
], ids=["420_8bit", "mono", "444_10bit_1080p", "422_12bit_4k_decoder_model", "no_order_hint_small"])
def test_stream_headers_and_segments(seq):
    packets, views = basic_stream(seq)
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    got = [header_view(h) for h in p.get_grain_headers()]
    assert got == views
    info = p.stream_info()
    assert info["bit_depth"] == seq.bit_depth and info["max_frame_width"] == seq.width
    assert (info["ss_x"], info["ss_y"]) == seq.subsampling() and info["packets"] == len(packets)
    segs = check_against_aggregate(p, views, (24000, 1001))
    # A, copy, A' -> one segment of three packets carrying the FIRST seed; B; (disable ends the run); B again over two
    assert [s.random_seed for s in segs] == [111, 222, 222]
    t = 1001 / 24000 * 1e7
    assert segs[0].start_time == 0 and segs[0].end_time == math.ceil(t + t + t)
    assert segs[2].start_time == segs[1].end_time + (math.ceil(t * 5) - math.ceil(t * 4))


def test_table_text_of_an_inspected_stream(tmp_path):
    seq = Seq()
    packets, views = basic_stream(seq)
    path = tmp_path / "clip.ivf"
    path.write_bytes(W.ivf(packets, seq.width, seq.height, 30, 1))
    out = tmp_path / "out.tbl"
    from grav1synth_b200.__main__ import main
    assert main(["inspect", str(path), "-o", str(out)]) == 0
    text = out.read_text().splitlines()
    assert text[0] == "filmgrn1"
    t = 1e7 / 30
    assert text[1] == f"E 0 {math.ceil(t * 3)} 1 111 1"
    assert text[2] == "\tp 2 7 1 10 0 1 128 192 256 120 180 300"
    assert text[3] == "\tsY 3  0 20 128 40 255 60" and text[4] == "\tsCb 2 0 10 255 30" and text[5] == "\tsCr 1 64 12"
    y, cb, cr = GA.coeffs(False, (1, 1))
    assert text[6] == "\tcY " + " ".join(map(str, y)) and text[7] == "\tcCb " + " ".join(map(str, cb))
    assert len(y) == 12 and len(cb) == 13
    assert sum(1 for ln in text if ln.startswith("E ")) == 3
    # the same through the API with an explicit frame rate, and from a Section-5 .obu file
    obu_path = tmp_path / "clip.obu"
    obu_path.write_bytes(b"".join(packets))
    p = I.BitstreamParser()
    assert p.push_file(str(obu_path), (30, 1)) == (30, 1)
    assert format_grain_table(p.aggregate_grain_headers(30, 1)) == out.read_text()


def test_no_grain_reports_nothing(tmp_path, caplog):
    seq = Seq(film_grain_params_present=False)
    packets = [W.temporal_delimiter() + seq.obu() + Frame(frame_type=0).frame_obu(seq),
               W.temporal_delimiter() + Frame(frame_type=1, order_hint=1).frame_obu(seq)]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    assert [h.kind for h in p.get_grain_headers()] == [I.DISABLE, I.DISABLE]
    assert p.aggregate_grain_headers(24, 1) == []
    path = tmp_path / "clip.ivf"
    path.write_bytes(W.ivf(packets, seq.width, seq.height))
    out = tmp_path / "none.tbl"
    from grav1synth_b200.__main__ import main
    assert main(["inspect", str(path), "-o", str(out)]) == 0 and not out.exists()   # src/main.rs:177-183


def test_hidden_frame_then_show_existing_frame():
    """An alt-ref style temporal unit: hidden frame + shown frame in one packet, later a show_existing_frame
    header.  Only shown headers are collected; show_existing_frame counts as CopyRefFrame (frame.rs:196-224)."""
    seq = Seq()
    hidden = Frame(frame_type=1, order_hint=8, show_frame=False, showable_frame=True, refresh_frame_flags=0x40, grain=GB)
    packets = [
        W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, grain=GA).frame_obu(seq),
        W.temporal_delimiter() + hidden.frame_obu(seq) + Frame(frame_type=1, order_hint=1, grain=GA2).frame_obu(seq),
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=2, grain=Grain(kind="copy")).frame_obu(seq),
        W.temporal_delimiter() + Frame(show_existing_frame=6).frame_header_obu(seq),
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=9, grain=GB).frame_obu(seq),
    ]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    views = [expected_view(I.UPDATE_GRAIN, GA, seq), expected_view(I.UPDATE_GRAIN, GA2, seq), (I.COPY_REF_FRAME, None),
             (I.COPY_REF_FRAME, None), expected_view(I.UPDATE_GRAIN, GB, seq)]
    assert [header_view(h) for h in p.get_grain_headers()] == views
    segs = check_against_aggregate(p, views, (25, 1))
    assert len(segs) == 2 and segs[0].end_time == 4 * 400000 and segs[1].start_time == 4 * 400000


def test_split_frame_header_and_tile_group_obus():
    """OBU_FRAME_HEADER + OBU_TILE_GROUP (two tiles) + a redundant frame header: the reference aborts on a standalone
    tile group (obu.rs:215-219); here the frame ends with its last tile and the next header is a new frame."""
    seq = Seq()
    f0 = Frame(frame_type=0, grain=GA, tile_cols_log2=1)
    f1 = Frame(frame_type=1, order_hint=1, grain=GB, tile_cols_log2=1)
    packets = [
        W.temporal_delimiter() + seq.obu() + f0.frame_header_obu(seq) + f0.frame_header_obu(seq)[:0] +
        W.obu(W.OBU_REDUNDANT_FRAME_HEADER, f0.header_bits(seq).push_bool(True).to_bytes()) + f0.tile_group_obu(seq),
        f1.frame_header_obu(seq) + f1.tile_group_obu(seq),          # no temporal delimiter in between
    ]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    assert [header_view(h) for h in p.get_grain_headers()] == [expected_view(I.UPDATE_GRAIN, GA, seq),
                                                               expected_view(I.UPDATE_GRAIN, GB, seq)]


def test_operating_point_layer_filter_and_missing_size_field():
    """OBUs outside operating point 0 are dropped (obu.rs:92-116); the last OBU of a packet may omit its size."""
    seq = Seq(operating_point_idc=(0x101, 0x303))   # op 0: temporal layer 0, spatial layer 0
    base = Frame(frame_type=0, grain=GA)
    enh = Frame(frame_type=1, order_hint=1, grain=GB)
    packets = [
        W.temporal_delimiter() + seq.obu() + base.frame_obu(seq, ext=(0, 0)) + enh.frame_obu(seq, ext=(1, 0)),
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=2, grain=GA2).frame_obu(seq, ext=(0, 0), has_size=False),
    ]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    assert [h.params.random_seed for h in p.get_grain_headers()] == [111, 999]


def test_skip_mode_bit_is_tracked_across_frames():
    """reference_select = 1 with one forward and one backward reference makes skip_mode_present appear
    (frame.rs:1683-1755); the parser has to follow refresh_frame_flags / order hints to know."""
    seq = Seq()
    packets = [
        W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, order_hint=0, grain=GA).frame_obu(seq),
        # hidden future frame stored in slot 7 (order hint 8)
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=8, show_frame=False, refresh_frame_flags=0x80,
                                       grain=Grain(kind="disable")).frame_obu(seq) +
        # frame 4 references slot 0 (hint 0, forward) and slot 7 (hint 8, backward) -> skip mode allowed
        Frame(frame_type=1, order_hint=4, refresh_frame_flags=0x02, ref_frame_idx=(0, 0, 0, 0, 7, 7, 7),
              reference_select=True, skip_mode_bit=True, grain=GB).frame_obu(seq),
        # frame 5: every reference is slot 0 (hint 0) -> one forward reference only, no skip mode bit
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=5, refresh_frame_flags=0x04, reference_select=True,
                                       grain=GA2).frame_obu(seq),
    ]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    assert [h.params.random_seed for h in p.get_grain_headers()] == [111, 222, 999]


def test_frame_refs_short_signaling_derives_the_references():
    """With short signalling only LAST and GOLDEN are coded; AV1 7.8 derives the rest from the slots' order hints, and
    those decide whether skip_mode_present exists.  (The reference stubs the process out: frame.rs:941.)"""
    seq = Seq()
    def stream(short, skip_bit):
        return [
            W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, order_hint=0, grain=GA).frame_obu(seq),
            # slot 7 <- a hidden future frame (order hint 8); slot 1 <- frame 2
            W.temporal_delimiter() + Frame(frame_type=1, order_hint=8, show_frame=False, refresh_frame_flags=0x80,
                                           grain=Grain(kind="disable")).frame_obu(seq) +
            Frame(frame_type=1, order_hint=2, refresh_frame_flags=0x02, grain=GA2).frame_obu(seq),
            # frame 4: LAST = slot 1 (hint 2), GOLDEN = slot 0 (hint 0); the derivation must find slot 7 (hint 8) as the
            # backward reference, which allows skip mode
            W.temporal_delimiter() + Frame(frame_type=1, order_hint=4, refresh_frame_flags=0x04, short_signaling=short,
                                           reference_select=True, skip_mode_bit=skip_bit, grain=GB).frame_obu(seq),
        ]
    p = I.BitstreamParser()
    for pk in stream((1, 0), True):
        p.push_packet(pk)
    assert [h.params.random_seed for h in p.get_grain_headers()] == [111, 999, 222]
    assert header_view(p.get_grain_headers()[2]) == expected_view(I.UPDATE_GRAIN, GB, seq)
    # without the skip_mode_present bit the same header no longer parses to GB
    q = I.BitstreamParser()
    try:
        for pk in stream((1, 0), None):
            q.push_packet(pk)
        hs = q.get_grain_headers()
        assert len(hs) < 3 or header_view(hs[2]) != expected_view(I.UPDATE_GRAIN, GB, seq)
    except G1SError:
        pass


def test_inherited_segmentation_decides_losslessness():
    """base_q_idx = 0 everywhere, but segment 1 carries ALT_Q = +20: the key frame is not lossless, and neither is the
    inter frame that keeps the data of its primary reference (segmentation enabled, nothing re-sent) -- so its
    loop-filter / CDEF / restoration / tx-mode syntax is present (spec 7.20 load_previous)."""
    seq = Seq(monochrome=True)
    packets = [
        W.temporal_delimiter() + seq.obu() +
        Frame(frame_type=0, base_q_idx=0, segmentation=True, seg_value=20, lossless=False, grain=GA).frame_obu(seq),
        W.temporal_delimiter() +
        Frame(frame_type=1, order_hint=1, base_q_idx=0, primary_ref_frame=0, segmentation=True, seg_inherit=True,
              lossless=False, grain=GB).frame_obu(seq),
        # and a frame that switches segmentation off is lossless again
        W.temporal_delimiter() +
        Frame(frame_type=1, order_hint=2, base_q_idx=0, primary_ref_frame=0, lossless=True, grain=GA2).frame_obu(seq),
    ]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    assert [header_view(h) for h in p.get_grain_headers()] == [
        expected_view(I.UPDATE_GRAIN, GA, seq), expected_view(I.UPDATE_GRAIN, GB, seq), expected_view(I.UPDATE_GRAIN, GA2, seq)]


def test_malformed_streams_are_reported():
    seq = Seq()
    good = W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, grain=GA).frame_obu(seq)
    for bad in (good[:-9], bytes([0x80]) + good, W.temporal_delimiter() + Frame(frame_type=0).frame_obu(seq)):
        p = I.BitstreamParser()
        with pytest.raises(G1SError) as e:
            p.push_packet(bad)
        assert e.value.code == abi.G1S_E_STREAM
    p = I.BitstreamParser()
    with pytest.raises(G1SError):
        p.push_file("/nonexistent/clip.ivf")


def test_inspect_symbols_exported():
    import ctypes
    from grav1synth_b200.diff import LIB_PATH
    L = ctypes.CDLL(LIB_PATH)
    for name in I.EXPORTS:
        assert hasattr(L, name), name


def test_frame_level_force_integer_mv_gates_allow_high_precision_mv():
    """spec 5.9.2: under SELECT_INTEGER_MV the frame codes its own force_integer_mv, and allow_high_precision_mv is
    absent when that bit is 1.  (The reference reads the bit and then tests the sequence-level value, frame.rs:469, so
    every later field of such an inter frame is off by one bit.)"""
    seq = Seq()
    for fim in (False, True):
        packets = [
            W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, order_hint=0, grain=GA).frame_obu(seq),
            W.temporal_delimiter() + Frame(frame_type=1, order_hint=1, allow_screen_content_tools=True,
                                           force_integer_mv=fim, grain=GB).frame_obu(seq),
        ]
        p = I.BitstreamParser()
        for pk in packets:
            p.push_packet(pk)
        hs = p.get_grain_headers()
        assert len(hs) == 2 and header_view(hs[1]) == expected_view(I.UPDATE_GRAIN, GB, seq), fim


def test_delta_frame_ids_are_coded_with_short_signaling_too():
    """spec 5.9.2: delta_frame_id_minus_1 follows every reference when frame ids are present, whether ref_frame_idx[i]
    itself was coded or derived by frame_refs_short_signaling."""
    seq = Seq(frame_id_numbers=True)
    for short in (None, (0, 0)):
        packets = [
            W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, order_hint=0, grain=GA).frame_obu(seq),
            W.temporal_delimiter() + Frame(frame_type=1, order_hint=1, short_signaling=short, grain=GB).frame_obu(seq),
        ]
        p = I.BitstreamParser()
        for pk in packets:
            p.push_packet(pk)
        hs = p.get_grain_headers()
        assert len(hs) == 2 and header_view(hs[1]) == expected_view(I.UPDATE_GRAIN, GB, seq), short


def test_temporal_point_info_in_shown_and_show_existing_headers():
    """Decoder model without equal_picture_interval: frame_presentation_time is coded in every shown frame header AND
    in show_existing_frame headers (the reference skips the latter and then fails its alignment check)."""
    seq = Seq(timing_info=True, decoder_model=True, equal_picture_interval=False)
    hidden = Frame(frame_type=1, order_hint=2, show_frame=False, showable_frame=True, refresh_frame_flags=0x02, grain=GB)
    packets = [
        W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, order_hint=0, grain=GA).frame_obu(seq),
        W.temporal_delimiter() + hidden.frame_obu(seq) + Frame(frame_type=1, order_hint=1, grain=GA2).frame_obu(seq),
        W.temporal_delimiter() + Frame(show_existing_frame=1).frame_header_obu(seq),
    ]
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    hs = p.get_grain_headers()
    assert [h.kind for h in hs] == [I.UPDATE_GRAIN, I.UPDATE_GRAIN, I.COPY_REF_FRAME]


def test_aggregation_compares_every_coefficient_the_count_names():
    """same_grain (grain.rs:83-105 equality) looks at exactly the coefficients the count byte names (count + 1 is
    stored): headers that differ only in the LAST chroma coefficient (the luma-correlation tap, 25th at lag 3) are two
    segments, headers that differ only in the seed are one."""
    seq = Seq()
    base = dict(points_y=((0, 30), (255, 90)), ar_coeff_lag=3, scaling_shift=11, ar_coeff_shift=9)
    cr_a = [((i * 3) % 23) - 11 for i in range(25)]
    cr_b = cr_a[:24] + [cr_a[24] + 1]

    def table(g1, g2):
        p = I.BitstreamParser()
        p.push_packet(W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, order_hint=0, grain=g1).frame_obu(seq))
        p.push_packet(W.temporal_delimiter() + Frame(frame_type=1, order_hint=1, grain=g2).frame_obu(seq))
        return p.aggregate_grain_headers(24, 1)

    assert len(table(Grain(seed=5, ar_coeffs_cr=cr_a, **base), Grain(seed=6, ar_coeffs_cr=cr_a, **base))) == 1
    assert len(table(Grain(seed=5, ar_coeffs_cr=cr_a, **base), Grain(seed=6, ar_coeffs_cr=cr_b, **base))) == 2
