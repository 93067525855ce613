"""`inspect` on REAL AV1 streams: libaom 3.13.1's own encoder (oracle/aom_encode.py, the libaom bundled with
opencv-python-headless) produces the bitstreams, csrc/g1s_obu.cpp walks them.  CPU only.

 * film-grain-test=K: the encoder signals libaom's built-in film grain test vector K on every frame; every
   UpdateGrain header we parse must equal the vector as it sits in the binary's .rodata (all 16 vectors);
 * film-grain-table=FILE: the encoder reads a `filmgrn1` table -- here the golden tables our `diff` path writes --
   and `inspect` must give the same parameters back: diff -> table -> libaom encoder -> AV1 -> inspect, the
   grav1synth workflow end to end, and proof that libaom's table reader accepts what g1s_write_grain_table emits;
 * encoder settings that change the frame-header syntax around the grain parameters (low delay, alt-refs with
   show_existing_frame, tiles and several tile groups per frame, no CDEF / restoration, error resilience, no order hints,
   delta q / lf, quantiser matrices, lossless, super-resolution, spatial resize with found_ref, screen content with
   intra block copy, 1080p and 8K).
Three of the streams are committed under tests/golden/aom/*.ivf so the walk is also checked without libaom.
"""
import json
import os

import pytest

from helpers import ROOT
from grav1synth_b200 import inspect as I
from grav1synth_b200.diff import format_grain_table
from oracle import aom_pin

ok, why = aom_pin.available()
needs_libaom = pytest.mark.skipif(not ok, reason="libaom pin unavailable: " + str(why))
GOLD = os.path.join(ROOT, "tests", "golden", "aom")
FIELDS = ("scaling_points_y", "scaling_points_cb", "scaling_points_cr", "scaling_shift", "ar_coeff_lag",
          "ar_coeff_shift", "cb_mult", "cb_luma_mult", "cb_offset", "cr_mult", "cr_luma_mult", "cr_offset",
          "overlap_flag", "chroma_scaling_from_luma", "grain_scale_shift")


def vector_view(v):
    """What a parsed header must hold for libaom's aom_film_grain_t `v` (4:2:0 stream)."""
    ny, ncb, ncr = len(v["scaling_points_y"]), len(v["scaling_points_cb"]), len(v["scaling_points_cr"])
    if v["chroma_scaling_from_luma"] or ny == 0:
        ncb = ncr = 0  # not coded (AV1 5.9.30)
    npl = 2 * v["ar_coeff_lag"] * (v["ar_coeff_lag"] + 1)
    npc = npl + 1 if ny else npl
    d = {k: v[k] for k in FIELDS}
    d["scaling_points_cb"], d["scaling_points_cr"] = v["scaling_points_cb"][:ncb], v["scaling_points_cr"][:ncr]
    d["overlap_flag"], d["chroma_scaling_from_luma"] = bool(v["overlap_flag"]), bool(v["chroma_scaling_from_luma"])
    d["ar_coeffs_y"] = v["ar_coeffs_y"][:npl] if ny else []
    d["ar_coeffs_cb"] = v["ar_coeffs_cb"][:npc] if (v["chroma_scaling_from_luma"] or ncb) else [0]
    d["ar_coeffs_cr"] = v["ar_coeffs_cr"][:npc] if (v["chroma_scaling_from_luma"] or ncr) else [0]
    for c, n in (("cb", ncb), ("cr", ncr)):
        if not n:
            d[c + "_mult"] = d[c + "_luma_mult"] = d[c + "_offset"] = 0
    return d


def header_view(h):
    g = h.params
    d = {k: getattr(g, k) for k in FIELDS}
    d["scaling_points_y"], d["scaling_points_cb"], d["scaling_points_cr"] = (
        [tuple(p) for p in g.scaling_points_y], [tuple(p) for p in g.scaling_points_cb],
        [tuple(p) for p in g.scaling_points_cr])
    d["ar_coeffs_y"], d["ar_coeffs_cb"], d["ar_coeffs_cr"] = g.ar_coeffs_y, g.ar_coeffs_cb, g.ar_coeffs_cr
    return d


def inspect_packets(packets):
    p = I.BitstreamParser()
    for pk in packets:
        p.push_packet(pk)
    return p, p.get_grain_headers()


@needs_libaom
@pytest.mark.parametrize("k", range(1, 17))
def test_every_libaom_film_grain_test_vector_survives_encode_and_inspect(k):
    from oracle import aom_encode as E
    want = vector_view(E.test_vector(k))
    packets = E.encode(E.synthetic_frames(7, 176, 144, seed=k), 176, 144, {"film-grain-test": str(k)})
    _, hs = inspect_packets(packets)
    assert len(hs) == 7 and hs[0].kind == I.UPDATE_GRAIN           # one header per displayed frame
    ups = [h for h in hs if h.kind == I.UPDATE_GRAIN]
    assert all(h.kind != I.DISABLE for h in hs)
    if E.test_vector(k)["update_parameters"]:
        assert len(ups) >= 4
    else:  # the vector asks for update_grain = 0: inter frames load the parameters of a reference frame
        assert sum(h.kind == I.COPY_REF_FRAME for h in hs) >= 4
    for h in ups:
        assert header_view(h) == want
        assert h.clip_to_restricted_range == bool(E.test_vector(k)["clip_to_restricted_range"])


ENCODER_VARIANTS = {
    "low_delay": dict(lag=0, opts={}),
    "altref_lag19": dict(lag=19, opts={"auto-alt-ref": "1"}),
    "two_tile_columns_4_rows": dict(lag=None, opts={"tile-columns": "1", "tile-rows": "2"}, size=(704, 576)),
    "no_cdef_no_restoration": dict(lag=None, opts={"enable-cdef": "0", "enable-restoration": "0"}),
    "error_resilient": dict(lag=0, opts={}, cfg={12: 1}),                      # g_error_resilient
    # hidden alt-refs under error resilience: every inter frame re-signals ref_order_hint[], which must REPLACE the saved
    # hints (the reference keeps stale ones, frame.rs:355-362, and then mis-reads skip_mode_present) -- found by fuzzing
    "error_resilient_altref": dict(lag=19, opts={}, cfg={12: 1}, frames=12),
    # 640 wide resized to 427: 7 superblock columns in 2 tiles (4 + 3), two tile groups with libaom's redundant frame
    # header between them -- the tile count must be the spec's ceil, not the reference's floor (found by fuzzing)
    "resized_uneven_tiles_redundant_headers": dict(lag=0, opts={"tile-columns": "1", "num-tile-groups": "2"},
                                                   cfg={12: 1, 16: 1, 17: 12, 18: 12}, size=(640, 360), frames=6),
    "sb128_no_global_motion": dict(lag=None, opts={"sb-size": "128", "enable-global-motion": "0"}),
    "screen_content_palette": dict(lag=None, opts={"tune-content": "screen"}),
    "full_hd": dict(lag=None, opts={}, size=(1920, 1080), frames=3),
    "superres_fixed": dict(lag=None, opts={}, cfg={19: 1, 20: 12, 21: 12}),     # rc_superres_mode FIXED, denominators
    "superres_random": dict(lag=None, opts={}, cfg={19: 2}, frames=12),
    "superres_screen_intrabc": dict(lag=None, opts={"tune-content": "screen", "enable-intrabc": "1"},
                                    cfg={19: 1, 20: 12, 21: 12}),
    "superres_lossless": dict(lag=None, opts={"lossless": "1"}, cfg={19: 1, 20: 12, 21: 12}),
    "resize_fixed_tiles_hd": dict(lag=None, opts={"tile-columns": "2"}, cfg={16: 1, 17: 12, 18: 12},   # rc_resize_mode
                                  size=(1920, 1080), frames=4),
    "resize_dynamic_cbr": dict(lag=0, opts={}, cfg={16: 3, 24: 1}, frames=12),
    "two_tile_groups": dict(lag=None, opts={"tile-columns": "1", "num-tile-groups": "2"}, size=(704, 576)),
    "four_tile_groups": dict(lag=None, opts={"tile-columns": "1", "tile-rows": "1", "num-tile-groups": "4"},
                             size=(704, 576)),
    "no_order_hint": dict(lag=None, opts={"enable-order-hint": "0"}),
    "delta_q_and_lf": dict(lag=None, opts={"deltaq-mode": "1", "delta-lf-mode": "1"}),
    "quant_matrices": dict(lag=None, opts={"enable-qm": "1"}),
    "lossless": dict(lag=None, opts={"lossless": "1"}),
    "cdf_update_off_reduced_tx": dict(lag=None, opts={"cdf-update-mode": "0", "reduced-tx-type-set": "1"}),
    "uhd_8k_one_frame": dict(lag=None, opts={}, size=(7680, 4320), frames=1),
    "forward_keyframes": dict(lag=19, opts={}, cfg={45: 1, 47: 8, 48: 8}, frames=24),      # fwd_kf_enabled, kf_min/max_dist
    "keyframe_every_5": dict(lag=19, opts={}, cfg={47: 5, 48: 5}, frames=24),
    "s_frames": dict(lag=0, opts={}, cfg={49: 4, 50: 1, 12: 1}, frames=24),                 # sframe_dist / mode, error resilient
    "monochrome": dict(lag=None, opts={}, cfg={52: 1}, frames=6, mono=True),
    "reduced_still_picture": dict(lag=None, opts={}, cfg={5: 1, 53: 0}, frames=1),          # g_limit = 1
    "full_header_still_picture": dict(lag=None, opts={}, cfg={5: 1, 53: 1}, frames=1),
    "main_profile_10bit": dict(lag=None, opts={}, enc=dict(bit_depth=10), frames=6),
    "high_profile_444": dict(lag=None, opts={}, enc=dict(chroma444=True), frames=6),
    "high_profile_444_10bit": dict(lag=0, opts={}, enc=dict(chroma444=True, bit_depth=10), frames=6),
}


@needs_libaom
@pytest.mark.parametrize("name", list(ENCODER_VARIANTS))
def test_encoder_variants_change_the_header_syntax_not_the_grain(name):
    from oracle import aom_encode as E
    v = ENCODER_VARIANTS[name]
    w, h = v.get("size", (352, 288))
    n = v.get("frames", 10)
    opts = {"film-grain-test": "1"}   # vector 1 updates its parameters on every frame
    opts.update(v["opts"])
    enc = v.get("enc", {})
    packets = E.encode(E.synthetic_frames(n, w, h, seed=3), w, h, opts, lag_in_frames=v["lag"], cfg_words=v.get("cfg"),
                       **enc)
    p, hs = inspect_packets(packets)
    assert len(hs) == n
    want = vector_view(E.test_vector(1))   # has luma points, so 4:4:4 codes the same syntax as 4:2:0
    if v.get("mono"):                       # no chroma syntax at all
        want.update(scaling_points_cb=[], scaling_points_cr=[], ar_coeffs_cb=[0], ar_coeffs_cr=[0], cb_mult=0,
                    cb_luma_mult=0, cb_offset=0, cr_mult=0, cr_luma_mult=0, cr_offset=0)
        assert p.stream_info()["monochrome"] == 1
    assert all(header_view(h) == want for h in hs if h.kind == I.UPDATE_GRAIN)
    assert hs[0].kind == I.UPDATE_GRAIN and all(h.kind != I.DISABLE for h in hs)
    assert sum(h.kind == I.UPDATE_GRAIN for h in hs) >= (n + 1) // 2
    info = p.stream_info()
    assert (info["max_frame_width"], info["max_frame_height"], info["bit_depth"]) == (w, h, enc.get("bit_depth", 8))
    assert (info["ss_x"], info["ss_y"]) == ((0, 0) if enc.get("chroma444") else (1, 1))
    segs = p.aggregate_grain_headers(24, 1)
    assert len(segs) == 1 and segs[0].start_time == 0 and segs[0].end_time == -(-n * 10_000_000 // 24)


@needs_libaom
@pytest.mark.parametrize("table", ["c2_small_8bit", "c3_small_10bit", "heavy_grain_12bit", "yuv444_8bit"])
def test_diff_table_through_libaom_and_back(table, tmp_path):
    """diff -> filmgrn1 table -> libaom (--film-grain-table) -> AV1 -> inspect gives the table back."""
    from oracle import aom_encode as E
    from grav1synth_b200.grain_table import parse_grain_table
    src = os.path.join(ROOT, "tests", "golden", table + ".tbl")
    segs = parse_grain_table(open(src).read())
    packets = E.encode(E.synthetic_frames(6, 352, 288, seed=9), 352, 288, {"film-grain-table": src}, lag_in_frames=0)
    p, hs = inspect_packets(packets)
    assert len(hs) == 6 and all(h.kind == I.UPDATE_GRAIN for h in hs)
    used = []
    for k, h in enumerate(hs):
        t = k * 10_000_000 // 24  # libaom looks the frame's start time up in the table
        seg = next(s for s in segs if s.start_time <= t < s.end_time)
        if seg not in used:
            used.append(seg)
        g = h.params
        for f in FIELDS + ("ar_coeffs_y", "ar_coeffs_cb", "ar_coeffs_cr"):
            assert getattr(g, f) == getattr(seg, f), (k, f)
    out = p.aggregate_grain_headers(24, 1)
    assert len(out) == len(used)
    # same text as the input table, apart from the E lines (times; libaom reseeds every frame)
    body = lambda text: [ln for ln in text.splitlines()[1:] if not ln.startswith("E ")]
    assert body(format_grain_table(out)) == body(format_grain_table(used))


@pytest.mark.parametrize("name", ["stream_fgtest1_altref", "stream_fgtest9_lowdelay", "stream_table_c2"])
def test_committed_libaom_streams(name, tmp_path):
    """The same walk on streams libaom produced earlier (tests/golden/make_aom_streams.py): runs without libaom."""
    want = json.load(open(os.path.join(GOLD, name + ".json")))
    p = I.BitstreamParser()
    fps = p.push_file(os.path.join(GOLD, name + ".ivf"))
    assert list(fps) == want["fps"]
    hs = p.get_grain_headers()
    assert [h.kind for h in hs] == want["kinds"]
    got = [json.loads(json.dumps(header_view(h))) for h in hs if h.kind == I.UPDATE_GRAIN]
    assert all(g == want["params"] for g in got)
    assert format_grain_table(p.aggregate_grain_headers(*fps)) == want["table"]
