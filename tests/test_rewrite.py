"""`apply` / `remove`: the header rewriter of csrc/g1s_obu.cpp -- CPU tests.

Witnesses: libaom's own encoder and decoder (oracle/aom_encode.py).
 * remove: stripping the film grain from a stream libaom encoded WITH `film-grain-test` must give, byte for byte, the
   stream libaom produces for the same frames WITHOUT film grain;
 * apply: the rewritten stream must still decode with libaom, `inspect` must read the applied table back on every
   frame (seed advancing by DEFAULT_GRAIN_SEED, update_grain on inter frames, apply_grain = 0 outside the table's
   segments), and applying then removing must restore the input.
Plus writer-level checks on hand-built streams that need no libaom.
"""
import os

import numpy as np
import pytest

import av1_writer as W
from av1_writer import Frame, Grain, Seq
from helpers import ROOT
from grav1synth_b200 import inspect as I
from grav1synth_b200.grain_table import parse_grain_table
from oracle import aom_pin
from test_inspect import GA, GB

ok, why = aom_pin.available()
needs_libaom = pytest.mark.skipif(not ok, reason="libaom pin unavailable: " + str(why))
T24 = 10_000_000 / 24


def golden_table(name):
    return parse_grain_table(open(os.path.join(ROOT, "tests", "golden", name + ".tbl")).read())


def ts(k):
    return int(np.ceil(k * T24))


@needs_libaom
@pytest.mark.parametrize("vector,lag", [(5, None), (1, 0), (12, 19), (3, 19), (16, None), (9, 0)])
def test_remove_gives_libaoms_own_grainless_stream(vector, lag):
    from oracle import aom_encode as E
    fr = E.synthetic_frames(10, 176, 144, seed=vector)
    # constant quality (rc_end_usage = AOM_Q): with rate control the bytes spent on grain headers feed back into the
    # quantiser choice and the two encodes legitimately diverge
    q = dict(options={"cq-level": "32"}, cfg_words={24: 3}, lag_in_frames=lag)
    plain = E.encode(fr, 176, 144, **q)
    q["options"] = {"cq-level": "32", "film-grain-test": str(vector)}
    grainy = E.encode(fr, 176, 144, **q)
    assert grainy != plain
    rw = I.GrainRewriter(None)
    assert [rw.rewrite_packet(p, ts(k)) for k, p in enumerate(grainy)] == plain
    # the handle still reports what the INPUT carried
    assert any(h.kind == I.UPDATE_GRAIN for h in rw.get_grain_headers())


@needs_libaom
@pytest.mark.parametrize("name,opts,kw", [
    ("two_tile_groups", {"tile-columns": "1", "num-tile-groups": "2"}, {}),            # OBU_FRAME_HEADER + OBU_TILE_GROUPs
    ("high_profile_444_10bit", {}, dict(chroma444=True, bit_depth=10)),
    ("superres", {}, dict(cfg_words={19: 1, 20: 12, 21: 12})),
    ("error_resilient_low_delay", {}, dict(cfg_words={12: 1}, lag_in_frames=0)),
    # uneven tiles + redundant frame headers between the tile groups: every copy of the header must be rewritten
    ("resized_uneven_tiles_redundant_headers", {"tile-columns": "1", "num-tile-groups": "2"},
     dict(cfg_words={12: 1, 16: 1, 17: 12, 18: 12}, lag_in_frames=0)),
])
def test_remove_and_apply_on_other_stream_structures(name, opts, kw):
    from oracle import aom_encode as E
    fr = E.synthetic_frames(8, 704, 576, seed=5)
    kw = dict(kw)
    cfg = dict(kw.pop("cfg_words", {}))
    cfg[24] = 3   # constant quality, see above
    base = dict({"cq-level": "32"}, **opts)
    plain = E.encode(fr, 704, 576, dict(base), cfg_words=cfg, **kw)
    grainy = E.encode(fr, 704, 576, dict(base, **{"film-grain-test": "3"}), cfg_words=cfg, **kw)
    rm = I.GrainRewriter(None)
    assert [rm.rewrite_packet(p, ts(k)) for k, p in enumerate(grainy)] == plain
    ap = I.GrainRewriter(golden_table("c2_small_8bit"))
    applied = [ap.rewrite_packet(p, ts(k)) for k, p in enumerate(plain)]
    if kw.get("bit_depth", 8) == 8:
        assert len(E.decode(applied)) == 8
    p = I.BitstreamParser()
    for pk in applied:
        p.push_packet(pk)
    hs = p.get_grain_headers()
    assert len(hs) == 8 and all(h.kind in (I.UPDATE_GRAIN, I.COPY_REF_FRAME) for h in hs)
    want = golden_table("c2_small_8bit")[0]
    assert all(h.params.scaling_points_y == want.scaling_points_y and h.params.ar_coeffs_cb == want.ar_coeffs_cb
               for h in hs if h.kind == I.UPDATE_GRAIN)
    rm2 = I.GrainRewriter(None)
    assert [rm2.rewrite_packet(pk, ts(k)) for k, pk in enumerate(applied)] == plain


@needs_libaom
@pytest.mark.parametrize("table", ["c2_small_8bit", "heavy_grain_12bit"])
def test_apply_diff_table_to_a_libaom_stream(table):
    from oracle import aom_encode as E
    segs = golden_table(table)
    fr = E.synthetic_frames(8, 176, 144, seed=4)
    plain = E.encode(fr, 176, 144, {}, lag_in_frames=19)
    rw = I.GrainRewriter(segs)
    out = [rw.rewrite_packet(p, ts(k)) for k, p in enumerate(plain)]
    assert rw.counters()[0] >= 8 and rw.counters()[1] == 0
    # libaom decodes the result, and film grain now changes the pictures
    dec_plain, dec_out = E.decode(plain), E.decode(out)
    assert len(dec_out) == len(dec_plain) == 8
    assert all(not np.array_equal(a[0], b[0]) for a, b in zip(dec_plain, dec_out))
    # inspect reads the table back
    p = I.BitstreamParser()
    for pk in out:
        p.push_packet(pk)
    hs = p.get_grain_headers()
    assert len(hs) == 8
    shown_ts = iter(ts(k) for k in range(8))
    seeds = {}
    for h in hs:
        if h.kind == I.COPY_REF_FRAME:   # show_existing_frame of an already rewritten hidden frame
            next(shown_ts)
            continue
        assert h.kind == I.UPDATE_GRAIN and h.clip_to_restricted_range
        g = h.params
        seg = next(s for s in segs if all(getattr(g, f) == getattr(s, f) for f in (
            "scaling_points_y", "scaling_points_cb", "scaling_points_cr", "ar_coeffs_y", "ar_coeffs_cb", "ar_coeffs_cr",
            "scaling_shift", "ar_coeff_shift", "ar_coeff_lag", "cb_mult", "cb_luma_mult", "cb_offset", "overlap_flag")))
        seeds.setdefault(id(seg), []).append(g.random_seed)
    for seg in segs:
        # frame.rs:637-640: the segment's seed advances by DEFAULT_GRAIN_SEED for every frame it is written to
        # (hidden frames included, so the shown frames may skip steps); strictly increasing step counts
        steps = {(seg.random_seed + n * 10956) % 65536: n for n in range(1, 40)}
        got = [steps[s] for s in seeds.get(id(seg), [])]
        assert got == sorted(got) and len(set(got)) == len(got)
    # apply then remove restores the input exactly
    rm = I.GrainRewriter(None)
    assert [rm.rewrite_packet(pk, ts(k)) for k, pk in enumerate(out)] == plain


@needs_libaom
def test_apply_outside_the_table_disables_grain():
    from oracle import aom_encode as E
    seg = golden_table("c2_small_8bit")[0]
    seg.raw.start_time, seg.raw.end_time = ts(2), ts(5)        # covers packets 2, 3, 4 only
    fr = E.synthetic_frames(7, 176, 144, seed=6)
    plain = E.encode(fr, 176, 144, {}, lag_in_frames=0)
    rw = I.GrainRewriter([seg])
    out = [rw.rewrite_packet(p, ts(k)) for k, p in enumerate(plain)]
    assert len(E.decode(out)) == 7
    p = I.BitstreamParser()
    for pk in out:
        p.push_packet(pk)
    assert [h.kind for h in p.get_grain_headers()] == [I.DISABLE] * 2 + [I.UPDATE_GRAIN] * 3 + [I.DISABLE] * 2
    assert rw.counters() == (3, 4)
    segs = p.aggregate_grain_headers(24, 1)
    assert len(segs) == 1 and (segs[0].start_time, segs[0].end_time) == (ts(2), ts(5))


def test_rewrite_hand_built_stream_roundtrip(tmp_path):
    """No libaom needed: apply on the writer's stream (OBU_FRAME with two tiles, standalone frame header + tile group,
    hidden frame, show_existing_frame, padding, an OBU without size field), read back with inspect, IVF in/out, CLI."""
    seq = Seq()
    split = Frame(frame_type=1, order_hint=3, grain=GB, tile_cols_log2=1)
    packets = [
        W.temporal_delimiter() + seq.obu() + Frame(frame_type=0, grain=GA).frame_obu(seq),
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=8, show_frame=False, refresh_frame_flags=0x40,
                                       grain=GB).frame_obu(seq) +
        Frame(frame_type=1, order_hint=1, grain=Grain(kind="copy"), tile_cols_log2=1).frame_obu(seq),
        # frame header, the header repeated (as OBU_FRAME_HEADER and as a redundant copy), the tile group, padding
        W.temporal_delimiter() + split.frame_header_obu(seq) + split.frame_header_obu(seq) +
        W.obu(W.OBU_REDUNDANT_FRAME_HEADER, split.header_bits(seq).push_bool(True).to_bytes()) +
        split.tile_group_obu(seq) + W.obu(W.OBU_PADDING, b"\0\0"),
        W.temporal_delimiter() + Frame(show_existing_frame=6).frame_header_obu(seq),
        W.temporal_delimiter() + Frame(frame_type=1, order_hint=9, grain=Grain(kind="disable")).frame_obu(seq, has_size=False),
    ]
    table = golden_table("c3_small_10bit")
    rw = I.GrainRewriter(table)
    out = [rw.rewrite_packet(p, ts(k)) for k, p in enumerate(packets)]
    assert rw.counters() == (5, 0)                               # 4 shown frames with a header + the hidden one
    p = I.BitstreamParser()
    for pk in out:
        p.push_packet(pk)
    hs = p.get_grain_headers()
    assert [h.kind for h in hs] == [I.UPDATE_GRAIN, I.UPDATE_GRAIN, I.UPDATE_GRAIN, I.COPY_REF_FRAME, I.UPDATE_GRAIN]
    want = table[0]
    for h in hs:
        if h.kind == I.UPDATE_GRAIN:
            assert h.params.scaling_points_y == want.scaling_points_y and h.params.ar_coeffs_cr == want.ar_coeffs_cr
    # the repeated headers inside packet 2 were rewritten to the same bytes as the first copy
    def obus(pk):
        out, off = [], 0
        while off < len(pk):
            t, has_size = (pk[off] >> 3) & 15, (pk[off] >> 1) & 1
            n, k, shift = 0, off + 1, 0
            while True:
                n |= (pk[k] & 0x7F) << shift
                shift += 7
                k += 1
                if not pk[k - 1] & 0x80:
                    break
            out.append((t, pk[k:k + n]))
            off = k + n
        return out
    copies = [body for t, body in obus(out[2]) if t in (W.OBU_FRAME_HEADER, W.OBU_REDUNDANT_FRAME_HEADER)]
    assert len(copies) == 3 and copies[0] == copies[1] == copies[2]
    assert copies[0] != [body for t, body in obus(packets[2]) if t == W.OBU_FRAME_HEADER][0]
    seeds = [h.params.random_seed for h in hs if h.kind == I.UPDATE_GRAIN]
    step = lambda n: (want.random_seed + n * 10956) % 65536
    assert seeds == [step(1), step(3), step(4), step(5)]          # step(2) went to the hidden frame
    # tile data and the other OBUs are untouched: removing again restores the input, except that the input's own
    # grain syntax is gone, which is exactly `remove` applied to the input
    rm1, rm2 = I.GrainRewriter(None), I.GrainRewriter(None)
    assert [rm1.rewrite_packet(pk, ts(k)) for k, pk in enumerate(out)] == \
           [rm2.rewrite_packet(pk, ts(k)) for k, pk in enumerate(packets)]
    # CLI: apply and remove on IVF files
    from grav1synth_b200.__main__ import main
    src, tbl, applied, removed = (tmp_path / n for n in ("in.ivf", "t.tbl", "applied.ivf", "removed.ivf"))
    src.write_bytes(W.ivf(packets, seq.width, seq.height, 24, 1))
    tbl.write_text(open(os.path.join(ROOT, "tests", "golden", "c3_small_10bit.tbl")).read())
    assert main(["apply", str(src), "-o", str(applied), "-g", str(tbl)]) == 0
    assert main(["remove", str(applied), "-o", str(removed)]) == 0
    q = I.BitstreamParser()
    q.push_file(str(applied))
    assert [h.kind for h in q.get_grain_headers()] == [h.kind for h in hs]
    r = I.BitstreamParser()
    r.push_file(str(removed))
    assert all(h.kind == I.DISABLE or h.kind == I.COPY_REF_FRAME for h in r.get_grain_headers())
    assert r.stream_info()["film_grain_params_present"] == 0


@needs_libaom
def test_random_tables_survive_apply_libaom_decode_and_inspect():
    """30 random (valid) film grain parameter sets: every branch of the grain syntax writer (no luma points, chroma
    scaling from luma, 0..3 lags, with / without chroma points and multipliers) must give a stream libaom decodes and
    `inspect` reads back exactly."""
    from oracle import aom_encode as E
    from grav1synth_b200.abi import CSegment, GrainTableSegment
    rng = np.random.default_rng(77)
    fr = E.synthetic_frames(3, 176, 144, seed=8)
    plain = E.encode(fr, 176, 144, {}, lag_in_frames=0)

    def points(n):
        xs = np.sort(rng.choice(256, n, replace=False))
        return [(int(x), int(rng.integers(0, 256))) for x in xs]

    for trial in range(30):
        s = CSegment()
        s.start_time, s.end_time, s.random_seed = 0, 2 ** 63, int(rng.integers(0, 65536))
        ny = int(rng.choice([0, 1, 2, 7, 14]))
        csfl = bool(rng.integers(0, 2)) and ny > 0
        # 4:2:0: grain goes to both chroma planes or to neither (libaom enforces the spec's constraint)
        ncb, ncr = (0, 0) if (csfl or ny == 0 or rng.integers(0, 3) == 0) else (int(rng.integers(1, 11)), int(rng.integers(1, 11)))
        for dst, n, cnt in ((s.scaling_points_y, ny, "num_y_points"), (s.scaling_points_cb, ncb, "num_cb_points"),
                            (s.scaling_points_cr, ncr, "num_cr_points")):
            setattr(s, cnt, n)
            for k, (x, y) in enumerate(points(n)):
                dst[k][0], dst[k][1] = x, y
        s.chroma_scaling_from_luma = int(csfl)
        s.scaling_shift, s.ar_coeff_lag = int(rng.integers(8, 12)), int(rng.integers(0, 4))
        s.ar_coeff_shift, s.grain_scale_shift = int(rng.integers(6, 10)), int(rng.integers(0, 4))
        for arr in (s.ar_coeffs_y, s.ar_coeffs_cb, s.ar_coeffs_cr):
            for k in range(len(arr)):
                arr[k] = int(rng.integers(-128, 128))
        s.cb_mult, s.cb_luma_mult, s.cb_offset = int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 512))
        s.cr_mult, s.cr_luma_mult, s.cr_offset = int(rng.integers(0, 256)), int(rng.integers(0, 256)), int(rng.integers(0, 512))
        s.overlap_flag = int(rng.integers(0, 2))
        seg = GrainTableSegment.from_c(s)
        rw = I.GrainRewriter([seg])
        out = [rw.rewrite_packet(p, ts(k)) for k, p in enumerate(plain)]
        assert len(E.decode(out)) == 3, trial
        p = I.BitstreamParser()
        for pk in out:
            p.push_packet(pk)
        hs = p.get_grain_headers()
        assert [h.kind for h in hs] == [I.UPDATE_GRAIN] * 3
        npl = 2 * s.ar_coeff_lag * (s.ar_coeff_lag + 1)
        npc = npl + 1 if ny else npl
        for h in hs:
            g = h.params
            assert g.scaling_points_y == seg.scaling_points_y and g.chroma_scaling_from_luma == csfl
            assert g.scaling_points_cb == seg.scaling_points_cb[:ncb] and g.scaling_points_cr == seg.scaling_points_cr[:ncr]
            assert (g.scaling_shift, g.ar_coeff_lag, g.ar_coeff_shift, g.grain_scale_shift, g.overlap_flag) == \
                   (seg.scaling_shift, seg.ar_coeff_lag, seg.ar_coeff_shift, seg.grain_scale_shift, seg.overlap_flag)
            assert g.ar_coeffs_y == (seg.ar_coeffs_y[:npl] if ny else [])
            assert g.ar_coeffs_cb == (seg.ar_coeffs_cb[:npc] if (csfl or ncb) else [0])
            assert g.ar_coeffs_cr == (seg.ar_coeffs_cr[:npc] if (csfl or ncr) else [0])
            assert (g.cb_mult, g.cb_luma_mult, g.cb_offset) == ((seg.cb_mult, seg.cb_luma_mult, seg.cb_offset) if ncb else (0, 0, 0))
            assert (g.cr_mult, g.cr_luma_mult, g.cr_offset) == ((seg.cr_mult, seg.cr_luma_mult, seg.cr_offset) if ncr else (0, 0, 0))
