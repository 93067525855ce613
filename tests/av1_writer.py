"""A small AV1 *header writer* for tests: builds sequence-header and frame OBUs (AV1 spec 5.5 / 5.9, field order
as the reference reads it in /root/reference/src/parser/sequence.rs:163-653 and frame.rs:162-699) with film grain
parameters, and wraps them in IVF or a Section-5 .obu stream.  It is the encoder side of what csrc/g1s_obu.cpp
parses, written independently of it; tile data is a few dummy bytes (inspect never decodes tiles).

BitBuilder mirrors the helper the reference's own unit tests use (frame.rs:2005-2075: push_bool / push_bits /
push_su / push_ns), so the reference's test vectors can be restated line by line in tests/test_inspect.py.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple


class BitBuilder:
    def __init__(self):
        self.bits: List[int] = []

    def push_bool(self, b) -> "BitBuilder":
        self.bits.append(1 if b else 0)
        return self

    def push_bits(self, value: int, width: int) -> "BitBuilder":
        for i in reversed(range(width)):
            self.bits.append((value >> i) & 1)
        return self

    def push_su(self, value: int, width: int) -> "BitBuilder":
        return self.push_bits(value & ((1 << width) - 1), width)

    def push_ns(self, value: int, n: int) -> "BitBuilder":
        w = n.bit_length()
        m = (1 << w) - n
        if value < m:
            return self.push_bits(value, w - 1)
        return self.push_bits(value + m, w)

    def push_uvlc(self, value: int) -> "BitBuilder":
        v = value + 1
        lz = v.bit_length() - 1
        self.push_bits(0, lz)
        return self.push_bits(v, lz + 1)

    def __len__(self):
        return len(self.bits)

    def to_bytes(self, pad_bit: int = 0) -> bytes:
        bits = list(self.bits)
        while len(bits) % 8:
            bits.append(pad_bit)
        return bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))

    def with_trailer(self) -> Tuple[bytes, int]:
        """Bytes with two recognisable trailer bytes after the payload, and the payload length in bits."""
        n = len(self.bits)
        b = BitBuilder()
        b.bits = list(self.bits)
        b.push_bits(0b1010_0101_1100_0011, 16)
        return b.to_bytes(), n


def leb128(v: int) -> bytes:
    out = bytearray()
    while True:
        byte = v & 0x7F
        v >>= 7
        out.append(byte | (0x80 if v else 0))
        if not v:
            return bytes(out)


def obu(obu_type: int, payload: bytes, has_size: bool = True, ext: Optional[Tuple[int, int]] = None) -> bytes:
    hdr = bytearray([(obu_type << 3) | ((1 if ext else 0) << 2) | ((1 if has_size else 0) << 1)])
    if ext:
        hdr.append((ext[0] << 5) | (ext[1] << 3))
    return bytes(hdr) + (leb128(len(payload)) if has_size else b"") + payload


OBU_SEQUENCE_HEADER, OBU_TEMPORAL_DELIMITER, OBU_FRAME_HEADER, OBU_TILE_GROUP, OBU_METADATA, OBU_FRAME, \
    OBU_REDUNDANT_FRAME_HEADER, OBU_PADDING = 1, 2, 3, 4, 5, 6, 7, 15


def temporal_delimiter() -> bytes:
    return obu(OBU_TEMPORAL_DELIMITER, b"")


@dataclass
class Seq:
    profile: int = 0
    width: int = 352
    height: int = 288
    bit_depth: int = 8
    monochrome: bool = False
    order_hint_bits: int = 7          # 0 disables order hints
    film_grain_params_present: bool = True
    use_128: bool = False
    enable_superres: bool = False
    enable_cdef: bool = True
    enable_restoration: bool = True
    enable_warped_motion: bool = False
    enable_ref_frame_mvs: bool = False
    choose_screen_content_tools: bool = True   # SELECT_SCREEN_CONTENT_TOOLS
    timing_info: bool = False
    equal_picture_interval: bool = True
    decoder_model: bool = False
    frame_id_numbers: bool = False
    operating_point_idc: Sequence[int] = (0,)
    color_description: Optional[Tuple[int, int, int]] = None
    ss: Tuple[int, int] = (1, 1)
    separate_uv_delta_q: bool = False
    reduced_still_picture_header: bool = False

    def payload(self) -> bytes:
        b = BitBuilder()
        b.push_bits(self.profile, 3).push_bool(self.reduced_still_picture_header)  # still_picture follows reduced
        b.push_bool(self.reduced_still_picture_header)
        if self.reduced_still_picture_header:
            b.push_bits(8, 5)
        else:
            b.push_bool(self.timing_info)
            if self.timing_info:
                b.push_bits(1001, 32).push_bits(24000, 32).push_bool(self.equal_picture_interval)
                if self.equal_picture_interval:
                    b.push_uvlc(0)  # num_ticks_per_picture_minus_1
                b.push_bool(self.decoder_model)
                if self.decoder_model:
                    b.push_bits(9, 5).push_bits(1, 32).push_bits(9, 5).push_bits(9, 5)
            b.push_bool(False)  # initial_display_delay_present_flag
            b.push_bits(len(self.operating_point_idc) - 1, 5)
            for idc in self.operating_point_idc:
                b.push_bits(idc, 12).push_bits(8, 5).push_bool(False)  # level 8 -> seq_tier bit
                if self.timing_info and self.decoder_model:
                    b.push_bool(True).push_bits(3, 10).push_bits(4, 10).push_bool(False)
        wb, hb = max(1, (self.width - 1).bit_length()), max(1, (self.height - 1).bit_length())
        b.push_bits(wb - 1, 4).push_bits(hb - 1, 4).push_bits(self.width - 1, wb).push_bits(self.height - 1, hb)
        if not self.reduced_still_picture_header:
            b.push_bool(self.frame_id_numbers)
            if self.frame_id_numbers:
                b.push_bits(2, 4).push_bits(1, 3)  # delta_frame_id_length 4, additional 2 -> id_len 7... (2+1+3+... )
        b.push_bool(self.use_128).push_bool(True).push_bool(True)  # filter_intra, intra_edge_filter
        if not self.reduced_still_picture_header:
            b.push_bool(False).push_bool(False).push_bool(self.enable_warped_motion).push_bool(True)
            b.push_bool(self.order_hint_bits > 0)
            if self.order_hint_bits > 0:
                b.push_bool(False).push_bool(self.enable_ref_frame_mvs)
            b.push_bool(self.choose_screen_content_tools)
            if not self.choose_screen_content_tools:
                b.push_bits(0, 1)  # seq_force_screen_content_tools = 0 -> force_integer_mv = SELECT, no more bits
            else:
                b.push_bool(True)  # seq_choose_integer_mv
            if self.order_hint_bits > 0:
                b.push_bits(self.order_hint_bits - 1, 3)
        b.push_bool(self.enable_superres).push_bool(self.enable_cdef).push_bool(self.enable_restoration)
        # color_config
        b.push_bool(self.bit_depth > 8)
        if self.profile == 2 and self.bit_depth > 8:
            b.push_bool(self.bit_depth == 12)
        if self.profile != 1:
            b.push_bool(self.monochrome)
        b.push_bool(self.color_description is not None)
        if self.color_description is not None:
            for v in self.color_description:
                b.push_bits(v, 8)
        if self.monochrome:
            b.push_bits(1, 1)  # color_range
        elif self.color_description == (1, 13, 0):
            b.push_bool(self.separate_uv_delta_q)
        else:
            b.push_bits(0, 1)  # color_range
            if self.profile == 2 and self.bit_depth == 12:
                b.push_bits(self.ss[0], 1)
                if self.ss[0]:
                    b.push_bits(self.ss[1], 1)
            if self.subsampling() == (1, 1):
                b.push_bits(0, 2)  # chroma_sample_position
            b.push_bool(self.separate_uv_delta_q)
        b.push_bool(self.film_grain_params_present)
        b.push_bool(True)  # trailing_bits
        return b.to_bytes()

    def subsampling(self) -> Tuple[int, int]:
        if self.monochrome:
            return (1, 1)
        if self.color_description == (1, 13, 0):
            return (0, 0)
        if self.profile == 0:
            return (1, 1)
        if self.profile == 1:
            return (0, 0)
        return tuple(self.ss) if self.bit_depth == 12 else (1, 0)

    def id_len(self) -> int:
        return (1 + 2 + 3) if self.frame_id_numbers else 0

    def obu(self, **kw) -> bytes:
        return obu(OBU_SEQUENCE_HEADER, self.payload(), **kw)


@dataclass
class Grain:
    """film_grain_params() content; kind: 'disable' | 'copy' | 'update'."""
    kind: str = "update"
    seed: int = 1234
    points_y: Sequence[Tuple[int, int]] = ((0, 20), (128, 40), (255, 60))
    chroma_scaling_from_luma: bool = False
    points_cb: Sequence[Tuple[int, int]] = ((0, 10), (255, 30))
    points_cr: Sequence[Tuple[int, int]] = ((64, 12),)
    scaling_shift: int = 10
    ar_coeff_lag: int = 2
    ar_coeffs_y: Optional[Sequence[int]] = None
    ar_coeffs_cb: Optional[Sequence[int]] = None
    ar_coeffs_cr: Optional[Sequence[int]] = None
    ar_coeff_shift: int = 7
    grain_scale_shift: int = 1
    cb: Tuple[int, int, int] = (128, 192, 256)
    cr: Tuple[int, int, int] = (120, 180, 300)
    overlap_flag: bool = True
    clip_to_restricted_range: bool = False
    ref_idx: int = 0

    def counts(self, mono: bool, ss: Tuple[int, int]):
        ny = len(self.points_y)
        if mono or self.chroma_scaling_from_luma or (ss == (1, 1) and ny == 0):
            ncb = ncr = 0
        else:
            ncb, ncr = len(self.points_cb), len(self.points_cr)
        npl = 2 * self.ar_coeff_lag * (self.ar_coeff_lag + 1)
        npc = npl + 1 if ny > 0 else npl
        return ny, ncb, ncr, npl, npc

    def coeffs(self, mono: bool, ss: Tuple[int, int]):
        ny, ncb, ncr, npl, npc = self.counts(mono, ss)
        y = list(self.ar_coeffs_y) if self.ar_coeffs_y is not None else [((i * 7) % 41) - 20 for i in range(npl)]
        cb = list(self.ar_coeffs_cb) if self.ar_coeffs_cb is not None else [((i * 5) % 31) - 15 for i in range(npc)]
        cr = list(self.ar_coeffs_cr) if self.ar_coeffs_cr is not None else [((i * 3) % 23) - 11 for i in range(npc)]
        has_cb = (not mono) and (self.chroma_scaling_from_luma or ncb > 0)
        has_cr = (not mono) and (self.chroma_scaling_from_luma or ncr > 0)
        return (y[:npl] if ny > 0 else []), (cb[:npc] if has_cb else None), (cr[:npc] if has_cr else None)

    def write(self, b: BitBuilder, frame_type: int, mono: bool, ss: Tuple[int, int]) -> None:
        if self.kind == "disable":
            b.push_bool(False)
            return
        b.push_bool(True).push_bits(self.seed, 16)
        if frame_type == 1:
            b.push_bool(self.kind == "update")
        if self.kind == "copy":
            assert frame_type == 1, "only inter frames can copy grain parameters"
            b.push_bits(self.ref_idx, 3)
            return
        ny, ncb, ncr, npl, npc = self.counts(mono, ss)
        b.push_bits(ny, 4)
        for v, s in self.points_y:
            b.push_bits(v, 8).push_bits(s, 8)
        if not mono:
            b.push_bool(self.chroma_scaling_from_luma)
        if not (mono or self.chroma_scaling_from_luma or (ss == (1, 1) and ny == 0)):
            b.push_bits(ncb, 4)
            for v, s in self.points_cb:
                b.push_bits(v, 8).push_bits(s, 8)
            b.push_bits(ncr, 4)
            for v, s in self.points_cr:
                b.push_bits(v, 8).push_bits(s, 8)
        b.push_bits(self.scaling_shift - 8, 2).push_bits(self.ar_coeff_lag, 2)
        y, cb, cr = self.coeffs(mono, ss)
        for c in y:
            b.push_bits(c + 128, 8)
        for c in (cb or []):
            b.push_bits(c + 128, 8)
        for c in (cr or []):
            b.push_bits(c + 128, 8)
        b.push_bits(self.ar_coeff_shift - 6, 2).push_bits(self.grain_scale_shift, 2)
        if ncb > 0:
            b.push_bits(self.cb[0], 8).push_bits(self.cb[1], 8).push_bits(self.cb[2], 9)
        if ncr > 0:
            b.push_bits(self.cr[0], 8).push_bits(self.cr[1], 8).push_bits(self.cr[2], 9)
        b.push_bool(self.overlap_flag).push_bool(self.clip_to_restricted_range)

    def expected(self, mono: bool, ss: Tuple[int, int]) -> dict:
        """What the reference's FilmGrainParams would hold after parsing this (grain.rs:136-295)."""
        ny, ncb, ncr, npl, npc = self.counts(mono, ss)
        y, cb, cr = self.coeffs(mono, ss)
        return dict(
            random_seed=self.seed, scaling_points_y=[tuple(p) for p in self.points_y],
            scaling_points_cb=[tuple(p) for p in self.points_cb] if ncb else [],
            scaling_points_cr=[tuple(p) for p in self.points_cr] if ncr else [],
            chroma_scaling_from_luma=bool(self.chroma_scaling_from_luma) and not mono,
            scaling_shift=self.scaling_shift, ar_coeff_lag=self.ar_coeff_lag, ar_coeffs_y=y,
            ar_coeffs_cb=cb if cb is not None else [0], ar_coeffs_cr=cr if cr is not None else [0],
            ar_coeff_shift=self.ar_coeff_shift, grain_scale_shift=self.grain_scale_shift,
            cb_mult=self.cb[0] if ncb else 0, cb_luma_mult=self.cb[1] if ncb else 0, cb_offset=self.cb[2] if ncb else 0,
            cr_mult=self.cr[0] if ncr else 0, cr_luma_mult=self.cr[1] if ncr else 0, cr_offset=self.cr[2] if ncr else 0,
            overlap_flag=self.overlap_flag,
        )


def tile_log2(blk: int, target: int) -> int:
    k = 0
    while (blk << k) < target:
        k += 1
    return k


@dataclass
class Frame:
    """One frame header (+ dummy tile group when as_frame_obu)."""
    frame_type: int = 0               # 0 key, 1 inter, 2 intra-only
    show_frame: bool = True
    showable_frame: bool = True       # written only for hidden frames
    error_resilient: bool = False     # written unless implied
    order_hint: int = 0
    refresh_frame_flags: int = 0xFF
    ref_frame_idx: Sequence[int] = (0, 0, 0, 0, 0, 0, 0)
    primary_ref_frame: int = 7
    base_q_idx: int = 60
    tile_cols_log2: int = 0
    reference_select: bool = False
    skip_mode_bit: Optional[bool] = None   # written after reference_select when the test knows skip mode is allowed
    allow_screen_content_tools: bool = False
    segmentation: bool = False
    seg_value: int = -7               # ALT_Q feature of segment 1 when the data is sent
    seg_inherit: bool = False         # segmentation enabled, update_map = update_data = 0: keep the reference's data
    lossless: Optional[bool] = None   # what the decoder will conclude (None: derived for the default cases)
    gm_translation_on_last: bool = False
    grain: Grain = field(default_factory=Grain)
    show_existing_frame: Optional[int] = None  # frame_to_show_map_idx: the header is only that
    short_signaling: Optional[Tuple[int, int]] = None  # (last_frame_idx, gold_frame_idx): frame_refs_short_signaling
    force_integer_mv: bool = False    # frame-level bit, coded when allow_screen_content_tools (sequence value is SELECT)

    def header_bits(self, s: Seq) -> BitBuilder:
        b = BitBuilder()
        mono, ss = s.monochrome, s.subsampling()
        num_planes = 1 if mono else 3
        if self.show_existing_frame is not None:
            b.push_bool(True).push_bits(self.show_existing_frame, 3)
            if s.timing_info and s.decoder_model and not s.equal_picture_interval:
                b.push_bits(5, 10)  # temporal_point_info: frame_presentation_time (length_minus_1 = 9)
            if s.frame_id_numbers:
                b.push_bits(0, s.id_len())
            return b
        ft, intra = self.frame_type, self.frame_type in (0, 2)
        err_res = self.error_resilient
        if not s.reduced_still_picture_header:
            b.push_bool(False).push_bits(ft, 2).push_bool(self.show_frame)
            if self.show_frame and s.timing_info and s.decoder_model and not s.equal_picture_interval:
                b.push_bits(5, 10)  # temporal_point_info
            if not self.show_frame:
                b.push_bool(self.showable_frame)
            if ft == 3 or (ft == 0 and self.show_frame):
                err_res = True
            else:
                b.push_bool(err_res)
        b.push_bool(False)  # disable_cdf_update
        sct = self.allow_screen_content_tools
        if s.reduced_still_picture_header or s.choose_screen_content_tools:
            b.push_bool(sct)
            if sct:
                b.push_bool(self.force_integer_mv)  # coded: seq_choose_integer_mv = 1 / reduced -> SELECT
        else:
            sct = False
        if s.frame_id_numbers:
            b.push_bits(self.order_hint & ((1 << s.id_len()) - 1), s.id_len())
        if not s.reduced_still_picture_header and ft != 3:
            b.push_bool(False)  # frame_size_override_flag
        b.push_bits(self.order_hint, s.order_hint_bits)
        primary = 7 if (intra or err_res) else self.primary_ref_frame
        if not (intra or err_res):
            b.push_bits(primary, 3)
        if s.timing_info and s.decoder_model:
            b.push_bool(False)  # buffer_removal_time_present_flag
        implied_refresh = ft == 3 or (ft == 0 and self.show_frame)
        if not implied_refresh:
            b.push_bits(self.refresh_frame_flags, 8)
        refresh = 0xFF if implied_refresh else self.refresh_frame_flags
        if (not intra or refresh != 0xFF) and err_res and s.order_hint_bits > 0:
            for i in range(8):
                b.push_bits((self.order_hint - 1 - i) & ((1 << s.order_hint_bits) - 1), s.order_hint_bits)
        if intra:
            if s.enable_superres:
                b.push_bool(False)
            b.push_bool(False)  # render_and_frame_size_different
            if sct:
                b.push_bool(False)  # allow_intrabc
        else:
            if s.order_hint_bits > 0:
                b.push_bool(self.short_signaling is not None)  # frame_refs_short_signaling
                if self.short_signaling is not None:
                    b.push_bits(self.short_signaling[0], 3).push_bits(self.short_signaling[1], 3)
            for idx in self.ref_frame_idx:
                if self.short_signaling is None:
                    b.push_bits(idx, 3)
                if s.frame_id_numbers:  # for every reference, short signaling or not (spec 5.9.2)
                    b.push_bits(0, 2 + 2)  # delta_frame_id_minus_1, delta_frame_id_length_minus_2 = 2
            if s.enable_superres:
                b.push_bool(False)
            b.push_bool(False)  # render_and_frame_size_different
            if not (sct and self.force_integer_mv):
                b.push_bool(True)   # allow_high_precision_mv: absent when the FRAME's force_integer_mv is 1
            b.push_bool(True)   # is_filter_switchable
            b.push_bool(False)  # is_motion_mode_switchable
            if not err_res and s.enable_ref_frame_mvs:
                b.push_bool(False)
        if not s.reduced_still_picture_header:
            b.push_bool(False)  # disable_frame_end_update_cdf
        # tile_info, uniform spacing
        mi_cols, mi_rows = 2 * ((s.width + 7) >> 3), 2 * ((s.height + 7) >> 3)
        sh = 5 if s.use_128 else 4
        sb_cols, sb_rows = (mi_cols + (1 << sh) - 1) >> sh, (mi_rows + (1 << sh) - 1) >> sh
        sb_size = sh + 2
        min_cols = tile_log2(4096 >> sb_size, sb_cols)
        max_cols = tile_log2(1, min(sb_cols, 64))
        max_rows = tile_log2(1, min(sb_rows, 64))
        min_tiles = max(min_cols, tile_log2((4096 * 2304) >> (2 * sb_size), sb_rows * sb_cols))
        b.push_bool(True)
        cols_log2 = min(max(self.tile_cols_log2, min_cols), max_cols)  # what the frame geometry allows
        for _ in range(cols_log2 - min_cols):
            b.push_bool(True)
        if cols_log2 < max_cols:
            b.push_bool(False)
        min_rows = max(min_tiles - cols_log2, 0)
        if min_rows < max_rows:
            b.push_bool(False)
        rows_log2 = min_rows
        if cols_log2 + rows_log2 > 0:
            b.push_bits(0, cols_log2 + rows_log2).push_bits(3, 2)
        self._tiles = (cols_log2, rows_log2, sb_cols, sb_rows)
        # quantization_params
        b.push_bits(self.base_q_idx, 8).push_bool(False)
        if num_planes > 1:
            if s.separate_uv_delta_q:
                b.push_bool(False)
            b.push_bool(True).push_su(-3, 7).push_bool(False)  # delta_q_u_dc = -3, u_ac = 0
        b.push_bool(False)  # using_qmatrix
        # segmentation_params
        b.push_bool(self.segmentation)
        if self.segmentation and self.seg_inherit:
            assert primary != 7
            b.push_bool(False).push_bool(False)  # segmentation_update_map, segmentation_update_data
        elif self.segmentation:
            if primary != 7:
                b.push_bool(True).push_bool(False).push_bool(True)  # update_map, temporal_update, update_data
            for seg in range(8):
                for feat in range(8):
                    on = seg == 1 and feat == 0
                    b.push_bool(on)
                    if on:
                        b.push_su(self.seg_value, 9)
        if self.base_q_idx > 0:
            b.push_bool(False)  # delta_q_present
        lossless = False  # delta_q_u_dc != 0 with chroma; monochrome: lossless iff base_q_idx == 0
        if num_planes == 1:
            lossless = self.base_q_idx == 0  # the test segment feature (-7 on ALT_Q) clamps back to qindex 0
        if self.lossless is not None:
            lossless = self.lossless
        if not lossless:
            b.push_bits(12, 6).push_bits(0, 6)  # loop_filter_level[0..1]
            if num_planes > 1:
                b.push_bits(3, 6).push_bits(4, 6)
            b.push_bits(2, 3).push_bool(True).push_bool(False)  # sharpness, delta_enabled, delta_update
            if s.enable_cdef:
                b.push_bits(1, 2).push_bits(1, 2)
                for _ in range(2):
                    b.push_bits(4, 4).push_bits(1, 2)
                    if num_planes > 1:
                        b.push_bits(2, 4).push_bits(3, 2)
            if s.enable_restoration:
                b.push_bits(1, 2)
                if num_planes > 1:
                    b.push_bits(0, 2).push_bits(2, 2)
                if s.use_128:
                    b.push_bool(True)
                else:
                    b.push_bool(True).push_bool(False)
                if num_planes > 1 and ss == (1, 1):
                    b.push_bool(True)  # lr_uv_shift
            b.push_bool(True)  # tx_mode_select
        if not intra:
            b.push_bool(self.reference_select)
            if self.skip_mode_bit is not None:
                b.push_bool(self.skip_mode_bit)
            if not err_res and s.enable_warped_motion:
                b.push_bool(False)
        b.push_bool(False)  # reduced_tx_set
        if not intra:
            for ref in range(7):
                if ref == 0 and self.gm_translation_on_last:
                    b.push_bool(True).push_bool(False).push_bool(True)
                    for _ in range(2):  # two parameters, decode_subexp(1025): more = 1 (i = 1, b2 = 3), more = 0, 3 bits
                        b.push_bool(True).push_bool(False).push_bits(5, 3)
                else:
                    b.push_bool(False)
        if s.film_grain_params_present and (self.show_frame or self.showable_frame):
            self.grain.write(b, ft, mono, ss)
        return b

    def frame_obu(self, s: Seq, tile_bytes: bytes = b"\x5a\xc3\x7e", **kw) -> bytes:
        assert self.show_existing_frame is None, "show_existing_frame travels in an OBU_FRAME_HEADER"
        b = self.header_bits(s)
        hdr = b.to_bytes()  # byte_alignment() zeros
        cols_log2, rows_log2, _, _ = self._tiles
        tg = BitBuilder()
        if cols_log2 + rows_log2 > 0:
            tg.push_bool(False)  # tile_start_and_end_present_flag
        return obu(OBU_FRAME, hdr + tg.to_bytes() + tile_bytes, **kw)

    def frame_header_obu(self, s: Seq, **kw) -> bytes:
        b = self.header_bits(s)
        b.push_bool(True)  # trailing_bits
        return obu(OBU_FRAME_HEADER, b.to_bytes(), **kw)

    def tile_group_obu(self, s: Seq, tile_bytes: bytes = b"\x11\x22", **kw) -> bytes:
        cols_log2, rows_log2, _, _ = self._tiles
        tg = BitBuilder()
        if cols_log2 + rows_log2 > 0:
            tg.push_bool(False)
        return obu(OBU_TILE_GROUP, tg.to_bytes() + tile_bytes, **kw)


def ivf(packets: Sequence[bytes], width: int, height: int, rate: int = 24, scale: int = 1) -> bytes:
    out = bytearray(b"DKIF" + struct.pack("<HH4sHHIIII", 0, 32, b"AV01", width, height, rate, scale, len(packets), 0))
    for i, p in enumerate(packets):
        out += struct.pack("<IQ", len(p), i) + p
    return bytes(out)
