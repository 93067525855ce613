"""Host-side mirror of the reference's `inspect` path (BASELINE configs[0]).

`BitstreamParser.get_grain_headers()` is what the reference's `Inspect` arm calls
(/root/reference/src/main.rs:172-196 -> src/parser.rs:120-173) and `aggregate_grain_headers`
(src/main.rs:713-772) turns the per-frame headers into grain-table segments.  Both run in C++ behind the
C ABI (csrc/g1s_obu.cpp: g1s_inspect_*); there is no Python parser behind this module.  CPU only -- the
reference does not touch pixels here either.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from . import abi
from .abi import CSegment, GrainTableSegment
from .diff import G1SError, lib

DISABLE, COPY_REF_FRAME, UPDATE_GRAIN = 0, 1, 2

EXPORTS = [
    "g1s_inspect_create", "g1s_inspect_destroy", "g1s_inspect_last_error", "g1s_inspect_push_packet",
    "g1s_inspect_push_file", "g1s_inspect_num_headers", "g1s_inspect_header", "g1s_inspect_finish",
    "g1s_inspect_stream_info", "g1s_obu_probe",
]


class CStreamInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "have_sequence_header", "seq_profile", "bit_depth", "monochrome", "ss_x", "ss_y", "max_frame_width",
        "max_frame_height", "film_grain_params_present", "color_primaries", "transfer_characteristics",
        "matrix_coefficients", "color_range", "order_hint_bits", "reduced_still_picture_header", "reserved_")] + \
        [("packets", C.c_uint64), ("obus", C.c_uint64)] + \
        [(n, C.c_int32) for n in ("last_frame_width", "last_frame_height", "last_tile_cols", "last_tile_rows")]


_bound = False


def _L():
    global _bound
    L = lib()
    if not _bound:
        L.g1s_inspect_create.argtypes = [C.POINTER(C.c_void_p)]
        L.g1s_inspect_destroy.argtypes = [C.c_void_p]
        L.g1s_inspect_destroy.restype = None
        L.g1s_inspect_last_error.argtypes = [C.c_void_p]
        L.g1s_inspect_last_error.restype = C.c_char_p
        L.g1s_inspect_push_packet.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.g1s_inspect_push_file.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.g1s_inspect_num_headers.argtypes = [C.c_void_p]
        L.g1s_inspect_num_headers.restype = C.c_size_t
        L.g1s_inspect_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(CSegment)]
        L.g1s_inspect_finish.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(CSegment), C.c_size_t,
                                         C.POINTER(C.c_size_t)]
        L.g1s_inspect_stream_info.argtypes = [C.c_void_p, C.POINTER(CStreamInfo)]
        L.g1s_obu_probe.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64), C.c_size_t,
                                    C.POINTER(C.c_int64), C.c_size_t, C.POINTER(CSegment)]
        L.g1s_obu_probe.restype = C.c_int64
        _bound = True
    return L


@dataclass
class FilmGrainHeader:
    """FilmGrainHeader of src/parser/grain.rs:11-16: kind + (for UPDATE_GRAIN) the parameters."""
    kind: int
    params: Optional[GrainTableSegment] = None
    clip_to_restricted_range: bool = False


class BitstreamParser:
    """`BitstreamParser::<false>` of the reference: feed packets, read the grain headers back."""

    def __init__(self):
        self._L = _L()
        self._h = C.c_void_p()
        rc = self._L.g1s_inspect_create(C.byref(self._h))
        if rc != 0:
            raise G1SError(rc, "g1s_inspect_create failed")
        self.frame_rate: Tuple[int, int] = (0, 0)

    def close(self):
        if self._h:
            self._L.g1s_inspect_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise G1SError(rc, self._L.g1s_inspect_last_error(self._h).decode())

    def push_packet(self, data: bytes) -> None:
        self._check(self._L.g1s_inspect_push_packet(self._h, data, len(data)))

    def push_file(self, path: str, frame_rate: Tuple[int, int] = (0, 0)) -> Tuple[int, int]:
        n, d = C.c_int64(frame_rate[0]), C.c_int64(frame_rate[1])
        self._check(self._L.g1s_inspect_push_file(self._h, path.encode(), C.byref(n), C.byref(d)))
        self.frame_rate = (n.value, d.value)
        return self.frame_rate

    def get_grain_headers(self) -> List[FilmGrainHeader]:
        out = []
        for i in range(self._L.g1s_inspect_num_headers(self._h)):
            kind, seg = C.c_int32(), CSegment()
            self._check(self._L.g1s_inspect_header(self._h, i, C.byref(kind), C.byref(seg)))
            if kind.value == UPDATE_GRAIN:
                out.append(FilmGrainHeader(kind.value, GrainTableSegment.from_c(seg), bool(seg.clip_to_restricted_range)))
            else:
                out.append(FilmGrainHeader(kind.value))
        return out

    def aggregate_grain_headers(self, fps_num: int, fps_den: int) -> List[GrainTableSegment]:
        n = C.c_size_t(0)
        cap = 64
        while True:
            arr = (CSegment * cap)()
            rc = self._L.g1s_inspect_finish(self._h, fps_num, fps_den, arr, cap, C.byref(n))
            if rc == abi.G1S_E_STATE and n.value > cap:
                cap = n.value
                continue
            self._check(rc)
            return [GrainTableSegment.from_c(arr[i]) for i in range(n.value)]

    def stream_info(self) -> dict:
        info = CStreamInfo()
        self._check(self._L.g1s_inspect_stream_info(self._h, C.byref(info)))
        return {n: int(getattr(info, n)) for n, _ in CStreamInfo._fields_ if n != "reserved_"}


def probe(what: str, data: bytes, args: Sequence[int] = (), nout: int = 16):
    """Parse one syntax group from raw bits (test hook): returns (bits_consumed, outputs, CSegment)."""
    L = _L()
    a = (C.c_int64 * max(1, len(args)))(*args)
    o = (C.c_int64 * nout)()
    seg = CSegment()
    rc = L.g1s_obu_probe(what.encode(), data, len(data), a, len(args), o, nout, C.byref(seg))
    return int(rc), [int(v) for v in o], seg


# ---------------------------------------------------------------- apply / remove
EXPORTS += ["g1s_rewrite_create", "g1s_rewrite_packet", "g1s_rewrite_take", "g1s_rewrite_counters"]


class GrainRewriter(BitstreamParser):
    """`BitstreamParser::<true>` (src/parser.rs:74-101, :175-348): packets in, packets with rewritten film grain
    headers out.  `table` = segments to apply (None removes film grain).  All rewriting is in csrc/g1s_obu.cpp."""

    def __init__(self, table: Optional[Sequence[GrainTableSegment]] = None):
        self._L = _L()
        L = self._L
        L.g1s_rewrite_create.argtypes = [C.POINTER(CSegment), C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.g1s_rewrite_packet.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_size_t)]
        L.g1s_rewrite_take.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.g1s_rewrite_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self._h = C.c_void_p()
        self.frame_rate = (0, 0)
        if table is not None:
            arr = abi.segments_to_c(table)
            for i, s in enumerate(table):
                # tables read from text carry no clip flag: the reference sets it (src/parser/grain.rs:130)
                arr[i].clip_to_restricted_range = 1
            rc = L.g1s_rewrite_create(arr, len(table), 1, C.byref(self._h))
        else:
            rc = L.g1s_rewrite_create(None, 0, 0, C.byref(self._h))
        if rc != 0:
            raise G1SError(rc, "g1s_rewrite_create failed")

    def rewrite_packet(self, data: bytes, packet_ts: int) -> bytes:
        n = C.c_size_t(0)
        self._check(self._L.g1s_rewrite_packet(self._h, data, len(data), packet_ts, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        self._check(self._L.g1s_rewrite_take(self._h, buf, n.value))
        return buf.raw[: n.value]

    def counters(self) -> Tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._L.g1s_rewrite_counters(self._h, C.byref(a), C.byref(b))
        return a.value, b.value


def pts_to_av1_ts(pts: int, time_base: Tuple[int, int]) -> int:
    """ffmpeg_pts_to_av1_ts (src/parser.rs:103-118): ceil(pts * num * 1e7 / den), 0 for negative pts."""
    num, den = time_base
    if pts < 0 or den == 0:
        return 0
    return -(-(pts * num * 10_000_000) // den)


def rewrite_ivf(data: bytes, rewriter: GrainRewriter) -> bytes:
    """IVF in, IVF out (the reference remuxes through FFmpeg; IVF is the container this build reads and writes)."""
    import struct
    if data[:4] != b"DKIF" or data[8:12] != b"AV01":
        raise ValueError("not an AV01 IVF file")
    hdr_len = struct.unpack_from("<H", data, 6)[0]
    rate, scale = struct.unpack_from("<II", data, 16)
    out = bytearray(data[:max(32, hdr_len)])
    off = max(32, hdr_len)
    while off + 12 <= len(data):
        size, pts = struct.unpack_from("<IQ", data, off)
        pkt = data[off + 12: off + 12 + size]
        if len(pkt) != size:
            raise ValueError("truncated IVF frame")
        new = rewriter.rewrite_packet(pkt, pts_to_av1_ts(pts, (scale, rate)))
        out += struct.pack("<IQ", len(new), pts) + new
        off += 12 + size
    return bytes(out)


# ---------------------------------------------------------------- generate (photon noise)
EXPORTS += ["g1s_generate_photon_noise"]
TRANSFER_BT1886, TRANSFER_SMPTE2084, TRANSFER_BT470BG = 0, 1, 2


def generate_photon_noise_params(start_time: int, end_time: int, iso: int, width: int, height: int,
                                 transfer: int = TRANSFER_BT1886, chroma_grain: bool = False,
                                 random_seed: Optional[int] = None, full_range: bool = True) -> GrainTableSegment:
    """av1_grain::generate_photon_noise_params as the reference calls it (src/main.rs:288-303)."""
    L = _L()
    L.g1s_generate_photon_noise.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int32,
                                            C.c_uint64, C.c_uint64, C.POINTER(CSegment)]
    seg = CSegment()
    rc = L.g1s_generate_photon_noise(iso, width, height, transfer, int(chroma_grain), int(full_range),
                                     -1 if random_seed is None else random_seed, start_time, end_time, C.byref(seg))
    if rc != 0:
        raise G1SError(rc, "g1s_generate_photon_noise: bad arguments")
    return GrainTableSegment.from_c(seg)
