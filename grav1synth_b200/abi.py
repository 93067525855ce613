"""ctypes mirror of include/g1s.h (the C-ABI structs) plus small helpers.

The structs restate, field for field, the data the reference moves across the
`av1_grain::DiffGenerator` seam: `GrainTableSegment` as consumed at
/root/reference/src/parser/grain.rs:108-133 and src/main.rs:705-713, and the borrowed
planar `Frame<T>` built at src/reader.rs:172-212.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

NUM_Y_POINTS = 14
NUM_UV_POINTS = 10
NUM_Y_COEFFS = 24
NUM_UV_COEFFS = 25

G1S_OK = 0
G1S_E_ARG = -1
G1S_E_DIMS = -2
G1S_E_CUDA = -3
G1S_E_NCCL = -4
G1S_E_NOMEM = -5
G1S_E_STATE = -6
G1S_E_IO = -7
G1S_E_STREAM = -8

GRAM_EXACT_INT, GRAM_REF_ORDER = 0, 1
MODEL_AUTO, MODEL_HOST, MODEL_DEVICE = 0, 1, 2
MODE_FULL = 0
MODE_PRODUCER = 1
MODE_CONSUMER = 2


class CSegment(C.Structure):
    _fields_ = [
        ("start_time", C.c_uint64),
        ("end_time", C.c_uint64),
        ("num_y_points", C.c_uint8),
        ("num_cb_points", C.c_uint8),
        ("num_cr_points", C.c_uint8),
        ("scaling_shift", C.c_uint8),
        ("ar_coeff_lag", C.c_uint8),
        ("ar_coeff_shift", C.c_uint8),
        ("grain_scale_shift", C.c_uint8),
        ("overlap_flag", C.c_uint8),
        ("chroma_scaling_from_luma", C.c_uint8),
        ("cb_mult", C.c_uint8),
        ("cb_luma_mult", C.c_uint8),
        ("cr_mult", C.c_uint8),
        ("cr_luma_mult", C.c_uint8),
        ("num_ar_coeffs_plus1", C.c_uint8 * 3),
        ("cb_offset", C.c_uint16),
        ("cr_offset", C.c_uint16),
        ("random_seed", C.c_uint16),
        ("clip_to_restricted_range", C.c_uint16),
        ("scaling_points_y", (C.c_uint8 * 2) * NUM_Y_POINTS),
        ("scaling_points_cb", (C.c_uint8 * 2) * NUM_UV_POINTS),
        ("scaling_points_cr", (C.c_uint8 * 2) * NUM_UV_POINTS),
        ("ar_coeffs_y", C.c_int8 * NUM_Y_COEFFS),
        ("ar_coeffs_cb", C.c_int8 * NUM_UV_COEFFS),
        ("ar_coeffs_cr", C.c_int8 * NUM_UV_COEFFS),
    ]


class CFrame(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("stride_bytes", C.c_size_t * 3), ("width", C.c_int32),
                ("height", C.c_int32)]


class CDiffConfig(C.Structure):
    _fields_ = [
        ("fps_num", C.c_int64),
        ("fps_den", C.c_int64),
        ("src_bit_depth", C.c_int32),
        ("den_bit_depth", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("ss_x", C.c_int32),
        ("ss_y", C.c_int32),
        ("monochrome", C.c_int32),
        ("device", C.c_int32),
        ("batch_frames", C.c_int32),
        ("mode", C.c_int32),
        ("gram_kernel", C.c_int32),
        ("host_threads", C.c_int32),
        ("host_narrow", C.c_int32),
        ("gram_order", C.c_int32),
        ("n_devices", C.c_int32),
        ("device_ids", C.c_int32 * 8),
        ("model_placement", C.c_int32),
        ("reserved_", C.c_int32 * 3),
    ]


class CFilterOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32), ("d", C.c_int32)]


FILTER_CROP, FILTER_RESIZE = 0, 1
RESIZE_ALGS = {"hermite": 0, "catmullrom": 1, "mitchell": 2, "lanczos": 3, "spline36": 4}


@dataclass
class GrainTableSegment:
    """Python view of one segment (same field names as av1_grain::GrainTableSegment)."""

    start_time: int
    end_time: int
    scaling_points_y: List[Tuple[int, int]]
    scaling_points_cb: List[Tuple[int, int]]
    scaling_points_cr: List[Tuple[int, int]]
    scaling_shift: int
    ar_coeff_lag: int
    ar_coeffs_y: List[int]
    ar_coeffs_cb: List[int]
    ar_coeffs_cr: List[int]
    ar_coeff_shift: int
    cb_mult: int
    cb_luma_mult: int
    cb_offset: int
    cr_mult: int
    cr_luma_mult: int
    cr_offset: int
    overlap_flag: bool
    chroma_scaling_from_luma: bool
    grain_scale_shift: int
    random_seed: int
    raw: CSegment = field(repr=False, compare=False, default=None)

    @staticmethod
    def from_c(s: CSegment) -> "GrainTableSegment":
        copy = CSegment.from_buffer_copy(bytes(s))
        return GrainTableSegment(
            start_time=int(s.start_time),
            end_time=int(s.end_time),
            scaling_points_y=[(int(p[0]), int(p[1])) for p in list(s.scaling_points_y)[: s.num_y_points]],
            scaling_points_cb=[(int(p[0]), int(p[1])) for p in list(s.scaling_points_cb)[: s.num_cb_points]],
            scaling_points_cr=[(int(p[0]), int(p[1])) for p in list(s.scaling_points_cr)[: s.num_cr_points]],
            scaling_shift=int(s.scaling_shift),
            ar_coeff_lag=int(s.ar_coeff_lag),
            ar_coeffs_y=_coeffs(s.ar_coeffs_y, s.num_ar_coeffs_plus1[0]),
            ar_coeffs_cb=_coeffs(s.ar_coeffs_cb, s.num_ar_coeffs_plus1[1]),
            ar_coeffs_cr=_coeffs(s.ar_coeffs_cr, s.num_ar_coeffs_plus1[2]),
            ar_coeff_shift=int(s.ar_coeff_shift),
            cb_mult=int(s.cb_mult),
            cb_luma_mult=int(s.cb_luma_mult),
            cb_offset=int(s.cb_offset),
            cr_mult=int(s.cr_mult),
            cr_luma_mult=int(s.cr_luma_mult),
            cr_offset=int(s.cr_offset),
            overlap_flag=bool(s.overlap_flag),
            chroma_scaling_from_luma=bool(s.chroma_scaling_from_luma),
            grain_scale_shift=int(s.grain_scale_shift),
            random_seed=int(s.random_seed),
            raw=copy,
        )


def _coeffs(arr, count_plus1: int) -> List[int]:
    """All coefficients (diff path, count byte 0) or the explicit count the inspect path recorded."""
    vals = [int(v) for v in arr]
    return vals if count_plus1 == 0 else vals[: count_plus1 - 1]


def segments_to_c(segs: Sequence[GrainTableSegment]):
    arr = (CSegment * max(1, len(segs)))()
    for i, s in enumerate(segs):
        arr[i] = s.raw
    return arr


def frame_from_planes(planes: Sequence[np.ndarray]) -> Tuple[CFrame, list]:
    """Borrow numpy planes (uint8 or uint16, C-contiguous rows) as a CFrame.

    Returns the struct and the list of arrays that must stay alive while it is used.
    """
    f = CFrame()
    f.height, f.width = planes[0].shape
    keep = []
    for i in range(3):
        if i < len(planes) and planes[i] is not None:
            p = planes[i]
            if p.dtype not in (np.uint8, np.uint16) or p.ndim != 2 or p.strides[1] != p.itemsize:
                raise ValueError("planes must be 2-D uint8/uint16 arrays with contiguous rows")
            f.plane[i] = p.ctypes.data
            f.stride_bytes[i] = p.strides[0]
            keep.append(p)
        else:
            f.plane[i] = None
            f.stride_bytes[i] = 0
    return f, keep
