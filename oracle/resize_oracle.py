"""Test infrastructure (never imported by the product): numpy statement of the source-filter chain the engine runs on
the device (grav1synth_b200/csrc/g1s_filters.cu), i.e. of what the reference does to the SOURCE frame before diff_frame
(/root/reference/src/main.rs:621-624 -> src/filters.rs:112-182 -> video_resize::{crop, resize}).

PARITY UNPINNED for resize: the arithmetic lives in the un-vendored crate video-resize 0.2.0, absent from this box, and
nothing executable here descends from it.  Restated: the published algorithm it ports (zimg's separable resampler) --
kernels hermite / catmullrom / mitchell (bicubic, support 2), lanczos (3 taps a side), spline36 (support 3); weights from
scale = dst / src, step = min(scale, 1), 2 * ceil(support / step) taps centred on (i + 0.5) / scale, mirrored at the
edges, normalised -- with the choices listed at the top of g1s_filters.cu (f32 accumulate in tap order, floor(x + 0.5),
horizontal pass first and rounded, chroma resized like luma without a siting shift).  crop is exact (slicing).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

ALGS = ("hermite", "catmullrom", "mitchell", "lanczos", "spline36")


def _poly3(x, c0, c1, c2, c3):
    return c0 + x * (c1 + x * (c2 + x * c3))


def _bicubic(x, b, c):
    x = abs(x)
    if x < 1.0:
        return _poly3(x, (6.0 - 2.0 * b) / 6.0, 0.0, (-18.0 + 12.0 * b + 6.0 * c) / 6.0, (12.0 - 9.0 * b - 6.0 * c) / 6.0)
    if x < 2.0:
        return _poly3(x, (8.0 * b + 24.0 * c) / 6.0, (-12.0 * b - 48.0 * c) / 6.0, (6.0 * b + 30.0 * c) / 6.0,
                      (-b - 6.0 * c) / 6.0)
    return 0.0


def _sinc(x):
    return 1.0 if x == 0.0 else math.sin(math.pi * x) / (math.pi * x)


def kernel(alg: str, x: float) -> float:
    if alg == "hermite":
        return _bicubic(x, 0.0, 0.0)
    if alg == "catmullrom":
        return _bicubic(x, 0.0, 0.5)
    if alg == "mitchell":
        return _bicubic(x, 1.0 / 3.0, 1.0 / 3.0)
    x = abs(x)
    if alg == "lanczos":
        return _sinc(x) * _sinc(x / 3.0) if x < 3.0 else 0.0
    if x < 1.0:
        return _poly3(x, 1.0, -3.0 / 209.0, -453.0 / 209.0, 13.0 / 11.0)
    if x < 2.0:
        return _poly3(x - 1.0, 0.0, -156.0 / 209.0, 270.0 / 209.0, -6.0 / 11.0)
    if x < 3.0:
        return _poly3(x - 2.0, 0.0, 26.0 / 209.0, -45.0 / 209.0, 1.0 / 11.0)
    return 0.0


def support(alg: str) -> int:
    return 3 if alg in ("lanczos", "spline36") else 2


def table(alg: str, src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """(left [dst] int32, coef [dst, taps] float32)."""
    scale = dst / src
    step = min(scale, 1.0)
    fsize = max(math.ceil(support(alg) / step) * 2, 1)
    taps = min(fsize, src)
    left = np.zeros(dst, np.int32)
    coef = np.zeros((dst, taps), np.float32)
    for i in range(dst):
        pos = (i + 0.5) / scale
        begin = math.floor(pos - fsize / 2.0 + 0.5) + 0.5
        total = 0.0
        for j in range(fsize):
            total += kernel(alg, (begin + j - pos) * step)
        idx, wt = [], []
        for j in range(fsize):
            xpos = begin + j
            real = -xpos if xpos < 0.0 else (2.0 * src - xpos if xpos >= src else xpos)
            real = min(max(real, 0.0), math.nextafter(float(src), 0.0))
            idx.append(int(math.floor(real)))
            wt.append(kernel(alg, (xpos - pos) * step) / total)
        l = max(min(min(idx), src - taps), 0)
        row = [0.0] * taps
        for k, w in zip(idx, wt):
            row[min(k - l, taps - 1)] += w
        left[i] = l
        coef[i] = np.array(row, np.float64).astype(np.float32)
    return left, coef


def _pass(plane: np.ndarray, left: np.ndarray, coef: np.ndarray, maxv: int, axis: int) -> np.ndarray:
    """One separable pass along `axis` (1: horizontal) in the device's arithmetic: f32, taps in order."""
    src = plane.astype(np.float32)
    if axis == 0:
        src = src.T
    rows, _ = src.shape
    dst_n, taps = coef.shape
    acc = np.zeros((rows, dst_n), np.float32)
    for j in range(taps):
        acc = (acc + coef[:, j][None, :] * src[:, (left + j)]).astype(np.float32)
    out = np.clip(np.floor((acc + np.float32(0.5)).astype(np.float32)), 0, maxv).astype(plane.dtype)
    return np.ascontiguousarray(out.T if axis == 0 else out)


def resize_planes(planes: Sequence[np.ndarray], width: int, height: int, alg: str, bit_depth: int, ss: Tuple[int, int],
                  tables=table) -> List[np.ndarray]:
    maxv = (1 << bit_depth) - 1
    out = []
    for c, p in enumerate(planes):
        ow = width if c == 0 else (width + ss[0]) >> ss[0]
        oh = height if c == 0 else (height + ss[1]) >> ss[1]
        lh, ch = tables(alg, p.shape[1], ow)
        lv, cv = tables(alg, p.shape[0], oh)
        out.append(_pass(_pass(p, lh, ch, maxv, 1), lv, cv, maxv, 0))
    return out


def crop_planes(planes: Sequence[np.ndarray], top: int, bottom: int, left: int, right: int, ss: Tuple[int, int]):
    h, w = planes[0].shape
    out = [planes[0][top:h - bottom, left:w - right]]
    for p in planes[1:]:
        ch, cw = p.shape
        out.append(p[top >> ss[1]:ch - (bottom >> ss[1]), left >> ss[0]:cw - (right >> ss[0])])
    return [np.ascontiguousarray(p) for p in out]


def apply_chain(planes: Sequence[np.ndarray], ops, bit_depth: int, ss: Tuple[int, int], tables=table):
    """ops: list of ("crop", top, bottom, left, right) / ("resize", width, height, alg)."""
    cur = list(planes)
    for op in ops:
        if op[0] == "crop":
            cur = crop_planes(cur, *op[1:], ss)
        else:
            cur = resize_planes(cur, op[1], op[2], op[3], bit_depth, ss, tables)
    return cur
