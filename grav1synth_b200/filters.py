"""`--filters` chain of the `diff` command: grammar and messages of the reference's FilterChain::new
(/root/reference/src/filters.rs:16-110).  Filters apply to the SOURCE frame only, before the diff
(src/main.rs:621-624).

This module is the grammar only.  The chain itself runs on the device (csrc/g1s_filters.cu, reached through
DiffGenerator.set_source_filters / g1s_diff_set_source_filters): crop as a pointer offset, resize as two separable
passes.  The resampling arithmetic of the un-vendored crate `video-resize 0.2.0` cannot be pinned
here; the published kernels it ports are restated (parity unpinned).  `apply` below is the host-side crop kept for
callers that want cropped numpy planes; it refuses resize.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Union

import numpy as np

ALGORITHMS = ("hermite", "catmullrom", "mitchell", "lanczos", "spline36")


class FilterError(ValueError):
    pass


@dataclass
class Crop:
    top: int = 0
    bottom: int = 0
    left: int = 0
    right: int = 0


@dataclass
class Resize:
    width: int
    height: int
    alg: str = "catmullrom"


def _parse_usize(value: str) -> int:
    if not value.isdigit():
        raise FilterError("invalid digit found in string")  # Rust's ParseIntError text
    return int(value)


class FilterChain:
    def __init__(self, filters: str):
        self.filters: List[Union[Crop, Resize]] = []
        if not filters:
            return
        for flt in filters.split(";"):
            if ":" not in flt:
                raise FilterError(f'Invalid filter syntax in "{flt}"')
            name, args = flt.split(":", 1)
            if name == "crop":
                c = Crop()
                for arg in args.split(","):
                    if "=" not in arg:
                        raise FilterError(f'Invalid filter syntax in "{arg}"')
                    k, v = arg.split("=", 1)
                    if k not in ("top", "bottom", "left", "right"):
                        raise FilterError(f'Unrecognized crop arg "{k}"')
                    setattr(c, k, _parse_usize(v))
                self.filters.append(c)
            elif name == "resize":
                w = h = 0
                alg = "catmullrom"
                for arg in args.split(","):
                    if "=" not in arg:
                        raise FilterError(f'Invalid filter syntax in "{arg}"')
                    k, v = arg.split("=", 1)
                    if k == "width":
                        w = _parse_usize(v)
                    elif k == "height":
                        h = _parse_usize(v)
                    elif k == "alg":
                        if v not in ALGORITHMS:
                            raise FilterError(f'Unrecognized resize algorithm "{v}"')
                        alg = v
                    else:
                        raise FilterError(f'Unrecognized resize arg "{k}"')
                if w == 0 or h == 0:
                    raise FilterError("Both width and height must be provided to resize filter")
                self.filters.append(Resize(w, h, alg))
            else:
                raise FilterError(f'Unrecognized filter "{name}"')

    def output_size(self, width: int, height: int):
        for f in self.filters:
            if isinstance(f, Resize):
                width, height = f.width, f.height
            else:
                width, height = width - f.left - f.right, height - f.top - f.bottom
        return width, height

    def apply(self, planes: Sequence[np.ndarray], ss_x: int, ss_y: int) -> List[np.ndarray]:
        out = list(planes)
        for f in self.filters:
            if isinstance(f, Resize):
                raise FilterError("resize runs on the device: use DiffGenerator.set_source_filters")
            h, w = out[0].shape
            if f.top + f.bottom >= h or f.left + f.right >= w:
                raise FilterError("crop removes the whole frame")
            if (f.top | f.bottom) & ((1 << ss_y) - 1) or (f.left | f.right) & ((1 << ss_x) - 1):
                raise FilterError("crop offsets must be multiples of the chroma subsampling")
            cropped = [out[0][f.top:h - f.bottom, f.left:w - f.right]]
            for p in out[1:]:
                ch, cw = p.shape
                cropped.append(p[f.top >> ss_y:ch - (f.bottom >> ss_y), f.left >> ss_x:cw - (f.right >> ss_x)])
            out = cropped
        return [np.ascontiguousarray(p) for p in out] if self.filters else out
