"""YUV4MPEG2 (.y4m) reader — the FFmpeg-free stand-in for the reference's BitstreamReader on the `diff`
path.  Produces what decode_frame hands to DiffGenerator
(/root/reference/src/reader.rs:172-212): planar Y, Cb, Cr planes, uint8 for 8-bit and little-endian
uint16 above, plus the stream details the Diff arm needs (size, bit depth, subsampling, frame rate;
src/reader.rs:27-34).  Pure CPU plumbing: no pixel is touched here beyond the file read.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import BinaryIO, Iterator, List, Optional, Tuple

import numpy as np

# colourspace tag -> (ss_x, ss_y, bit_depth, monochrome)
_COLOURSPACES = {
    "420": (1, 1, 8, False), "420jpeg": (1, 1, 8, False), "420mpeg2": (1, 1, 8, False), "420paldv": (1, 1, 8, False),
    "422": (1, 0, 8, False), "444": (0, 0, 8, False), "mono": (1, 1, 8, True),
    "420p10": (1, 1, 10, False), "422p10": (1, 0, 10, False), "444p10": (0, 0, 10, False),
    "420p12": (1, 1, 12, False), "422p12": (1, 0, 12, False), "444p12": (0, 0, 12, False),
    "mono10": (1, 1, 10, True), "mono12": (1, 1, 12, True),
}


@dataclass
class VideoDetails:
    """Same fields as the reference's VideoDetails (src/reader.rs:27-34)."""
    width: int
    height: int
    bit_depth: int
    ss_x: int
    ss_y: int
    monochrome: bool
    fps_num: int
    fps_den: int


class Y4MReader:
    def __init__(self, path: str):
        self._f: BinaryIO = open(path, "rb")
        header = self._f.readline()
        if not header.startswith(b"YUV4MPEG2"):
            raise ValueError(f"{path}: not a YUV4MPEG2 stream")
        w = h = None
        fps = (25, 1)
        cs = "420"
        for tok in header.decode("ascii", "replace").split()[1:]:
            if tok[0] == "W":
                w = int(tok[1:])
            elif tok[0] == "H":
                h = int(tok[1:])
            elif tok[0] == "F":
                n, d = tok[1:].split(":")
                fps = (int(n), int(d))
            elif tok[0] == "C":
                cs = tok[1:]
        if not w or not h:
            raise ValueError(f"{path}: missing frame size")
        if cs not in _COLOURSPACES:
            raise ValueError(f"unsupported video format C{cs}")
        ss_x, ss_y, bd, mono = _COLOURSPACES[cs]
        if fps[0] <= 0 or fps[1] <= 0:
            raise ValueError(f"{path}: bad frame rate")
        self.details = VideoDetails(w, h, bd, ss_x, ss_y, mono, fps[0], fps[1])
        self._dtype = np.dtype(np.uint8) if bd == 8 else np.dtype("<u2")
        cw, ch = (w + ss_x) >> ss_x, (h + ss_y) >> ss_y
        self._shapes = [(h, w)] if mono else [(h, w), (ch, cw), (ch, cw)]

    def get_video_details(self) -> VideoDetails:
        return self.details

    def get_frame(self) -> Optional[List[np.ndarray]]:
        """Next frame as a list of planes, or None at end of stream (BitstreamReader::get_frame)."""
        line = self._f.readline()
        if not line:
            return None
        if not line.startswith(b"FRAME"):
            raise ValueError("corrupt y4m stream: FRAME marker expected")
        planes = []
        for (ph, pw) in self._shapes:
            n = ph * pw * self._dtype.itemsize
            buf = self._f.read(n)
            if len(buf) != n:
                return None  # truncated last frame: treated as end of stream
            planes.append(np.frombuffer(buf, dtype=self._dtype).reshape(ph, pw))
        return planes

    def __iter__(self) -> Iterator[List[np.ndarray]]:
        while True:
            f = self.get_frame()
            if f is None:
                return
            yield f

    def close(self):
        self._f.close()


def write_y4m(path: str, frames, bit_depth: int, fps: Tuple[int, int] = (24, 1), ss: Tuple[int, int] = (1, 1)) -> None:
    """Small writer used by the tests to build input clips."""
    tag = {(1, 1): "420", (1, 0): "422", (0, 0): "444"}[ss] + ("" if bit_depth == 8 else f"p{bit_depth}")
    if tag == "420":
        tag = "420jpeg"
    h, w = frames[0][0].shape
    with open(path, "wb") as f:
        f.write(f"YUV4MPEG2 W{w} H{h} F{fps[0]}:{fps[1]} Ip A1:1 C{tag}\n".encode())
        for planes in frames:
            f.write(b"FRAME\n")
            for p in planes:
                f.write(np.ascontiguousarray(p.astype(np.uint8 if bit_depth == 8 else "<u2")).tobytes())
