"""Shared test helpers: the synthetic corpus and a vectorised numpy statement of the
per-frame integer record (third, independent implementation used to cross-check both
the C oracle and the CUDA kernels)."""
from __future__ import annotations

import os
import sys
from typing import List, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from grav1synth_b200.synth import SynthSpec, make_pair_numpy  # noqa: E402

COORDS = [(x, y) for y in range(-3, 1) for x in range(-3, 4) if not (y == 0 and x >= 0)]
assert len(COORDS) == 24

# name -> (spec, frames, fps)
CORPUS = {
    "c2_small_8bit": (SynthSpec(256, 192, 8, textured=0.2, sigma0=1.0, sigma1=1.5), 4, (24, 1)),
    "c3_small_10bit": (SynthSpec(320, 176, 10, textured=0.3, sigma0=1.0, sigma1=2.0, seed=7), 3, (30000, 1001)),
    "odd_size_8bit": (SynthSpec(203, 117, 8, textured=0.1, sigma0=1.2, sigma1=1.0, seed=11), 2, (25, 1)),
    "tiny_64x48": (SynthSpec(64, 48, 8, textured=0.0, sigma0=1.5, sigma1=0.5, seed=3), 2, (24, 1)),
    "yuv444_8bit": (SynthSpec(160, 128, 8, ss_x=0, ss_y=0, textured=0.2, sigma0=1.0, sigma1=1.5, seed=5), 2, (24, 1)),
    "yuv422_10bit": (SynthSpec(192, 96, 10, ss_x=1, ss_y=0, textured=0.2, sigma0=1.0, sigma1=1.5, seed=9), 2, (24, 1)),
    "heavy_grain_12bit": (SynthSpec(192, 160, 12, textured=0.0, sigma0=4.0, sigma1=8.0, seed=13), 2, (24, 1)),
}
# streams without goldens (compared in-test), kept out of CORPUS so that the parametrised suites do not pick them up
EXTRA = {
    # long enough for runs of frames to reach the batched fold (NoiseModel::fold_run)
    "tiny_long": (SynthSpec(96, 64, 8, textured=0.0, sigma0=1.5, sigma1=0.5, seed=31), 72, (24, 1)),
}


def corpus_frames(name: str):
    spec, n, fps = (CORPUS.get(name) or EXTRA[name])
    return spec, fps, [make_pair_numpy(spec, k) for k in range(n)]


def to8(p: np.ndarray, bd: int) -> np.ndarray:
    return p if bd == 8 else (p >> (bd - 8)).astype(np.uint8)


def obs_rects(flat: np.ndarray, nbw: int, nbh: int, pw: int, ph: int, bw: int, bh: int):
    """Observation rectangle per flat block: (bx, by, x0, x1, y0, y1) in block-local coordinates."""
    out = []
    for by in range(nbh):
        for bx in range(nbw):
            b = by * nbw + bx
            if not flat[b]:
                continue
            y0 = 0 if (by > 0 and flat[b - nbw]) else 3
            x0 = 0 if (bx > 0 and flat[b - 1]) else 3
            y1 = min(ph - by * bh, bh)
            x1 = min(pw - bx * bw - 3, bw if (bx + 1 < nbw and flat[b + 1]) else bw - 3)
            out.append((bx, by, x0, x1, y0, y1))
    return out


def numpy_record(src: Sequence[np.ndarray], den: Sequence[np.ndarray], src_bd: int, den_bd: int, ss_x: int,
                 ss_y: int, flat: np.ndarray) -> dict:
    """gram[3][26][26], nobs[3], luma_sum[nb], rsum[3][nb], rsq[3][nb] from numpy (int64 exact)."""
    h, w = src[0].shape
    nbw, nbh = (w + 31) // 32, (h + 31) // 32
    nb = nbw * nbh
    s8 = [to8(p, src_bd).astype(np.int64) for p in src]
    d8 = [to8(p, den_bd).astype(np.int64) for p in den]
    res = [a - b for a, b in zip(s8, d8)]
    gram = np.zeros((3, 26, 26), np.int64)
    nobs = np.zeros(3, np.int64)
    luma_sum = np.zeros(nb, np.uint32)
    rsum = np.zeros((3, nb), np.int32)
    rsq = np.zeros((3, nb), np.uint32)
    planes = len(src)
    for c in range(planes):
        sx, sy = (ss_x, ss_y) if c else (0, 0)
        bw, bh = 32 >> sx, 32 >> sy
        pw, ph = w >> sx, h >> sy
        r = res[c]
        H, W = r.shape
        mask = np.zeros((H, W), np.int64)
        for (bx, by, x0, x1, y0, y1) in obs_rects(flat, nbw, nbh, pw, ph, bw, bh):
            if x1 > x0 and y1 > y0:
                mask[by * bh + y0:by * bh + y1, bx * bw + x0:bx * bw + x1] = 1
        nobs[c] = mask.sum()
        pad = np.pad(r, 3)
        taps = []
        for (cx, cy) in COORDS:
            taps.append(pad[3 + cy:3 + cy + H, 3 + cx:3 + cx + W])
        if c:
            l = res[0][: ph << sy, : pw << sx].reshape(ph, 1 << sy, pw, 1 << sx).sum(axis=(1, 3))
            l4 = np.zeros((H, W), np.int64)
            l4[:ph, :pw] = l
            taps.append(l4)
        else:
            taps.append(np.zeros((H, W), np.int64))
        taps.append(r)
        T = np.stack([t * mask for t in taps]).reshape(26, -1)
        U = np.stack(taps).reshape(26, -1)
        gram[c] = T @ U.T
        for by in range(nbh):
            for bx in range(nbw):
                b = by * nbw + bx
                if not flat[b]:
                    continue
                mw, mh = min(pw - bx * bw, bw), min(ph - by * bh, bh)
                blk = r[by * bh:by * bh + mh, bx * bw:bx * bw + mw]
                rsum[c, b] = blk.sum()
                rsq[c, b] = (blk * blk).sum()
                if c == 0:
                    luma_sum[b] = s8[0][by * 32:by * 32 + min(h - by * 32, 32), bx * 32:bx * 32 + min(w - bx * 32, 32)].sum()
    return dict(gram=gram, nobs=nobs, luma_sum=luma_sum, rsum=rsum, rsq=rsq, num_flat=int((flat != 0).sum()))


def gram_to_pairs(G: np.ndarray) -> np.ndarray:
    iu = np.triu_indices(26)
    return G[iu]


def table_text(segs) -> str:
    from grav1synth_b200.diff import format_grain_table
    return format_grain_table(segs)
