"""Inputs shared by the libaom pin (tests/test_aom_pin.py, tests/golden/make_aom_golden.py, the GPU parity
tests): the seeded corpus plus the edge cases of tests/test_gpu_parity.py, every one as 8-bit-reducible
frame pairs.  name -> (frames, bit_depth, (ss_x, ss_y), fps)."""
from __future__ import annotations

import numpy as np

from helpers import CORPUS, corpus_frames
from grav1synth_b200.synth import SynthSpec, make_pair_numpy


def _segcut():
    a = SynthSpec(256, 192, 8, textured=0.0, sigma0=1.0, sigma1=0.5, ar_strength=0.0, seed=1)
    b = SynthSpec(256, 192, 8, textured=0.0, sigma0=2.2, sigma1=0.5, ar_strength=0.6, seed=2)
    return [make_pair_numpy(a, k) for k in range(3)] + [make_pair_numpy(b, k) for k in range(3)]


def _saturated():
    rng = np.random.default_rng(5)
    h, w = 128, 192
    den = [np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.full((h // 2, w // 2), 255, np.uint8)]
    src = [np.full((h, w), 255, np.uint8), np.full((h // 2, w // 2), 255, np.uint8),
           np.zeros((h // 2, w // 2), np.uint8)]
    src[0][:, : w // 2] = rng.integers(0, 256, (h, w // 2), dtype=np.uint8)
    den[0][:, : w // 2] = rng.integers(0, 256, (h, w // 2), dtype=np.uint8)
    return [(src, den), (den, src)]


def _sparse_overflow():
    _, _, frames = corpus_frames("c2_small_8bit")
    frames = [([p.copy() for p in s], [p.copy() for p in d]) for s, d in frames]
    for k, (s, d) in enumerate(frames):
        s[0][37 + k, 70], d[0][37 + k, 70] = 255, 3
        s[1][50, 101], d[1][50, 101] = 0, 200
        s[2][5, 3], d[2][5, 3] = 129, 0
        s[0][100, 200], d[0][100, 200] = 0, 128
    return frames


def _flat_everything():
    y = np.full((96, 128), 100, np.uint8)
    c = np.full((48, 64), 128, np.uint8)
    return [([y, c, c], [y, c, c])]


def _single_block():
    rng = np.random.default_rng(0)
    y = rng.integers(0, 255, (32, 32), dtype=np.uint8)
    c = np.full((16, 16), 128, np.uint8)
    return [([y, c, c], [y, c, c])]


def _zero_frame_mid_stream():
    _, _, frames = corpus_frames("c2_small_8bit")
    return list(frames[:2]) + [(frames[2][1], frames[2][1])] + list(frames[2:])


def _random_spec(seed):
    r = np.random.default_rng(1000 + seed)
    w = int(r.integers(40, 420))
    h = int(r.integers(40, 260))
    bd = int(r.choice([8, 10, 12]))
    return SynthSpec(w, h, bd, textured=float(r.choice([0.0, 0.3, 0.6, 0.9])), sigma0=float(r.uniform(0.8, 2.5)),
                     sigma1=float(r.uniform(0.0, 2.0)), ar_strength=float(r.uniform(0, 0.5)), seed=seed)


def _random(seed):
    spec = _random_spec(seed)
    return [make_pair_numpy(spec, k) for k in range(2)], spec.bit_depth


def _long():
    spec = SynthSpec(640, 352, 8, textured=0.15, sigma0=1.5, sigma1=2.5, seed=77)
    return [make_pair_numpy(spec, k) for k in range(12)]


def _hd_frame():
    spec = SynthSpec(1920, 1080, 8, textured=0.1, sigma0=1.0, sigma1=1.5, seed=2026)
    return [make_pair_numpy(spec, 0)]


def _uhd_frame():
    # BASELINE configs[2] geometry: one 3840x2160 10-bit 4:2:0 pair (the frame of test_full_size_4k_10bit_frame_against_oracle)
    spec = SynthSpec(3840, 2160, 10, textured=0.1, sigma0=1.0, sigma1=1.5, seed=2026)
    return [make_pair_numpy(spec, 0)]


def codec_source(n=3, w=352, h=288, seed=3):
    """Textured, noisy 8-bit 4:2:0 source frames for the codec pair (deterministic)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(0, 1, w)[None, :]
    ys = np.linspace(0, 1, h)[:, None]
    out = []
    for k in range(n):
        base = 30 + 190 * (0.5 + 0.5 * np.sin(3 * xs + 2 * ys + 0.05 * k))
        tex = 25 * (np.sin(40 * xs) * np.sin(37 * ys) > 0.8)
        y = base + tex + rng.normal(0, 3.0 + 3 * base / 255, (h, w))
        u = 128 + 25 * np.sin(5 * xs[:, ::2] + 0 * ys[::2]) + rng.normal(0, 1.5, (h // 2, w // 2))
        v = 128 + 25 * np.cos(4 * ys[::2] + 0 * xs[:, ::2]) + rng.normal(0, 1.5, (h // 2, w // 2))
        out.append([np.clip(np.rint(p), 0, 255).astype(np.uint8) for p in (y, u, v)])
    return out


def _codec_pair():
    """Source frames against what a real codec made of them: the 'denoised' side is libaom's DECODE of the committed
    stream tests/golden/aom/codec_pair_cq28.ivf (libaom's own encode of the source at constant quality, no film grain;
    tests/golden/make_aom_golden.py writes it).  Needs the bundled libaom at test time to decode."""
    import os
    import struct
    from oracle import aom_encode as E
    d = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aom", "codec_pair_cq28.ivf"), "rb").read()
    off, packets = 32, []
    while off + 12 <= len(d):
        sz = struct.unpack_from("<I", d, off)[0]
        packets.append(d[off + 12: off + 12 + sz])
        off += 12 + sz
    den = E.decode(packets)
    return [(s, list(dd)) for s, dd in zip(codec_source(), den)]


NEEDS_LIBAOM = {"codec_pair_cq28"}
CASES = {}
for _name, (_spec, _n, _fps) in CORPUS.items():
    CASES[_name] = (lambda n=_name: corpus_frames(n)[2], _spec.bit_depth, (_spec.ss_x, _spec.ss_y), _fps)
CASES["segment_cut"] = (_segcut, 8, (1, 1), (30000, 1001))
CASES["saturated_residual"] = (_saturated, 8, (1, 1), (24, 1))
CASES["sparse_int8_overflow"] = (_sparse_overflow, 8, (1, 1), (24, 1))
CASES["flat_everything"] = (_flat_everything, 8, (1, 1), (24, 1))
CASES["single_block"] = (_single_block, 8, (1, 1), (24, 1))
CASES["zero_frame_mid_stream"] = (_zero_frame_mid_stream, 8, (1, 1), (24, 1))
for _s in range(8):
    CASES[f"random_{_s}"] = (lambda s=_s: _random(s)[0], _random_spec(_s).bit_depth, (1, 1), (24, 1))
CASES["long_12_frames"] = (_long, 8, (1, 1), (24, 1))
CASES["hd_1080p_frame"] = (_hd_frame, 8, (1, 1), (24, 1))
CASES["uhd_4k_10bit_frame"] = (_uhd_frame, 10, (1, 1), (24, 1))
CASES["codec_pair_cq28"] = (_codec_pair, 8, (1, 1), (24, 1))

# Cases where the exact-integer Gram (what the CUDA engine accumulates) lands on the other side of a structural
# tie inside fit_piecewise than the reference's per-term f64 accumulation does: the
# REF_ORDER oracle equals libaom there, the EXACT_INT oracle and the engine pick a different, equally valid point set.
EXACT_INT_TIE_FLIPS = {"segment_cut"}


def load(name):
    make, bd, ss, fps = CASES[name]
    return make(), bd, ss, fps
