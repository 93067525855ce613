"""Source filters on the device (`-m gpu`): g1s_diff_set_source_filters (crop + resize on the SOURCE frame, as
/root/reference/src/main.rs:621-624 applies FilterChain::apply before diff_frame).  The CUDA chain must equal the numpy
oracle (oracle/resize_oracle.py) bit for bit: a handle WITH filters fed raw source frames gives the records and tables
of a plain handle fed oracle-filtered frames.  (Parity with video-resize 0.2.0 itself is unpinned: crate absent.)"""
import numpy as np
import pytest

from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from grav1synth_b200.synth import SynthSpec, make_pair_numpy
from oracle import resize_oracle as R

pytestmark = pytest.mark.gpu


def upscaled_source(spec, k, raw_w, raw_h, pad=(0, 0, 0, 0)):
    """A raw source of raw_w x raw_h whose filtered version is comparable to the denoised frame: the synthetic source
    resized up with numpy (any deterministic content does -- the test compares two paths over the same data)."""
    s, d = make_pair_numpy(spec, k)
    ss = (spec.ss_x, spec.ss_y)
    raw = R.resize_planes(s, raw_w, raw_h, "catmullrom", spec.bit_depth, ss)
    if any(pad):
        t, b, l, r = pad
        raw = [np.pad(raw[0], ((t, b), (l, r)), mode="edge")] + \
              [np.pad(p, ((t >> ss[1], b >> ss[1]), (l >> ss[0], r >> ss[0])), mode="edge") for p in raw[1:]]
    return [np.ascontiguousarray(p) for p in raw], d


def run(spec, frames, ops=None, raw_size=None, **kw):
    g = D.DiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y, **kw)
    if ops:
        g.set_source_filters(ops, *raw_size)
    recs = []
    g.set_record_tap(lambda i, r: recs.append(r))
    for s, d in frames:
        g.diff_frame(s, d)
    segs = g.finish()
    rl = D.RecordLayout(g.num_blocks)
    return segs, [rl.unpack(r) for r in recs]


def device_tables(alg, src, dst):
    return D.resize_table(alg, src, dst)


CHAINS = [
    ("crop only", (8, 8), [("crop", 4, 8, 2, 6)], (0, 0), (4, 8, 2, 6)),
    ("downscale catmullrom", (8, 8), [("resize", None, None, "catmullrom")], (480, 264), (0, 0, 0, 0)),
    ("downscale lanczos 10-bit", (10, 10), [("resize", None, None, "lanczos")], (400, 300), (0, 0, 0, 0)),
    ("upscale spline36", (8, 8), [("resize", None, None, "spline36")], (200, 120), (0, 0, 0, 0)),
    ("crop then mitchell", (10, 10), [("crop", 2, 6, 4, 8), ("resize", None, None, "mitchell")], (352, 240), (2, 6, 4, 8)),
    ("hermite then crop", (8, 8), [("resize", None, None, "hermite"), ("crop", 0, 4, 0, 8)], (300, 200), (0, 0, 0, 0)),
]


@pytest.mark.parametrize("name,bd,ops,raw,pad", CHAINS, ids=[c[0] for c in CHAINS])
def test_device_chain_equals_the_numpy_oracle(name, bd, ops, raw, pad):
    W, H = 320, 176
    spec = SynthSpec(W, H, bd[0], textured=0.2, sigma0=1.0, sigma1=1.5, seed=21)
    ss = (1, 1)
    # size the chain so that it ends at W x H
    post = [o for o in ops[ops.index(next(o for o in ops if o[0] == "resize")) + 1:]] if any(o[0] == "resize" for o in ops) else []
    grow_w = sum(o[3] + o[4] for o in post if o[0] == "crop")
    grow_h = sum(o[1] + o[2] for o in post if o[0] == "crop")
    full_ops = [("resize", W + grow_w, H + grow_h, o[3]) if o[0] == "resize" else o for o in ops]
    if raw == (0, 0):
        raw = (W, H)
    frames_raw, frames_ref = [], []
    for k in range(3):
        if any(o[0] == "resize" for o in ops):
            s, d = upscaled_source(spec, k, raw[0], raw[1], pad)
        else:
            s0, d = make_pair_numpy(spec, k)
            t, b, l, r = pad
            s = [np.pad(s0[0], ((t, b), (l, r)), mode="edge"), np.pad(s0[1], ((t >> 1, b >> 1), (l >> 1, r >> 1)), mode="edge"),
                 np.pad(s0[2], ((t >> 1, b >> 1), (l >> 1, r >> 1)), mode="edge")]
            s = [np.ascontiguousarray(p) for p in s]
        frames_raw.append((s, d))
        filtered = R.apply_chain(s, full_ops, bd[0], ss, tables=device_tables)
        assert filtered[0].shape == (H, W)
        frames_ref.append((filtered, d))
    rh, rw = frames_raw[0][0][0].shape
    got_segs, got = run(spec, frames_raw, full_ops, (rw, rh))
    want_segs, want = run(spec, frames_ref)
    for a, b in zip(got, want):
        for key in ("gram", "nobs", "flat", "score", "rsum", "rsq", "luma_sum"):
            assert np.array_equal(a[key], b[key]), (name, key)
    assert got_segs == want_segs


def test_filter_errors():
    g = D.DiffGenerator(24, 1, 8, 8, 320, 176)
    with pytest.raises(ValueError, match="Luma dimensions do not match"):   # the chain must end at the handle's size
        g.set_source_filters([("resize", 300, 176, "lanczos")], 640, 352)  # (verify_dimensions_match in the reference)
    with pytest.raises(ValueError):
        g.set_source_filters([("crop", 400, 0, 0, 0)], 320, 176)
    g.set_source_filters([("crop", 0, 16, 0, 32)], 352, 192)
    rng = np.random.default_rng(0)
    src = [rng.integers(0, 255, (192, 352), dtype=np.uint8), rng.integers(0, 255, (96, 176), dtype=np.uint8),
           rng.integers(0, 255, (96, 176), dtype=np.uint8)]
    den = [p[: p.shape[0] - (16 >> (i > 0)), : p.shape[1] - (32 >> (i > 0))].copy() for i, p in enumerate(src)]
    g.diff_frame(src, den)
    with pytest.raises(ValueError):      # a source that is not the configured pre-filter size
        g.diff_frame(den, den)
