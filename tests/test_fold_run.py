"""DiffSequencer::consume_latest_batch / NoiseModel::fold_run (the sequential half of the model with the solves of the
combined state handed to helper threads) against the plain frame-by-frame fold, on streams with scene cuts, frames
without enough flat blocks and frames whose evaluation failed: the tables must be identical to the last bit."""
import numpy as np
import pytest

from helpers import gram_to_pairs, numpy_record, table_text
from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from grav1synth_b200.synth import SynthSpec, make_pair_numpy
from oracle import oracle as O

W, H, BD = 256, 160, 8


def scene_digests(sigma0, sigma1, seed, n):
    spec = SynthSpec(W, H, BD, textured=0.1, sigma0=sigma0, sigma1=sigma1, seed=seed)
    o = O.OracleDiffGenerator(24, 1, BD, BD, ss_x=1, ss_y=1)
    rl = D.RecordLayout(((W + 31) // 32) * ((H + 31) // 32))
    helper = D.DiffGenerator(24, 1, BD, BD, W, H, 1, 1, mode=abi.MODE_CONSUMER)
    out = []
    for k in range(n):
        s, d = make_pair_numpy(spec, k)
        o.diff_frame(s, d)
        flat, scores, _ = o.last_flat()
        r = numpy_record(s, d, BD, BD, 1, 1, flat)
        pairs = np.stack([gram_to_pairs(r["gram"][c]) for c in range(3)])
        rec = rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat)
        out.append(helper.digest_from_record(rec))
    helper.close()
    return out


@pytest.fixture(scope="module")
def scenes():
    return {"A": scene_digests(1.0, 1.5, 21, 4), "B": scene_digests(4.0, 7.0, 22, 4), "C": scene_digests(2.0, 0.5, 23, 3)}


def fold(stream, chunk):
    g = D.DiffGenerator(24, 1, BD, BD, W, H, 1, 1, mode=abi.MODE_CONSUMER)
    buf = np.ascontiguousarray(np.stack(stream))
    for k in range(0, len(stream), chunk):
        n = min(chunk, len(stream) - k)
        g.consume_digests(buf[k:].ctypes.data, n)
    segs = g.finish()
    g.close()
    return segs


def same_tables(a, b):
    assert len(a) == len(b)
    assert table_text(a) == table_text(b)
    for x, y in zip(a, b):
        assert x == y


def build(scenes, plan):
    stream = []
    for name, n in plan:
        if name == "noflat":
            d = scenes["A"][0].copy()
            d[0] = 0.0   # enough_flat = false: NoiseStatus::Error, nothing merged
            stream += [d] * n
        elif name == "fail1":
            d = scenes["B"][1].copy()
            d[1], d[2], d[3] = 2.0, 1.0, 2.0   # Cb's strength solve failed: luma merged, chroma not
            stream += [d] * n
        else:
            src = scenes[name]
            stream += [src[k % len(src)] for k in range(n)]
    return stream


PLANS = {
    "one_scene": [("A", 70)],
    "cuts": [("A", 45), ("B", 60), ("C", 33), ("A", 17)],
    "cut_inside_the_first_sixteen": [("A", 5), ("B", 40), ("A", 3), ("C", 30)],
    "frames_that_fail": [("A", 30), ("noflat", 1), ("A", 25), ("B", 20), ("fail1", 2), ("B", 30), ("noflat", 3), ("C", 20)],
    "short_runs_only": [("A", 10), ("noflat", 1), ("A", 12), ("B", 9), ("noflat", 1), ("B", 15)],
}


@pytest.mark.parametrize("plan", sorted(PLANS))
def test_batched_fold_equals_the_frame_by_frame_fold(scenes, plan):
    stream = build(scenes, PLANS[plan])
    want = fold(stream, 1)            # one digest per call: every frame takes NoiseModel::fold
    assert len(want) >= (3 if plan == "cuts" else 1)
    for chunk in (len(stream), 64, 37, 16):
        same_tables(fold(stream, chunk), want)


def test_batched_fold_with_more_helpers_than_frames_and_with_none(scenes, monkeypatch):
    stream = build(scenes, PLANS["cuts"])
    want = fold(stream, 1)
    for helpers in ("1", "3", "16"):
        monkeypatch.setenv("G1S_FOLD_THREADS", helpers)
        same_tables(fold(stream, len(stream)), want)


def test_records_in_batches_equal_records_one_by_one():
    """The single-GPU route: records -> per-frame model half on the host pool -> fold.  72 frames handed over in one
    call (the batched fold) against one call per frame."""
    from helpers import corpus_frames
    spec, fps, frames = corpus_frames("tiny_long")
    o = O.OracleDiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, ss_x=spec.ss_x, ss_y=spec.ss_y)
    rl = D.RecordLayout(((spec.width + 31) // 32) * ((spec.height + 31) // 32))
    recs = []
    for s, d in frames:
        o.diff_frame(s, d)
        flat, scores, _ = o.last_flat()
        r = numpy_record(s, d, spec.bit_depth, spec.bit_depth, spec.ss_x, spec.ss_y, flat)
        pairs = np.stack([gram_to_pairs(r["gram"][c]) for c in range(3)])
        recs.append(rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat))
    recs = np.stack(recs)

    def run(chunk):
        g = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y,
                            mode=abi.MODE_CONSUMER)
        for k in range(0, len(recs), chunk):
            g.consume_records(recs[k:k + chunk])
        segs = g.finish()
        g.close()
        return segs

    want = run(1)
    assert want == o.finish()      # and that is the oracle's table
    for chunk in (len(recs), 40, 16):
        same_tables(run(chunk), want)
