"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI, against the CPU oracle.

Bit-exact bar: flat flags, f32 flatness scores (bit patterns), integer Gram sums, observation
counts, per-block noise statistics and the final grain tables must all be identical.
Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

from helpers import CORPUS, ROOT, corpus_frames, gram_to_pairs, numpy_record
from grav1synth_b200 import diff as D
from grav1synth_b200.synth import SynthSpec, make_pair_numpy
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def gpu_run(spec, fps, frames, batch=0, src_bd=None, den_bd=None, gram_kernel=0):
    g = D.DiffGenerator(fps[0], fps[1], src_bd or spec.bit_depth, den_bd or spec.bit_depth, spec.width, spec.height,
                        spec.ss_x, spec.ss_y, batch_frames=batch, gram_kernel=gram_kernel)
    recs = []
    g.set_record_tap(lambda i, r: recs.append((i, r)))
    for s, d in frames:
        g.diff_frame(s, d)
    segs = g.finish()
    assert [i for i, _ in recs] == list(range(len(frames)))
    rl = D.RecordLayout(g.num_blocks)
    return segs, [rl.unpack(r) for _, r in recs], g


def oracle_run(spec, fps, frames, src_bd=None, den_bd=None):
    o = O.OracleDiffGenerator(fps[0], fps[1], src_bd or spec.bit_depth, den_bd or spec.bit_depth, O.GRAM_EXACT_INT,
                              O.EXP_FIXED, spec.ss_x, spec.ss_y)
    per = []
    for s, d in frames:
        o.diff_frame(s, d)
        flat, scores, feat = o.last_flat()
        per.append(dict(flat=flat, scores=scores, gram=[o.last_gram(c) for c in range(3)], status=o.last_status))
    return o.finish(), per


def compare_records(spec, frames, got, want, src_bd=None, den_bd=None):
    for k, ((s, d), g, w) in enumerate(zip(frames, got, want)):
        assert np.array_equal(g["flat"], w["flat"]), f"frame {k}: flat flags differ"
        assert np.array_equal(g["score"].view(np.uint32), w["scores"].view(np.uint32)), f"frame {k}: scores differ"
        assert g["num_flat"] == int((w["flat"] != 0).sum())
        if g["num_flat"] <= 1:
            continue
        ref = numpy_record(s, d, src_bd or spec.bit_depth, den_bd or spec.bit_depth, spec.ss_x, spec.ss_y, w["flat"])
        for c in range(3):
            G, nobs = w["gram"][c]
            assert int(g["nobs"][c]) == nobs, f"frame {k} plane {c}: nobs"
            assert np.array_equal(D.gram_pairs_to_matrix(g["gram"][c]), G), f"frame {k} plane {c}: Gram"
            m = w["flat"] != 0
            assert np.array_equal(g["rsum"][c][m], ref["rsum"][c][m])
            assert np.array_equal(g["rsq"][c][m], ref["rsq"][c][m])
        assert np.array_equal(g["luma_sum"][w["flat"] != 0], ref["luma_sum"][w["flat"] != 0])


@pytest.mark.parametrize("path", ["tensorcore-vec", "tensorcore-scalar", "generic"])
@pytest.mark.parametrize("name", list(CORPUS))
def test_corpus_bit_exact(name, path, monkeypatch):
    """Every Gram path against the oracle: residual kernel (128-bit loads) + TMA-fed int8 tensor-core Gram
    kernel, the same with the residual kernel's scalar loads (what unaligned planes get), and the generic
    int32 kernel."""
    if path == "tensorcore-scalar":
        monkeypatch.setenv("G1S_SCALAR_LOADS", "1")
    spec, fps, frames = corpus_frames(name)
    segs, recs, g = gpu_run(spec, fps, frames, gram_kernel=1 if path == "generic" else 0)
    if spec.ss_x == 1 and spec.ss_y == 1:
        c = g.counters()
        assert (c["tma_batches"] > 0) == (path != "generic")
        assert (c["vector_batches"] > 0) == (path == "tensorcore-vec")
    want, per = oracle_run(spec, fps, frames)
    compare_records(spec, frames, recs, per)
    assert segs == want
    with open(os.path.join(ROOT, "tests", "golden", name + ".tbl")) as f:
        assert D.format_grain_table(segs) == f.read()


@pytest.mark.parametrize("batch", [1, 2, 3])
def test_batching_does_not_change_results(batch):
    spec, fps, frames = corpus_frames("c2_small_8bit")
    a, ra, _ = gpu_run(spec, fps, frames, batch=batch)
    b, rb, _ = gpu_run(spec, fps, frames, batch=0)
    assert a == b
    for x, y in zip(ra, rb):
        for key in ("gram", "nobs", "flat", "score", "rsum", "rsq", "luma_sum"):
            assert np.array_equal(x[key], y[key])


def test_mixed_bit_depths():
    spec, fps, frames = corpus_frames("c3_small_10bit")
    frames = [(s, [(p >> 2).astype(np.uint8) for p in d]) for s, d in frames]
    segs, recs, _ = gpu_run(spec, fps, frames, src_bd=10, den_bd=8)
    want, per = oracle_run(spec, fps, frames, src_bd=10, den_bd=8)
    compare_records(spec, frames, recs, per, 10, 8)
    assert segs == want


def test_device_resident_frames_equal_host_frames():
    import torch
    spec, fps, frames = corpus_frames("c3_small_10bit")
    host, _, _ = gpu_run(spec, fps, frames)
    g = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y)
    keep = []
    for s, d in frames:
        ts = [torch.from_numpy(p.view(np.int16)).cuda() for p in s]
        td = [torch.from_numpy(p.view(np.int16)).cuda() for p in d]
        keep += ts + td
        g.diff_frame_device([t.data_ptr() for t in ts], [t.stride(0) * 2 for t in ts],
                            [t.data_ptr() for t in td], [t.stride(0) * 2 for t in td])
    torch.cuda.synchronize()
    assert g.finish() == host


def test_saturated_residual_and_zero_residual():
    # |r| = 255 everywhere in one half (int8 overflow territory for any packed path), zero in the other
    rng = np.random.default_rng(5)
    h, w = 128, 192
    den = [np.zeros((h, w), np.uint8), np.zeros((h // 2, w // 2), np.uint8), np.full((h // 2, w // 2), 255, np.uint8)]
    src = [np.full((h, w), 255, np.uint8), np.full((h // 2, w // 2), 255, np.uint8), np.zeros((h // 2, w // 2), np.uint8)]
    src[0][:, : w // 2] = rng.integers(0, 256, (h, w // 2), dtype=np.uint8)
    den[0][:, : w // 2] = rng.integers(0, 256, (h, w // 2), dtype=np.uint8)
    spec = SynthSpec(w, h, 8)
    frames = [(src, den), (den, src)]
    segs, recs, _ = gpu_run(spec, (24, 1), frames)
    want, per = oracle_run(spec, (24, 1), frames)
    compare_records(spec, frames, recs, per)
    assert segs == want


def test_sparse_int8_overflow_falls_back_per_block():
    # ordinary grain everywhere, plus a few isolated |r| > 127 samples: only the touched super-units may take
    # the generic kernel, and the sums must still be exact
    spec, fps, frames = corpus_frames("c2_small_8bit")
    frames = [([p.copy() for p in s], [p.copy() for p in d]) for s, d in frames]
    for k, (s, d) in enumerate(frames):
        s[0][37 + k, 70], d[0][37 + k, 70] = 255, 3       # luma, +252
        s[1][50, 101], d[1][50, 101] = 0, 200             # Cb, -200
        s[2][5, 3], d[2][5, 3] = 129, 0                   # Cr, +129 (just outside int8)
        s[0][100, 200], d[0][100, 200] = 0, 128           # luma, -128 (still inside int8)
    segs, recs, _ = gpu_run(spec, fps, frames)
    want, per = oracle_run(spec, fps, frames)
    compare_records(spec, frames, recs, per)
    assert segs == want


def test_unaligned_device_planes():
    # device planes whose base address / stride are not 8-byte aligned take the scalar-load variant
    import torch
    spec, fps, frames = corpus_frames("c2_small_8bit")
    host, _, _ = gpu_run(spec, fps, frames)
    g = D.DiffGenerator(fps[0], fps[1], 8, 8, spec.width, spec.height)
    keep = []

    def odd(p):
        h, w = p.shape
        buf = torch.zeros((h, w + 13), dtype=torch.uint8, device="cuda")
        buf[:, 3:3 + w] = torch.from_numpy(p).cuda()
        keep.append(buf)
        return buf.data_ptr() + 3, w + 13

    for s, d in frames:
        sp, dp = [odd(p) for p in s], [odd(p) for p in d]
        g.diff_frame_device([a for a, _ in sp], [b for _, b in sp], [a for a, _ in dp], [b for _, b in dp])
    torch.cuda.synchronize()
    assert g.finish() == host


def test_flat_everything_rule_and_error_swallowing():
    y = np.full((96, 128), 100, np.uint8)
    c = np.full((48, 64), 128, np.uint8)
    spec = SynthSpec(128, 96, 8)
    segs, recs, _ = gpu_run(spec, (24, 1), [([y, c, c], [y, c, c])])
    assert np.all(recs[0]["flat"] == 1) and np.all(recs[0]["score"] == 0)
    want, _ = oracle_run(spec, (24, 1), [([y, c, c], [y, c, c])])
    assert segs == want


def test_single_block_frame_not_enough_flat_blocks():
    rng = np.random.default_rng(0)
    y = rng.integers(0, 255, (32, 32), dtype=np.uint8)
    c = np.full((16, 16), 128, np.uint8)
    spec = SynthSpec(32, 32, 8)
    segs, recs, _ = gpu_run(spec, (24, 1), [([y, c, c], [y, c, c])])
    want, _ = oracle_run(spec, (24, 1), [([y, c, c], [y, c, c])])
    assert recs[0]["num_flat"] <= 1 and segs == want


def test_dimension_mismatch_raises():
    g = D.DiffGenerator(24, 1, 8, 8, 64, 64)
    a = [np.zeros((64, 64), np.uint8), np.zeros((32, 32), np.uint8), np.zeros((32, 32), np.uint8)]
    b = [np.zeros((64, 96), np.uint8), np.zeros((32, 48), np.uint8), np.zeros((32, 48), np.uint8)]
    with pytest.raises(ValueError):
        g.diff_frame(a, b)


def test_producer_digests_fold_to_the_same_table():
    """PRODUCER handle (kernels + per-frame model half, digests into a sink) -> CONSUMER handle: the
    multi-GPU data flow on one GPU must reproduce the FULL handle's table."""
    import torch
    from grav1synth_b200 import abi
    spec, fps, frames = corpus_frames("c3_small_10bit")
    full, _, _ = gpu_run(spec, fps, frames)
    args = (fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y)
    prod = D.DiffGenerator(*args, mode=abi.MODE_PRODUCER, batch_frames=2)
    cons = D.DiffGenerator(*args, mode=abi.MODE_CONSUMER)
    sink = torch.zeros((len(frames), D.digest_bytes() // 8), dtype=torch.float64).pin_memory()
    prod.set_digest_sink(sink.data_ptr(), len(frames))
    for s, d in frames:
        prod.diff_frame(s, d)
    prod.flush()
    assert prod.digest_count == len(frames)
    cons.consume_digests(sink.data_ptr(), len(frames))
    assert cons.finish() == full


def test_pinned_host_planes_take_the_direct_path():
    import torch
    spec, fps, frames = corpus_frames("c3_small_10bit")
    want, _, _ = gpu_run(spec, fps, frames)
    g = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y)
    keep = []
    for s, d in frames:
        ps = [torch.from_numpy(p.view(np.int16)).pin_memory() for p in s]
        pd = [torch.from_numpy(p.view(np.int16)).pin_memory() for p in d]
        keep += ps + pd
        g.diff_frame([t.numpy().view(np.uint16) for t in ps], [t.numpy().view(np.uint16) for t in pd])
    assert g.finish() == want


def test_8k_two_scenes_cut_a_segment():
    """BASELINE configs[4] geometry (7680x4320 10-bit 4:2:0): two frames with different grain must give the
    oracle's table, segment cut included."""
    a = SynthSpec(7680, 4320, 10, textured=0.1, sigma0=1.0, sigma1=0.5, ar_strength=0.0, seed=41)
    b = SynthSpec(7680, 4320, 10, textured=0.1, sigma0=2.0, sigma1=0.5, ar_strength=0.6, seed=42)
    frames = [make_pair_numpy(a, 0), make_pair_numpy(b, 0)]
    segs, recs, _ = gpu_run(a, (60, 1), frames)
    want, per = oracle_run(a, (60, 1), frames)
    assert np.array_equal(recs[0]["flat"], per[0]["flat"]) and np.array_equal(recs[1]["flat"], per[1]["flat"])
    for c in range(3):
        for k in range(2):
            assert np.array_equal(D.gram_pairs_to_matrix(recs[k]["gram"][c]), per[k]["gram"][c][0])
    assert segs == want and len(segs) == 2 and segs[0].end_time == 10_000_000 // 60


@pytest.mark.parametrize("seed", range(8))
def test_random_geometries(seed):
    """Random sizes (odd, not multiples of 32/64, narrow last block columns), bit depths and texture fractions:
    every record field against the oracle.  Exercises super-unit edges, the TMA zero fill and the margin logic."""
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.integers(40, 420))
    h = int(rng.integers(40, 260))
    bd = int(rng.choice([8, 10, 12]))
    spec = SynthSpec(w, h, bd, textured=float(rng.choice([0.0, 0.3, 0.6, 0.9])), sigma0=float(rng.uniform(0.8, 2.5)),
                     sigma1=float(rng.uniform(0.0, 2.0)), ar_strength=float(rng.uniform(0, 0.5)), seed=seed)
    frames = [make_pair_numpy(spec, k) for k in range(2)]
    segs, recs, _ = gpu_run(spec, (24, 1), frames)
    want, per = oracle_run(spec, (24, 1), frames)
    compare_records(spec, frames, recs, per)
    assert segs == want


def test_monochrome():
    spec, fps, frames = corpus_frames("c2_small_8bit")
    g = D.DiffGenerator(fps[0], fps[1], 8, 8, spec.width, spec.height, monochrome=True)
    o = O.OracleDiffGenerator(fps[0], fps[1], 8, 8)
    for s, d in frames:
        g.diff_frame(s[:1], d[:1])
        o.diff_frame(s[:1], d[:1])
    assert g.finish() == o.finish()


def test_segment_cut_on_gpu():
    a = SynthSpec(256, 192, 8, textured=0.0, sigma0=1.0, sigma1=0.5, ar_strength=0.0, seed=1)
    b = SynthSpec(256, 192, 8, textured=0.0, sigma0=2.2, sigma1=0.5, ar_strength=0.6, seed=2)
    frames = [make_pair_numpy(a, k) for k in range(3)] + [make_pair_numpy(b, k) for k in range(3)]
    segs, _, _ = gpu_run(a, (30000, 1001), frames, batch=2)
    want, _ = oracle_run(a, (30000, 1001), frames)
    assert len(want) >= 2 and segs == want


def test_full_size_4k_10bit_frame_against_oracle():
    """BASELINE configs[2] geometry: one 3840x2160 10-bit 4:2:0 pair, full record vs the oracle."""
    spec = SynthSpec(3840, 2160, 10, textured=0.1, sigma0=1.0, sigma1=1.5, seed=2026)
    frames = [make_pair_numpy(spec, 0)]
    segs, recs, g = gpu_run(spec, (24, 1), frames)
    want, per = oracle_run(spec, (24, 1), frames)
    compare_records(spec, frames, recs, per)
    assert segs == want
    c = g.counters()
    assert c["gram_launches"] >= 1 and c["flat_launches"] >= 1 and c["frames_done"] == 1


def test_full_size_properties_1080p_batch():
    """BASELINE configs[1] geometry, 6 frames: determinism, frame-order independence of records, and
    the closed-form observation count implied by the flat mask."""
    from helpers import obs_rects
    spec = SynthSpec(1920, 1080, 8, textured=0.2, sigma0=1.0, sigma1=1.5, seed=99)
    frames = [make_pair_numpy(spec, k) for k in range(6)]
    _, r1, _ = gpu_run(spec, (24, 1), frames, batch=4)
    _, r2, _ = gpu_run(spec, (24, 1), frames[::-1], batch=6)
    for a, b in zip(r1, r2[::-1]):
        for key in ("gram", "nobs", "flat", "score", "rsum", "rsq", "luma_sum"):
            assert np.array_equal(a[key], b[key]), key
    nbw, nbh = 60, 34
    for rec in r1:
        for c in range(3):
            s = 1 if c else 0
            n = sum(max(0, x1 - x0) * max(0, y1 - y0) if (x1 > x0 and y1 > y0) else 0
                    for (_, _, x0, x1, y0, y1) in obs_rects(rec["flat"], nbw, nbh, 1920 >> s, 1080 >> s, 32 >> s, 32 >> s))
            assert int(rec["nobs"][c]) == n
        # centre-sample square sum over observations can never exceed the sum over whole flat blocks
        G = D.gram_pairs_to_matrix(rec["gram"][0])
        assert 0 < G[25, 25] <= int(rec["rsq"][0][rec["flat"] != 0].astype(np.int64).sum())


@pytest.mark.parametrize("name", ["c2_small_8bit", "c3_small_10bit", "odd_size_8bit", "yuv444_8bit", "yuv422_10bit",
                                  "heavy_grain_12bit", "saturated_residual", "sparse_int8_overflow",
                                  "zero_frame_mid_stream", "random_4", "random_5", "long_12_frames",
                                  "hd_1080p_frame", "uhd_4k_10bit_frame", "codec_pair_cq28"])
def test_engine_gives_libaom_tables(name):
    """The CUDA engine against the UPSTREAM BINARY: tests/golden/aom/*.json is what libaom 3.13.1's own
    noise_model.c (the code av1-grain's `diff` ports; oracle/aom_pin.py, tests/golden/make_aom_golden.py)
    returns for the same frames reduced to 8 bits -- flat-block maps per frame and every integer of every segment."""
    import hashlib
    import json
    from test_aom_pin import load_or_skip, seg_view
    with open(os.path.join(ROOT, "tests", "golden", "aom", name + ".json")) as f:
        want = json.load(f)
    frames, bd, ss, fps = load_or_skip(name)
    h, w = frames[0][0][0].shape
    spec = SynthSpec(w, h, bd, ss_x=ss[0], ss_y=ss[1])
    segs, recs, _ = gpu_run(spec, fps, frames)
    for k, (r, wf) in enumerate(zip(recs, want["frames"])):
        assert hashlib.sha256(np.ascontiguousarray(r["flat"]).tobytes()).hexdigest()[:16] == wf["flat_sha"], k
        if wf["status"] in (0, 3):
            for c in range(3):
                assert int(r["nobs"][c]) == wf["state"][f"latest{c}"]["nobs"], (k, c)
    assert [seg_view(s) for s in segs] == want["segments"]
