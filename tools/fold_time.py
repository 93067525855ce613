import sys, time, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from helpers import corpus_frames, gram_to_pairs, numpy_record
from grav1synth_b200 import diff as D, abi
from oracle import oracle as O
name = "c3_small_10bit"
spec, fps, frames = corpus_frames(name)
o = O.OracleDiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, ss_x=spec.ss_x, ss_y=spec.ss_y)
rl = D.RecordLayout(((spec.width + 31) // 32) * ((spec.height + 31) // 32))
helper = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y, mode=abi.MODE_CONSUMER)
dgs = []
for s, d in frames:
    o.diff_frame(s, d)
    flat, scores, _ = o.last_flat()
    r = numpy_record(s, d, spec.bit_depth, spec.bit_depth, spec.ss_x, spec.ss_y, flat)
    pairs = np.stack([gram_to_pairs(r["gram"][c]) for c in range(3)])
    rec = rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat)
    dgs.append(helper.digest_from_record(rec))
N = 30000
buf = np.stack([dgs[i % len(dgs)] for i in range(N)])
cons = D.DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y, mode=abi.MODE_CONSUMER)
cons.consume_digests(buf.ctypes.data, 656, borrowed=True)
cons.flush()
t0 = time.perf_counter()
CH = 656  # what rank 0 of an 8-GPU run receives per exchange (8 ranks x 82 frames)
for k in range(0, N, CH):
    cons.consume_digests(buf[k:].ctypes.data, min(CH, N - k), borrowed=True)
cons.flush()
dt = time.perf_counter() - t0
print(f"{N} digests folded in {dt:.3f} s: {dt / N * 1e6:.2f} us per frame, {N / dt:.0f} frames/s")
tbl = cons.finish()
print(len(tbl), "segments")
