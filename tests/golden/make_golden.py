"""Regenerates tests/golden/*.tbl from the CPU oracle on the seeded synthetic corpus.

These are RESTATEMENT goldens (oracle/g1s_oracle.c, EXACT_INT + fixed exp), not outputs of the
reference binary: the reference cannot be built here (no Rust toolchain; the arithmetic lives in the
un-vendored crate av1-grain 0.4.2) and it ships no golden for `diff`.  If a grav1synth binary ever
becomes available, regenerate with it and diff.

    python tests/golden/make_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import CORPUS, corpus_frames  # noqa: E402
from oracle import oracle as O  # noqa: E402

for name in CORPUS:
    spec, fps, frames = corpus_frames(name)
    g = O.OracleDiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, O.GRAM_EXACT_INT, O.EXP_FIXED,
                              spec.ss_x, spec.ss_y)
    for s, d in frames:
        g.diff_frame(s, d)
    O.write_grain_table(g.finish(), os.path.join(HERE, name + ".tbl"))
    print("wrote", name)
