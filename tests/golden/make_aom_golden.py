"""Writes tests/golden/aom/*.json: what libaom 3.13.1's own noise_model.c (the upstream of av1-grain's `diff`,
run from the binary bundled with opencv-python-headless, see oracle/aom_pin.py) returns for every case of
tests/aom_cases.py -- per frame the status, the flat-block map and the AR / strength solutions (exact f64 bit
patterns), per segment every integer of the grain parameters.

These are UPSTREAM-BINARY goldens (not restatement goldens): tests compare the oracle and the CUDA engine with
them whether or not libaom is present at test time.

    python tests/golden/make_aom_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from aom_cases import CASES, load  # noqa: E402
from oracle import aom_pin as P  # noqa: E402

ok, where = P.available()
if not ok:
    sys.exit("libaom pin unavailable: " + where)
os.makedirs(os.path.join(HERE, "aom"), exist_ok=True)
# the codec pair's stream: libaom's encode of aom_cases.codec_source() at constant quality 28, no film grain
_ivf = os.path.join(HERE, "aom", "codec_pair_cq28.ivf")
if not os.path.exists(_ivf):
    import av1_writer as W
    from aom_cases import codec_source
    from oracle import aom_encode as E
    _pk = E.encode([tuple(f) for f in codec_source()], 352, 288, {"cq-level": "28", "cpu-used": "6"}, lag_in_frames=0,
                   cfg_words={24: 3})
    with open(_ivf, "wb") as f:
        f.write(W.ivf(_pk, 352, 288, 24, 1))
for name in (sys.argv[1:] or CASES):
    frames, bd, ss, fps = load(name)
    a = P.AomNoiseModel(ss[0], ss[1])
    per_frame = []
    for s, d in frames:
        st = a.update([P.to_u8(p, bd) for p in s], [P.to_u8(p, bd) for p in d])
        states = {}
        for which in ("latest", "combined"):
            for c in range(3):
                stt = a.state(which, c)
                h = hashlib.sha256()
                h.update(stt["x"].tobytes()); h.update(stt["strength_x"].tobytes())
                states[f"{which}{c}"] = dict(nobs=stt["num_observations"], ar_gain=stt["ar_gain"].hex(),
                                             digest=h.hexdigest()[:16])
        per_frame.append(dict(status=st, num_flat=int(a.num_flat),
                              flat_sha=hashlib.sha256(a.flat.tobytes()).hexdigest()[:16], state=states))
    segs = a.finish()
    out = dict(case=name, producer=os.path.basename(where), bit_depth=bd, ss=list(ss), frames=per_frame,
               segment_first_frame=a.segment_first_frame, segments=segs)
    with open(os.path.join(HERE, "aom", name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", name, "segments", len(segs))
