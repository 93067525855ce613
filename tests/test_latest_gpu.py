"""latest_kernel (`-m gpu`): the per-frame half of the host model evaluated on the device (csrc/g1s_latest.cu) against
the host's NoiseModel::compute_latest on the same record (g1s_diff_digest_from_record): every double of the digest --
AR systems and solutions, gains, strength systems and solutions of the three channels -- must have the same BITS."""
import numpy as np
import pytest

from aom_cases import CASES
from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from test_aom_pin import load_or_skip

pytestmark = pytest.mark.gpu


def device_digests(frames, bd, ss, fps, **kw):
    h, w = frames[0][0][0].shape
    # a PRODUCER handle with a digest sink hands out exactly what the device computed; the record tap keeps the records
    g = D.DiffGenerator(fps[0], fps[1], bd, bd, w, h, ss[0], ss[1], mode=abi.MODE_PRODUCER,
                        model_placement=abi.MODEL_DEVICE, **kw)
    sink = np.zeros((len(frames), D.digest_bytes() // 8))
    g.set_digest_sink(sink.ctypes.data, len(frames))
    recs = []
    g.set_record_tap(lambda i, r: recs.append(r))
    for s, d in frames:
        g.diff_frame(s, d)
    g.flush()
    assert g.digest_count == len(frames)
    host = np.stack([g.digest_from_record(r) for r in recs])
    g.close()
    return sink, host


@pytest.mark.parametrize("name", ["c2_small_8bit", "c3_small_10bit", "odd_size_8bit", "tiny_64x48", "yuv444_8bit",
                                  "yuv422_10bit", "heavy_grain_12bit", "segment_cut", "saturated_residual",
                                  "flat_everything", "single_block", "zero_frame_mid_stream", "random_3", "random_6",
                                  "hd_1080p_frame", "uhd_4k_10bit_frame"])
@pytest.mark.parametrize("strict", [False, True], ids=["exact_int", "ref_order"])
def test_device_digest_has_the_hosts_bits(name, strict):
    if strict and name in ("hd_1080p_frame", "uhd_4k_10bit_frame"):
        pytest.skip("strict mode on the large frames is covered by test_gpu_strict")
    frames, bd, ss, fps = load_or_skip(name)
    dev, host = device_digests(frames, bd, ss, fps, gram_order=abi.GRAM_REF_ORDER if strict else abi.GRAM_EXACT_INT)
    for k in range(len(frames)):
        same = dev[k].view(np.uint64) == host[k].view(np.uint64)
        both_nan = np.isnan(dev[k]) & np.isnan(host[k])
        bad = np.flatnonzero(~(same | both_nan))
        assert bad.size == 0, (name, k, bad[:8], dev[k][bad[:8]], host[k][bad[:8]])


def test_host_model_switch_gives_the_same_table(monkeypatch):
    frames, bd, ss, fps = load_or_skip("segment_cut")
    h, w = frames[0][0][0].shape
    tables = []
    for place in (abi.MODEL_DEVICE, abi.MODEL_HOST):
        g = D.DiffGenerator(fps[0], fps[1], bd, bd, w, h, ss[0], ss[1], model_placement=place)
        for s, d in frames:
            g.diff_frame(s, d)
        tables.append(g.finish())
    assert tables[0] == tables[1] and len(tables[0]) >= 2
