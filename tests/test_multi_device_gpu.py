"""Several GPUs behind ONE handle (`-m gpu`): g1s_diff_config.n_devices / device_ids.  The reference's
caller is one process with one DiffGenerator (/root/reference/src/main.rs:420-524); a multi-device handle keeps that
shape: batches of frames are dealt round-robin to the devices, every device runs the kernels and the per-frame half of
the model, the digests are folded in frame order.  The table must be the single-device table, segment cut and short
last batch included.  On a box with one GPU the same dealing / folding logic is exercised with that GPU listed several
times (the children are independent handles); with two or more GPUs the devices are distinct."""
import numpy as np
import pytest

from grav1synth_b200 import abi
from grav1synth_b200 import diff as D
from grav1synth_b200.synth import SynthSpec, make_pair_numpy

pytestmark = pytest.mark.gpu


def stream(n=23):
    a = SynthSpec(320, 208, 10, textured=0.2, sigma0=1.0, sigma1=1.5, seed=5)
    b = SynthSpec(320, 208, 10, textured=0.2, sigma0=2.5, sigma1=0.5, ar_strength=0.6, seed=6)
    return a, [make_pair_numpy(a if k < n // 2 else b, k) for k in range(n)]


def devices_for_test(k):
    import torch
    n = torch.cuda.device_count()
    return [i % n for i in range(k)]


@pytest.mark.parametrize("ndev,batch", [(2, 3), (3, 2), (4, 5)])
def test_multi_device_handle_gives_the_single_device_table(ndev, batch):
    spec, frames = stream()
    one = D.DiffGenerator(30000, 1001, 10, 10, spec.width, spec.height, batch_frames=batch)
    for s, d in frames:
        one.diff_frame(s, d)
    want = one.finish()
    assert len(want) >= 2
    devs = devices_for_test(ndev)
    many = D.DiffGenerator(30000, 1001, 10, 10, spec.width, spec.height, batch_frames=batch, devices=devs)
    assert many.batch_frames == batch
    assert [many.frame_device(k) for k in range(0, 4 * batch, batch)] == [devs[i % ndev] for i in range(4)]
    for s, d in frames:
        many.diff_frame(s, d)
    assert many.frames_pushed == len(frames)
    got = many.finish()
    assert got == want
    c = many.counters()
    assert c["frames_done"] == len(frames) and c["kernels_launched"] > 0


def test_multi_device_strict_mode():
    spec, frames = stream(9)
    one = D.DiffGenerator(30000, 1001, 10, 10, spec.width, spec.height, gram_order=abi.GRAM_REF_ORDER)
    many = D.DiffGenerator(30000, 1001, 10, 10, spec.width, spec.height, batch_frames=2, devices=devices_for_test(2),
                           gram_order=abi.GRAM_REF_ORDER)
    for s, d in frames:
        one.diff_frame(s, d)
        many.diff_frame(s, d)
    assert many.finish() == one.finish()


def test_multi_device_rejects_bad_configurations():
    with pytest.raises(D.G1SError) as e:
        D.DiffGenerator(24, 1, 8, 8, 64, 64, devices=[0, 99])
    assert e.value.code in (abi.G1S_E_NCCL, abi.G1S_E_CUDA)
    with pytest.raises(D.G1SError) as e:
        D.DiffGenerator(24, 1, 8, 8, 64, 64, devices=[0, 0], mode=abi.MODE_PRODUCER)
    assert e.value.code == abi.G1S_E_ARG
