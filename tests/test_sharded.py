"""world_size-2 gloo test (CPU) of the frame-sharded driver: dealing, record gather, in-order fold.
The CUDA producer is replaced by a CPU record producer (oracle flat mask + numpy sums); the
exchange and consumer code under test is the code the GPU run uses."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, corpus_frames, gram_to_pairs, numpy_record


class CpuRecordProducer:
    """Stand-in for the CUDA producer: per-frame record from the oracle's flat mask + numpy sums, turned into a
    digest by the product's own per-frame model code (g1s_diff_digest_from_record), written into the sink."""

    def __init__(self, spec, fps):
        import ctypes as C
        from grav1synth_b200 import abi
        from grav1synth_b200.diff import DiffGenerator, RecordLayout, digest_bytes
        from oracle import oracle as O
        self.spec = spec
        self.o = O.OracleDiffGenerator(24, 1, spec.bit_depth, spec.bit_depth, ss_x=spec.ss_x, ss_y=spec.ss_y)
        self.rl = RecordLayout(((spec.width + 31) // 32) * ((spec.height + 31) // 32))
        self.helper = DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x,
                                    spec.ss_y, mode=abi.MODE_CONSUMER)
        self.pending, self.sink, self.cap, self.count, self.C = [], None, 0, 0, C
        self.ndbl = digest_bytes() // 8

    def set_digest_sink(self, ptr, cap):
        self.sink, self.cap, self.count = ptr, cap, 0

    @property
    def digest_count(self):
        return self.count

    def diff_frame(self, s, d):
        spec = self.spec
        self.o.diff_frame(s, d)
        flat, scores, _ = self.o.last_flat()
        r = numpy_record(s, d, spec.bit_depth, spec.bit_depth, spec.ss_x, spec.ss_y, flat)
        pairs = np.stack([gram_to_pairs(r["gram"][c]) for c in range(3)])
        self.pending.append(self.rl.pack(pairs, r["nobs"], r["num_flat"], r["luma_sum"], r["rsum"], r["rsq"], scores, flat))

    def flush(self):
        for rec in self.pending:
            dg = self.helper.digest_from_record(rec)
            dst = (self.C.c_double * self.ndbl).from_address(self.sink + 8 * self.ndbl * (self.count % self.cap))
            np.frombuffer(dst, np.float64)[:] = dg
            self.count += 1
        self.pending.clear()

    def wait_retired(self, frames):
        self.flush()


def _worker(rank, world, port, name, B, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grav1synth_b200.diff import format_grain_table
    from grav1synth_b200.sharded import ShardedDiff, owner_of
    spec, fps, frames = corpus_frames(name)
    sd = ShardedDiff(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y,
                     frames_per_rank=B, producer_factory=lambda: CpuRecordProducer(spec, fps))
    n, base, folded = len(frames), 0, 0
    while base < n:
        for k in range(base, min(n, base + world * B)):
            if owner_of(k, world, B) == rank:
                sd.push_local(*frames[k])
        sd.exchange()
        base += world * B
    folded = sd.exchange(final=True)
    segs = sd.finish()
    if rank == 0:
        assert folded == n
        with open(out_path, "w") as f:
            f.write(format_grain_table(segs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,B", [("c2_small_8bit", 1), ("c2_small_8bit", 3), ("c3_small_10bit", 2)])
def test_two_rank_gloo_matches_golden(name, B, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "t.tbl")
    mp.spawn(_worker, args=(2, port, name, B, out), nprocs=2, join=True)
    with open(os.path.join(ROOT, "tests", "golden", name + ".tbl")) as f:
        assert open(out).read() == f.read()


def test_four_rank_gloo_long_stream_takes_the_batched_fold(tmp_path):
    """72 frames over 4 ranks, 9 per rank and exchange: rank 0 folds 36 digests at a time, i.e. through
    DiffSequencer::consume_latest_batch / NoiseModel::fold_run.  Expected: the same frames folded one by one."""
    import ctypes as C
    from grav1synth_b200 import abi
    from grav1synth_b200.diff import DiffGenerator, format_grain_table
    name = "tiny_long"
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "t.tbl")
    mp.spawn(_worker, args=(4, port, name, 9, out), nprocs=4, join=True)
    spec, fps, frames = corpus_frames(name)
    prod = CpuRecordProducer(spec, fps)
    ring = np.zeros((len(frames), prod.ndbl), np.float64)
    prod.set_digest_sink(ring.ctypes.data, len(frames))
    for sf, df in frames:
        prod.diff_frame(sf, df)
    prod.flush()
    one = DiffGenerator(fps[0], fps[1], spec.bit_depth, spec.bit_depth, spec.width, spec.height, spec.ss_x, spec.ss_y,
                        mode=abi.MODE_CONSUMER)
    for k in range(len(frames)):
        one.consume_digests(ring[k:].ctypes.data, 1)
    assert open(out).read() == format_grain_table(one.finish())


def test_owner_dealing():
    from grav1synth_b200.sharded import owner_of
    assert [owner_of(k, 2, 2) for k in range(8)] == [0, 0, 1, 1, 0, 0, 1, 1]
    assert [owner_of(k, 4, 1) for k in range(6)] == [0, 1, 2, 3, 0, 1]
