"""Frame-sharded multi-GPU driver (one process per GPU, torch.distributed for the plumbing).

The path shards by frame: every per-pixel quantity is a function of one frame
pair, and the only cross-frame logic is the small sequential model merge.  So each rank runs a
PRODUCER handle on its own frames: kernels, plus the per-frame half of the host model, whose result
is a fixed-size ~11 KB digest per frame written straight into a pinned ring.  Once per super-batch
the digests are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and rank 0 folds them
in global frame order into one CONSUMER handle (asynchronously, on the handle's own thread).
A super-batch is world*B consecutive frames; rank r owns frames [base + r*B, base + (r+1)*B), so the
rank-major gather result is already in frame order.  The gather of super-batch k is issued after the
frames of super-batch k+1 have been queued, so it overlaps the kernels.  There is no data-path
collective besides that gather.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import abi
from .diff import DiffGenerator, digest_bytes

RING = 4  # super-batches the digest ring holds


class ShardedDiff:
    def __init__(self, fps_num: int, fps_den: int, source_bit_depth: int, denoised_bit_depth: int, width: int,
                 height: int, ss_x: int = 1, ss_y: int = 1, frames_per_rank: int = 8, device: Optional[int] = None,
                 producer_factory: Optional[Callable[[], object]] = None, group=None, batch_frames: int = 0,
                 model_placement: Optional[int] = None):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.B = frames_per_rank
        self.ndbl = digest_bytes() // 8
        args = (fps_num, fps_den, source_bit_depth, denoised_bit_depth, width, height, ss_x, ss_y)
        if producer_factory is not None:          # tests inject a CPU digest producer
            self.producer = producer_factory()
        else:
            if model_placement is None:
                # one process per GPU on one host: about 3 cores per rank keep the per-frame model half on the host fed;
                # with fewer (8 GPUs on 32 cores) the device evaluates it (latest_kernel) and only digests come back
                import os
                model_placement = abi.MODEL_DEVICE if 3 * self.world > (os.cpu_count() or 2) // 2 else abi.MODEL_HOST
            import os
            self.producer = DiffGenerator(*args, device=device or 0, batch_frames=batch_frames,
                                          mode=abi.MODE_PRODUCER, model_placement=model_placement,
                                          host_threads=max(2, min(16, (os.cpu_count() or 2) // self.world)))
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self.on_gpu = backend == "nccl"
        # the producer writes digests straight into this ring (pinned when a GPU is involved)
        self.ring = torch.zeros((RING * self.B, self.ndbl), dtype=torch.float64)
        rows = self.B + 1  # + one row carrying the rank's frame count of the super-batch
        if self.on_gpu:
            self.ring = self.ring.pin_memory()
            dev = torch.device("cuda", torch.cuda.current_device())
            self.dev_local = torch.empty((rows, self.ndbl), dtype=torch.float64, device=dev)
            # two gather buffers: the device->host copy of exchange k runs while exchange k+1 is being gathered,
            # and rank 0's fold thread reads the pinned copy in place (no second copy on the Python thread)
            self.dev_all = [torch.empty((self.world * rows, self.ndbl), dtype=torch.float64, device=dev) for _ in range(2)]
            if self.rank == 0:
                self.host_all = [torch.empty((self.world * rows, self.ndbl), dtype=torch.float64).pin_memory()
                                 for _ in range(2)]
                self.copied = [torch.cuda.Event() for _ in range(2)]
            self.pending = None            # gather slot whose digests have not been handed to the consumer yet
            self.borrowed = [False, False]  # slot is (possibly still) being read by the consumer's fold thread
        self.producer.set_digest_sink(self.ring.data_ptr(), RING * self.B)
        self.consumer = DiffGenerator(*args, mode=abi.MODE_CONSUMER) if self.rank == 0 else None
        self.pushed = 0        # local frames pushed
        self.sb_pushed = []    # local frame count of every super-batch closed so far
        self.sb_done = 0       # super-batches exchanged
        self.folded = 0        # frames handed to the consumer (rank 0)

    # -- per super-batch --------------------------------------------------------------
    def push_local(self, source, denoised, device_resident: bool = False) -> None:
        """Queue one of this rank's frames of the current super-batch (at most B per super-batch)."""
        if device_resident:
            self.producer.diff_frame_device(*source, *denoised)
        else:
            self.producer.diff_frame(source, denoised)
        self.pushed += 1

    def _exchange_one(self, final: bool) -> None:
        k = self.sb_done
        n_local = self.sb_pushed[k]
        first = sum(self.sb_pushed[:k])
        if final:
            self.producer.flush()
        else:
            self.producer.wait_retired(first + n_local)
        if self.producer.digest_count < first + n_local:
            raise RuntimeError("digests of the super-batch are not complete")
        lo = (k % RING) * self.B
        local = self.ring[lo:lo + self.B]
        rows = self.B + 1
        if self.world > 1 and self.on_gpu:
            slot = k % 2
            if self.consumer is not None and self.borrowed[slot]:
                self.consumer.flush()  # the fold thread is done with what this slot held two exchanges ago
                self.borrowed[slot] = False
            self.dev_local[: self.B].copy_(local, non_blocking=True)
            self.dev_local[self.B].fill_(float(n_local))
            dist.all_gather_into_tensor(self.dev_all[slot], self.dev_local, group=self.group)
            if self.consumer is not None:
                self.host_all[slot].copy_(self.dev_all[slot], non_blocking=True)
                self.copied[slot].record()
                prev, self.pending = self.pending, slot
                if prev is not None:
                    self._hand_over(prev)  # the previous exchange's digests: their copy finished long ago
            self.sb_done += 1
            return
        elif self.world > 1:
            mine = torch.cat([local, torch.full((1, self.ndbl), float(n_local), dtype=torch.float64)])
            allbuf = torch.empty((self.world * rows, self.ndbl), dtype=torch.float64)
            dist.all_gather_into_tensor(allbuf, mine, group=self.group)
            allv = allbuf.view(self.world, rows, self.ndbl)
        else:
            allv = torch.cat([local, torch.full((1, self.ndbl), float(n_local), dtype=torch.float64)]).unsqueeze(0)
        if self.consumer is not None:
            self._fold_gathered(allv, borrowed=False)
        self.sb_done += 1

    def _fold_gathered(self, allv, borrowed: bool) -> None:
        seen_short = False
        for r in range(self.world):
            c = int(allv[r, self.B, 0].item())
            if seen_short and c:  # a short rank may only be followed by empty ranks (tail of the stream)
                raise RuntimeError("frames are not contiguous across ranks in this super-batch")
            seen_short |= c < self.B
            if c:
                self.consumer.consume_digests(allv[r].data_ptr(), c, borrowed=borrowed)  # folded asynchronously
                self.folded += c

    def _hand_over(self, slot: int) -> None:
        """rank 0, GPU path: give the gathered digests of `slot` to the consumer, read in place."""
        self.copied[slot].synchronize()
        self._fold_gathered(self.host_all[slot].view(self.world, self.B + 1, self.ndbl), borrowed=True)
        self.borrowed[slot] = True

    def exchange(self, final: bool = False) -> int:
        """Close the current super-batch.  Unless `final`, only the super-batch BEFORE it is gathered now
        (its kernels finished while this one was being queued), so the gather overlaps compute.  With
        `final` everything outstanding is drained, gathered and folded.  Returns frames folded so far."""
        closed = self.pushed - sum(self.sb_pushed)
        if closed > self.B:
            raise RuntimeError("more than frames_per_rank frames pushed in one super-batch")
        self.sb_pushed.append(closed)
        keep = 0 if final else 1
        while len(self.sb_pushed) - self.sb_done > keep:
            self._exchange_one(final)
        if final and self.consumer is not None:
            if self.on_gpu and self.world > 1 and self.pending is not None:
                self._hand_over(self.pending)
                self.pending = None
            self.consumer.flush()
            if self.on_gpu and self.world > 1:
                self.borrowed = [False, False]
        return self.folded

    def finish(self):
        """rank 0: the grain table; other ranks: None.  Call exchange(final=True) first."""
        if len(self.sb_pushed) != self.sb_done or self.pushed != sum(self.sb_pushed):
            self.exchange(final=True)
        return self.consumer.finish() if self.consumer is not None else None


def owner_of(frame_index: int, world: int, frames_per_rank: int) -> int:
    """Rank that owns a global frame index under the round-robin batch dealing."""
    return (frame_index // frames_per_rank) % world


def parity_check(device: int, frames_per_rank: int = 3, super_batches: int = 5, gram_order: int = 0):
    """Multi-rank parity, on an initialised process group: a small two-scene stream (segment cut, short tail
    super-batch) goes through ShardedDiff -- every rank pushes only the frames it owns -- and through ONE single-GPU
    handle on rank 0; the two grain tables must be identical.  Returns (frames, segments) on rank 0, (frames, None)
    elsewhere; raises on a mismatch.  Used by bench.py before its timed region at N > 1 and by tools/sharded_check.py."""
    from .synth import SynthSpec, make_pair_numpy
    rank, world = dist.get_rank(), dist.get_world_size()
    B, nsb = frames_per_rank, super_batches
    a = SynthSpec(320, 208, 10, textured=0.2, sigma0=1.0, sigma1=1.5, seed=5)
    b = SynthSpec(320, 208, 10, textured=0.2, sigma0=2.5, sigma1=0.5, ar_strength=0.6, seed=6)
    total = world * B * (nsb - 1) + (world - 1) * B + 1  # the tail super-batch: full ranks, then one frame
    frames = {k: make_pair_numpy(a if k < total // 2 else b, k) for k in range(total)
              if rank == 0 or owner_of(k, world, B) == rank}
    sd = ShardedDiff(30000, 1001, 10, 10, a.width, a.height, 1, 1, frames_per_rank=B, device=device, batch_frames=2)
    k = 0
    while k < total:
        base = k
        for j in range(world * B):
            g = base + j
            if g < total and owner_of(g, world, B) == rank:
                sd.push_local(*frames[g])
        k = min(total, base + world * B)
        sd.exchange(final=k >= total)
    table = sd.finish()
    sd.producer.close()
    nseg = None
    if rank == 0:
        ref = DiffGenerator(30000, 1001, 10, 10, a.width, a.height, device=device)
        for g in range(total):
            ref.diff_frame(*frames[g])
        want = ref.finish()
        ref.close()
        if table != want:
            raise AssertionError("sharded grain table differs from the single-handle table")
        if len(want) < 2:
            raise AssertionError("the parity stream is meant to cut a segment")
        nseg = len(want)
    dist.barrier()
    return total, nseg
