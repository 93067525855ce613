"""Frame-sharded multi-GPU driver (one process per GPU, torch.distributed for the plumbing).

The path shards by frame: every per-pixel quantity is a function of one frame
pair, and the only cross-frame logic is the small sequential model merge.  So each rank runs a
PRODUCER handle on its own frames: kernels, plus the per-frame half of the host model, whose result
is a fixed-size ~27 KB digest per frame written straight into a pinned tensor.  Once per super-batch
the digests are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and rank 0 folds them
in global frame order into one CONSUMER handle.  A super-batch is world*B consecutive frames; rank r
owns frames [base + r*B, base + (r+1)*B), so the rank-major gather result is already in frame order.
There is no data-path collective besides that gather.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import abi
from .diff import DiffGenerator, digest_bytes


class ShardedDiff:
    def __init__(self, fps_num: int, fps_den: int, source_bit_depth: int, denoised_bit_depth: int, width: int,
                 height: int, ss_x: int = 1, ss_y: int = 1, frames_per_rank: int = 8, device: Optional[int] = None,
                 producer_factory: Optional[Callable[[], object]] = None, group=None, batch_frames: int = 0):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.B = frames_per_rank
        self.ndbl = digest_bytes() // 8
        args = (fps_num, fps_den, source_bit_depth, denoised_bit_depth, width, height, ss_x, ss_y)
        if producer_factory is not None:          # tests inject a CPU digest producer
            self.producer = producer_factory()
        else:
            self.producer = DiffGenerator(*args, device=device or 0, batch_frames=batch_frames,
                                          mode=abi.MODE_PRODUCER)
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self.on_gpu = backend == "nccl"
        # the producer writes digests straight into this (pinned when a GPU is involved) tensor
        self.sink = torch.zeros((self.B, self.ndbl), dtype=torch.float64)
        if self.on_gpu:
            self.sink = self.sink.pin_memory()
            dev = torch.device("cuda", torch.cuda.current_device())
            self.dev_local = torch.empty((self.B + 1, self.ndbl), dtype=torch.float64, device=dev)
            self.dev_all = torch.empty((self.world * (self.B + 1), self.ndbl), dtype=torch.float64, device=dev)
            self.host_all = torch.empty((self.world * (self.B + 1), self.ndbl), dtype=torch.float64).pin_memory()
        self.producer.set_digest_sink(self.sink.data_ptr(), self.B)
        self.consumer = DiffGenerator(*args, mode=abi.MODE_CONSUMER) if self.rank == 0 else None

    # -- per super-batch --------------------------------------------------------------
    def push_local(self, source, denoised, device_resident: bool = False) -> None:
        """Queue one of this rank's frames of the current super-batch (at most B per super-batch)."""
        if device_resident:
            self.producer.diff_frame_device(*source, *denoised)
        else:
            self.producer.diff_frame(source, denoised)

    def exchange(self) -> int:
        """Finish the super-batch: drain the local kernels, gather all ranks' digests, fold them
        into the model on rank 0 in global frame order.  Returns the frames folded (rank 0)."""
        self.producer.flush()
        n_local = self.producer.digest_count
        if n_local > self.B:
            raise RuntimeError("more than frames_per_rank frames pushed in one super-batch")
        folded = 0
        if self.world > 1 and self.on_gpu:
            # row B of every rank's block carries its frame count
            self.dev_local[: self.B].copy_(self.sink, non_blocking=True)
            self.dev_local[self.B].fill_(float(n_local))
            dist.all_gather_into_tensor(self.dev_all, self.dev_local, group=self.group)
            if self.consumer is not None:
                self.host_all.copy_(self.dev_all, non_blocking=False)
                allv = self.host_all.view(self.world, self.B + 1, self.ndbl)
        elif self.world > 1:
            local = torch.cat([self.sink, torch.full((1, self.ndbl), float(n_local), dtype=torch.float64)])
            allbuf = torch.empty((self.world * (self.B + 1), self.ndbl), dtype=torch.float64)
            dist.all_gather_into_tensor(allbuf, local, group=self.group)
            allv = allbuf.view(self.world, self.B + 1, self.ndbl)
        else:
            allv = torch.cat([self.sink, torch.full((1, self.ndbl), float(n_local), dtype=torch.float64)]).unsqueeze(0)
        if self.consumer is not None:
            seen_short = False
            for r in range(self.world):
                c = int(allv[r, self.B, 0].item())
                if seen_short and c:  # a short rank may only be followed by empty ranks (tail of the stream)
                    raise RuntimeError("frames are not contiguous across ranks in this super-batch")
                seen_short |= c < self.B
                if c:
                    self.consumer.consume_digests(allv[r].data_ptr(), c)
                    folded += c
        self.producer.set_digest_sink(self.sink.data_ptr(), self.B)  # reset the count for the next super-batch
        return folded

    def finish(self):
        """rank 0: the grain table; other ranks: None."""
        return self.consumer.finish() if self.consumer is not None else None


def owner_of(frame_index: int, world: int, frames_per_rank: int) -> int:
    """Rank that owns a global frame index under the round-robin batch dealing."""
    return (frame_index // frames_per_rank) % world
