"""Frame-sharded multi-GPU driver (one process per GPU, torch.distributed for the plumbing).

The path shards by frame: every per-pixel quantity is a function of one frame
pair, and the only cross-frame logic is the small sequential model update.  So each rank runs a
PRODUCER handle (kernels only) on its own frames, the fixed-size per-frame integer records are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) once per super-batch, and rank 0
feeds them in global frame order to one CONSUMER handle.  A super-batch is world*B consecutive
frames; rank r owns frames [base + r*B, base + (r+1)*B), so the rank-major gather result is
already in frame order.  There is no data-path collective besides that record gather.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import abi
from .diff import DiffGenerator, RecordLayout


class ShardedDiff:
    def __init__(self, fps_num: int, fps_den: int, source_bit_depth: int, denoised_bit_depth: int, width: int,
                 height: int, ss_x: int = 1, ss_y: int = 1, frames_per_rank: int = 8, device: Optional[int] = None,
                 producer_factory: Optional[Callable[[], object]] = None, group=None):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.B = frames_per_rank
        self.nb = ((width + 31) // 32) * ((height + 31) // 32)
        self.layout = RecordLayout(self.nb)
        self._records: List[np.ndarray] = []
        args = (fps_num, fps_den, source_bit_depth, denoised_bit_depth, width, height, ss_x, ss_y)
        if producer_factory is not None:          # tests inject a CPU record producer
            self.producer = producer_factory()
        else:
            self.producer = DiffGenerator(*args, device=device or 0, batch_frames=frames_per_rank,
                                          mode=abi.MODE_PRODUCER)
        self.producer.set_record_tap(lambda i, r: self._records.append(r))
        self.consumer = DiffGenerator(*args, mode=abi.MODE_CONSUMER) if self.rank == 0 else None
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self.comm_device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    # -- per super-batch --------------------------------------------------------------
    def push_local(self, source, denoised, device_resident: bool = False) -> None:
        """Queue one of this rank's frames of the current super-batch (at most B per super-batch)."""
        if device_resident:
            self.producer.diff_frame_device(*source, *denoised)
        else:
            self.producer.diff_frame(source, denoised)

    def exchange(self) -> int:
        """Finish the super-batch: drain the local kernels, gather all ranks' records, fold them
        into the model on rank 0 in global frame order.  Returns the frames folded (rank 0)."""
        self.producer.flush()
        n_local = len(self._records)
        if n_local > self.B:
            raise RuntimeError("more than frames_per_rank frames pushed in one super-batch")
        buf = torch.zeros((self.B, self.layout.bytes), dtype=torch.uint8)
        if n_local:
            buf[:n_local] = torch.from_numpy(np.stack(self._records))
        self._records.clear()
        count = torch.tensor([n_local], dtype=torch.int64)
        if self.world > 1:
            buf = buf.to(self.comm_device)
            count = count.to(self.comm_device)
            allbuf = torch.empty((self.world * self.B, self.layout.bytes), dtype=torch.uint8, device=self.comm_device)
            allcnt = torch.empty((self.world,), dtype=torch.int64, device=self.comm_device)
            dist.all_gather_into_tensor(allbuf, buf, group=self.group)
            dist.all_gather_into_tensor(allcnt, count, group=self.group)
            allbuf, allcnt = allbuf.cpu().view(self.world, self.B, self.layout.bytes), allcnt.cpu()
        else:
            allbuf, allcnt = buf.unsqueeze(0), count.unsqueeze(0)
        folded = 0
        if self.consumer is not None:
            counts = [int(c) for c in allcnt.view(-1)]
            # a short rank may only be followed by empty ranks (tail of the stream)
            seen_short = False
            for r, c in enumerate(counts):
                if seen_short and c:
                    raise RuntimeError("frames are not contiguous across ranks in this super-batch")
                seen_short |= c < self.B
                if c:
                    self.consumer.consume_records(allbuf[r].numpy()[:c])
                    folded += c
        return folded

    def finish(self):
        """rank 0: the grain table; other ranks: None."""
        return self.consumer.finish() if self.consumer is not None else None


def owner_of(frame_index: int, world: int, frames_per_rank: int) -> int:
    """Rank that owns a global frame index under the round-robin batch dealing."""
    return (frame_index // frames_per_rank) % world
