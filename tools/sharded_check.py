#!/usr/bin/env python
"""torchrun check of the NCCL frame-sharded path: the table rank 0 builds from the gathered digests must equal
the table of one handle fed the same frames in global order.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_check.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

from grav1synth_b200 import diff as D
from grav1synth_b200.sharded import ShardedDiff, owner_of
from grav1synth_b200.synth import SynthSpec, make_pair_numpy


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    B, nsb = 3, 5  # frames per rank per super-batch, super-batches (the last one short)
    a = SynthSpec(320, 208, 10, textured=0.2, sigma0=1.0, sigma1=1.5, seed=5)
    b = SynthSpec(320, 208, 10, textured=0.2, sigma0=2.5, sigma1=0.5, ar_strength=0.6, seed=6)
    total = world * B * (nsb - 1) + (world - 1) * B + 1  # the tail super-batch: full ranks, then one frame
    frames = [make_pair_numpy(a if k < total // 2 else b, k) for k in range(total)]
    sd = ShardedDiff(30000, 1001, 10, 10, a.width, a.height, 1, 1, frames_per_rank=B,
                     device=torch.cuda.current_device(), batch_frames=2)
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
    if rank == 0:
        ref = D.DiffGenerator(30000, 1001, 10, 10, a.width, a.height)
        for s, d in frames:
            ref.diff_frame(s, d)
        want = ref.finish()
        assert table == want, "sharded table differs from the single-handle table"
        assert len(want) >= 2, "the check wants a segment cut"
        print(f"sharded_check ok: world {world}, {total} frames, {len(want)} segments")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
