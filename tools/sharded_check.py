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

from grav1synth_b200.sharded import parity_check


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    total, nseg = parity_check(torch.cuda.current_device())
    if rank == 0:
        print(f"sharded_check ok: world {world}, {total} frames, {nseg} segments")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
