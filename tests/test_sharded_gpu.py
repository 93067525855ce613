"""NCCL path of the frame-sharded driver (needs two GPUs; skipped otherwise): tools/sharded_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_two_rank_nccl_matches_single_handle():
    import torch
    if torch.cuda.device_count() < 2:
        # loud, not a silent skip: on a one-GPU box the NCCL path is NOT verified by this suite (bench.py --gpus N
        # repeats the same check before its timed region and prints "parity_checked" in its JSON line)
        pytest.xfail("NCCL sharded parity NOT verified here: this box has fewer than two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sharded_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
