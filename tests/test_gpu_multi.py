"""Sharded prover / commitment over NCCL on >= 2 GPUs (skipped on a single-GPU box; the host-side logic of the same
path is covered on the CPU by tests/test_sharded_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_prover_and_commit_on_two_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    for gather in ("16", "2"):  # default: leave the sharded regime early; 2: stay sharded down to 4-entry shards
        env = dict(os.environ, ZB_GATHER_LOG2=gather)
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert f"multi_gpu_check ok on {world} GPUs" in r.stdout
