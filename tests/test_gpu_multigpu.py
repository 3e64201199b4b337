"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): one process per GPU over NCCL;
DistributedEngine3d must reproduce the single-GPU engine bit for bit (tools/check_multigpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_distributed_engine_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_multigpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTIGPU_CHECK PASS" in r.stdout
