"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): one process per GPU over NCCL; the
slice-sharded engine (ShardedEngine3d: median wavefront, boundary overlap exchange, table gather,
label-table broadcast) and the gather engine (DistributedEngine3d) must reproduce the single-GPU
engine bit for bit (tools/check_multigpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("engine,ks,fine", [("sharded", "3", "0"), ("sharded", "5", "0"), ("sharded", "1", "1"), ("gather", "3", "0")])
def test_distributed_engine_equals_single_gpu(engine, ks, fine):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_multigpu.py")]
    env = dict(os.environ, CHECK_ENGINE=engine, CHECK_KS=ks, CHECK_FINE=fine, CHECK_WITH_NETWORK="0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTIGPU_CHECK" in r.stdout and " PASS" in r.stdout
