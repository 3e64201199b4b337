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
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTIGPU_CHECK" in r.stdout and " PASS" in r.stdout


def test_multigpu_engine_front_end():
    """`MultiGPUEngine3d` as the widget constructs it (one process, reference keywords, worker
    processes started by the engine): same stacks, trackers and consensus as `Engine3d`."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_multigpu_front.py")],
                       capture_output=True, text=True, timeout=420, cwd=ROOT,
                       env=dict(os.environ, B200_EMPANADA_NCCL_TIMEOUT="120"))
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "MULTIGPU_FRONT PASS" in r.stdout


def test_multigpu_engine_needs_two_gpus():
    """multigpu.py:143-144: plain `Exception` below two GPUs."""
    import torch
    from empanada_napari_b200.multigpu import MultiGPUEngine3d
    if torch.cuda.device_count() > 1:
        pytest.skip("box has several GPUs")
    with pytest.raises(Exception, match="MultiGPU inference requires multiple GPUs"):
        MultiGPUEngine3d({"labels": [1], "thing_list": [1], "class_names": {1: "mito"}, "padding_factor": 16,
                          "norms": {"mean": 0.5, "std": 0.1}, "model": {}})
