"""World-size-2 test of the multi-process plumbing on CPU (gloo): contiguous slice ranges and the
gather of per-rank head blocks to rank 0 reassemble the plane in slice order."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from empanada_napari_b200.multigpu import gather_slices_to_root, slice_ranges
    ranges = slice_ranges(n, world)
    lo, hi = ranges[rank]
    nmax = max(b - a for a, b in ranges)
    full = torch.arange(n * 6, dtype=torch.float32).reshape(n, 2, 3)
    local = torch.zeros((nmax, 2, 3))
    local[: hi - lo] = full[lo:hi]
    blocks = gather_slices_to_root(local, ranges, rank, world)
    if rank == 0:
        out = torch.cat(blocks)
        np.save(os.path.join(tmpdir, "out.npy"), out.numpy())
    else:
        assert blocks is None
    dist.barrier()
    dist.destroy_process_group()


def test_slice_ranges_cover_and_balance():
    from empanada_napari_b200.multigpu import slice_ranges
    for n in (1, 7, 8, 1024, 1030):
        for w in (1, 2, 3, 8):
            r = slice_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gather_two_ranks_gloo(tmp_path):
    n, world = 11, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npy")
    assert np.array_equal(out, np.arange(n * 6, dtype=np.float32).reshape(n, 2, 3))
