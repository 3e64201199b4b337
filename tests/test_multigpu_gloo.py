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


def _export_child(q_out, q_in):
    import numpy as np
    from empanada_napari_b200 import multigpu as m
    exports = []
    arrays = (np.arange(400000, dtype=np.int32), np.arange(400000, dtype=np.int64) * 7, np.ones(400000, np.int64))
    q_out.put(m._export_arrays(arrays, exports))
    q_in.get()                   # the parent has read the segment
    m._release_exports(exports)
    q_out.put("released")


def test_large_tables_cross_processes_through_shared_memory():
    """The per-rank range tables of the sharded consensus travel as a shared-memory descriptor:
    written by one process, read (copied out) by another, unlinked by the writer."""
    import multiprocessing as mp
    import numpy as np
    from empanada_napari_b200 import multigpu as m
    ctx = mp.get_context("spawn")
    q_out, q_in = ctx.Queue(), ctx.Queue()
    p = ctx.Process(target=_export_child, args=(q_out, q_in))
    p.start()
    desc = q_out.get(timeout=120)
    assert desc[0] == "__b200_shm__"
    got = m._import_arrays(desc)
    assert np.array_equal(got[0], np.arange(400000, dtype=np.int32)) and got[0].dtype == np.int32
    assert np.array_equal(got[1], np.arange(400000, dtype=np.int64) * 7) and np.all(got[2] == 1)
    q_in.put("read")
    assert q_out.get(timeout=60) == "released"
    p.join(timeout=60)
    assert p.exitcode == 0
    # tuples of plain arrays are passed through untouched
    t = (np.arange(3), np.arange(3), np.arange(3))
    assert m._import_arrays(t) is t
