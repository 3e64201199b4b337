"""Host-side tracker containers (no GPU): the API-compatible InstanceTracker and the
self-completing PendingTracker used to overlap the matcher replay with the next plane."""
import json

import numpy as np


def test_pending_tracker_resolves_on_first_access(tmp_path):
    from empanada_napari_b200.tracking import InstanceTracker, PendingTracker
    calls = []
    tr = PendingTracker(1, 1000, (4, 5, 6), "xz")

    def resolve():
        calls.append(1)
        tr.instances = {7: {"box": (0, 0, 0, 1, 2, 3), "starts": np.array([3, 40]), "runs": np.array([2, 1])}}
        tr._b200_sizes = {7: 3}
        tr._b200_dense = "dense"
        tr.finish()

    tr._resolver = resolve
    assert tr.axis == "xz" and tr.class_id == 1 and not calls       # plain attributes do not resolve
    assert getattr(tr, "_b200_sizes", None) == {7: 3} and calls == [1]
    assert list(tr.instances.keys()) == [7] and calls == [1]         # resolved once
    assert tr.finished
    path = tmp_path / "t.json"
    tr.write_to_json(str(path))
    d = json.load(open(path))
    assert d["instances"]["7"]["rle"] == "3 2 40 1" and d["axis"] == "xz" and "_resolver" not in d
    back = InstanceTracker()
    back.load_from_json(str(path))
    assert np.array_equal(back.instances["7"]["starts"], [3, 40])


def test_pending_tracker_without_resolver_behaves_like_empty_tracker():
    from empanada_napari_b200.tracking import PendingTracker
    tr = PendingTracker(2, 1000, (1, 1, 1), "xy")
    assert tr.instances == {}
    assert getattr(tr, "_b200_dense", None) is None


def test_async_reraises_on_caller_thread():
    import pytest
    from empanada_napari_b200.inference import _Async
    assert _Async(lambda a, b: a + b, 2, 3).result() == 5

    def boom():
        raise ValueError("x")
    with pytest.raises(ValueError):
        _Async(boom).result()


def test_owned_ranges_cover_plane_and_shift_by_median_latency():
    from empanada_napari_b200.multigpu import owned_ranges
    for n, world, mid in ((20, 3, 1), (1024, 8, 2), (61, 4, 0), (50, 2, 2)):
        F, E = owned_ranges(n, world, mid)
        assert F[0][0] == 0 and F[-1][1] == n and E[0][0] == 0 and E[-1][1] == n
        for r in range(world - 1):
            assert F[r][1] == F[r + 1][0] and E[r][1] == E[r + 1][0]      # contiguous, disjoint
            assert E[r][1] == F[r][1] - mid                               # pushing slice t emits t - mid
        for r in range(1, world):
            assert E[r][0] == F[r][0] - mid


def test_merge_shard_tables_absolute_slice_keys():
    from empanada_napari_b200.multigpu import merge_shard_tables
    k = lambda s, q, c: np.uint64((s << 40) | (q << 20) | c)
    p0 = (np.array([2, 1], np.int32), np.zeros((2, 2, 5), np.int32), np.array([k(1, 1, 1)], np.uint64), np.array([7], np.int32))
    p1 = (np.array([3], np.int32), np.ones((1, 3, 5), np.int32), np.array([k(0, 1, 2)], np.uint64), np.array([9], np.int32))
    n_cc, table, keys, vals = merge_shard_tables([p0, p1], [(0, 2), (2, 3)])
    assert n_cc.tolist() == [2, 1, 3] and table.shape == (3, 3, 5) and table[2].min() == 1 and table[:2].max() == 0
    assert [int(x) >> 40 for x in keys] == [1, 2] and vals.tolist() == [7, 9]
    assert int(keys[1]) & ((1 << 40) - 1) == (1 << 20) | 2


def test_auto_slice_batch():
    from empanada_napari_b200.inference import auto_slice_batch
    assert auto_slice_batch(1024, 1024, 148) == 37          # 37 * 32 tiles = 8 waves of 148
    assert (auto_slice_batch(1024, 1024, 148) * 32) % 148 == 0
    assert auto_slice_batch(2048, 2048, 148) == 10          # capped by the 40 MPixel budget
    assert auto_slice_batch(64, 64, 148) == 32              # tiny slices: plain default
    assert auto_slice_batch(1024, 1024, 148, requested=16) == 16
    assert auto_slice_batch(2048, 2048, 148, requested=64) == 10
    assert auto_slice_batch(8192, 8192, 148) == 1
