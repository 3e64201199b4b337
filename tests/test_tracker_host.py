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
