"""-m gpu: zarr / dask volume I/O (/root/reference/empanada/zarr_utils.py:
97-184, patterns.py:204-213, data/volume_dataset.py:37-43, empanada_napari/inference.py:99-107,
474-489). zarr and dask are third-party packages that are not installed in this image, so the
tests stand in minimal objects with the same interface: a sliceable lazy array that records its
reads, and a `zarr` module whose arrays record every write (the product code only uses
zarr.open / create_array / chunk-aligned __setitem__)."""
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NORMS = {"mean": 0.57571, "std": 0.12765}


class LazyVolume:
    """zarr.Array-like: shape, dtype, chunks, slicing -> numpy; no __array__, no buffer protocol."""

    def __init__(self, data, chunks):
        self._data, self.shape, self.dtype, self.chunks = data, data.shape, data.dtype, chunks
        self.reads = []

    def __getitem__(self, key):
        self.reads.append(key)
        return self._data[key].copy()


class DaskLike(LazyVolume):
    """dask-like: slicing returns a lazy object with .compute(); chunks are tuples of tuples."""

    def __init__(self, data, chunks):
        super().__init__(data, tuple(tuple([c] * (-(-s // c))) for s, c in zip(data.shape, chunks)))

    def __getitem__(self, key):
        self.reads.append(key)
        block = self._data[key].copy()
        return types.SimpleNamespace(compute=lambda: block, shape=block.shape)


class FakeZarrArray:
    def __init__(self, shape, dtype, chunks):
        self.shape, self.dtype, self.chunks = tuple(shape), np.dtype(dtype), tuple(chunks)
        self.data = np.full(shape, 77, dtype=dtype)      # poison: every voxel must be written
        self.writes = []

    def __setitem__(self, key, value):
        assert np.asarray(value).dtype == self.dtype
        self.writes.append(tuple((k.start, k.stop) for k in key))
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]


class FakeZarrGroup:
    def __init__(self):
        self.arrays = {}

    def create_array(self, name, shape, dtype, chunks, overwrite=False):
        assert overwrite
        self.arrays[name] = FakeZarrArray(shape, dtype, chunks)
        return self.arrays[name]


@pytest.fixture
def fake_zarr(monkeypatch):
    stores = {}
    mod = types.ModuleType("zarr")
    mod.Array = FakeZarrArray

    def _open(url, mode=None):
        if mode == "w" or url not in stores:
            stores[url] = FakeZarrGroup()
        return stores[url]
    mod.open = _open
    monkeypatch.setitem(sys.modules, "zarr", mod)
    return stores


def _setup(shape=(40, 70, 52), seed=41):
    import torch
    import empanada_napari_b200.synthetic as syn
    from empanada_napari_b200.model import SyntheticHeadsModel
    vol, lab, _ = syn.make_volume(shape, seed=seed, scale=1.0)
    dev = torch.device("cuda:0")
    heads = {}
    for a in range(3):
        hs = [syn.analytic_heads(np.take(lab, i, axis=a), pad_to=16) for i in range(shape[a])]
        heads[a] = tuple(torch.from_numpy(np.stack([h[k] if k else h[0][0] for h in hs])).to(dev) for k in range(3))
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16, "norms": NORMS,
           "model": SyntheticHeadsModel(lambda a, s0, s1: tuple(t[s0:s1] for t in heads[a]))}
    return vol, cfg


def _assert_chunk_aligned_once(arr):
    D, H, W = arr.shape
    dc, hc, wc = arr.chunks
    want = {((z, min(D, z + dc)), (y, min(H, y + hc)), (x, min(W, x + wc)))
            for z in range(0, D, dc) for y in range(0, H, hc) for x in range(0, W, wc)}
    assert len(arr.writes) == len(want) and set(arr.writes) == want      # every chunk, exactly once


@pytest.mark.parametrize("lazy_cls", [LazyVolume, DaskLike])
def test_streamed_lazy_input_equals_numpy(lazy_cls):
    from empanada_napari_b200.inference import Engine3d
    from conftest import assert_instances_equal
    vol, cfg = _setup()
    kw = dict(median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=30, min_extent=3, save_panoptic=True,
              batch_size=8)
    ref = Engine3d(cfg, **kw)
    eng = Engine3d(cfg, **kw)
    lazy = lazy_cls(vol, (16, 32, 32))
    for axis_name in ("xy", "yz"):
        stack, trs = eng.infer_on_axis(lazy, axis_name)
        rstack, rtrs = ref.infer_on_axis(vol, axis_name)
        assert np.array_equal(stack, rstack)
        assert_instances_equal(trs[0].instances, rtrs[0].instances)
    # read once (cached for the second plane), in whole chunk rows of the store
    starts = [k.start for k in lazy.reads]
    assert starts == sorted(set(starts)) and all(s % 16 == 0 for s in starts)
    assert sum(k.stop - k.start for k in lazy.reads) == vol.shape[0]


def test_zarr_store_outputs(fake_zarr):
    from empanada_napari_b200.inference import Engine3d, stack_postprocessing, tracker_consensus
    vol, cfg = _setup()
    kw = dict(median_kernel_size=3, nms_kernel=3, confidence_thr=0.5, min_size=30, min_extent=3, save_panoptic=True,
              batch_size=8)
    chunks = (16, 32, 24)
    eng = Engine3d(cfg, store_url="mem://seg", chunk_size=chunks, **kw)
    ref = Engine3d(cfg, **kw)
    assert eng.zarr_store is fake_zarr["mem://seg"]
    trackers, rtrackers = {}, {}
    for axis_name in ("xy", "xz", "yz"):
        stack, trackers[axis_name] = eng.infer_on_axis(vol, axis_name)
        rstack, rtrackers[axis_name] = ref.infer_on_axis(vol, axis_name)
        assert stack is fake_zarr["mem://seg"].arrays[f"panoptic_{axis_name}"]
        assert stack.dtype == np.int32 and stack.chunks == chunks and np.array_equal(stack.data, rstack)
        _assert_chunk_aligned_once(stack)
    vote = dict(pixel_vote_thr=2, min_size=30, min_extent=3, dtype=np.uint32)
    (v, name, inst), = list(tracker_consensus(trackers, "mem://seg", cfg, chunk_size=chunks, **vote))
    (rv, _, rinst), = list(tracker_consensus(rtrackers, None, cfg, **vote))
    assert v is fake_zarr["mem://seg"].arrays["mito"] and v.dtype == np.uint32 and len(inst) == len(rinst) > 0
    assert np.array_equal(v.data, rv)
    _assert_chunk_aligned_once(v)
    (v, name, inst), = list(stack_postprocessing({"xy": trackers["xy"]}, "mem://seg", cfg, min_size=30, min_extent=3,
                                                  dtype=np.uint32, chunk_size=(7, 70, 52)))
    (rv, _, rinst), = list(stack_postprocessing({"xy": rtrackers["xy"]}, None, cfg, min_size=30, min_extent=3, dtype=np.uint32))
    assert np.array_equal(v.data, rv) and v.chunks == (7, 70, 52)
    _assert_chunk_aligned_once(v)


def test_store_url_without_zarr_raises(monkeypatch):
    from empanada_napari_b200 import _lib
    from empanada_napari_b200.inference import Engine3d
    monkeypatch.setitem(sys.modules, "zarr", None)       # import zarr -> ImportError
    _, cfg = _setup((8, 32, 32))
    with pytest.raises(_lib.B200EmpanadaError, match="zarr"):
        Engine3d(cfg, store_url="/tmp/x.zarr")
