"""Import shim for running the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY. Used by `oracle/make_golden.py` (fixture generation) and by the
pinning tests when /root/reference is present. Nothing in the product package imports this.

The reference imports third-party packages that are absent from this image (scikit-image,
zarr, dask, cztile, napari, ...). Only these are on the hot path: `skimage.measure.label` and `skimage.measure.regionprops` (call
sites empanada/inference/rle.py:22,75, filters.py:18,109) and, for the tracker morphology options,
`skimage.morphology.erosion` / `dilation` with their default footprint (filters.py:158,168 -
re-stated as scipy's grey erosion / dilation with the cross, borders reflected). They are re-stated here from their documented behaviour
(8-connected components of EQUAL-valued pixels, numbered in raster order of first pixel;
regionprops ascending by label, half-open bbox, row-major coords). scikit-image itself is a
lower-bounded, un-vendored dependency (setup.cfg:50) whose results are not pinned by any
reference test, so CC numbering is "parity unpinned" by the reference; this shim is what
defines the oracle for it.
"""
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("EMPANADA_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "empanada"))


# ----------------------------------------------------------------------------- skimage.measure
def sk_label(seg, background=0, connectivity=None):
    """Equal-value, full-connectivity connected components, raster-order numbering."""
    from scipy import ndimage as ndi

    seg = np.asarray(seg)
    structure = np.ones((3,) * seg.ndim, dtype=bool)
    out = np.zeros(seg.shape, dtype=np.int64)
    firsts = []  # (first flat index, value, local component id)
    for v in np.unique(seg):
        if v == background:
            continue
        lab, n = ndi.label(seg == v, structure=structure)
        if n == 0:
            continue
        flat = lab.ravel()
        idx = np.flatnonzero(flat)
        # first flat index of every local component
        order = np.argsort(flat[idx], kind="stable")
        sorted_labs = flat[idx][order]
        starts = np.flatnonzero(np.r_[True, sorted_labs[1:] != sorted_labs[:-1]])
        first_idx = idx[order][starts]
        for comp, fi in zip(sorted_labs[starts], first_idx):
            firsts.append((int(fi), v, int(comp)))
        out[lab > 0] = -1  # placeholder
    firsts.sort()
    # second pass: assign raster-order ids
    remap = {}
    for new_id, (fi, v, comp) in enumerate(firsts, start=1):
        remap[(v, comp)] = new_id
    for v in set(f[1] for f in firsts):
        lab, n = ndi.label(seg == v, structure=structure)
        lut = np.zeros(n + 1, dtype=np.int64)
        for comp in range(1, n + 1):
            lut[comp] = remap[(v, comp)]
        m = lab > 0
        out[m] = lut[lab[m]]
    return out


# ----------------------------------------------------------------------------- skimage.morphology
def sk_erosion(image, footprint=None, out=None):
    """skimage.morphology.erosion with its default footprint (the cross,
    ndi.generate_binary_structure(ndim, 1)): scipy's grey erosion, borders reflected."""
    from scipy import ndimage as ndi
    fp = ndi.generate_binary_structure(np.asarray(image).ndim, 1) if footprint is None else footprint
    return ndi.grey_erosion(image, footprint=fp)


def sk_dilation(image, footprint=None, out=None):
    from scipy import ndimage as ndi
    fp = ndi.generate_binary_structure(np.asarray(image).ndim, 1) if footprint is None else footprint
    return ndi.grey_dilation(image, footprint=fp)


class _RegionProp:
    __slots__ = ("label", "bbox", "coords", "area")

    def __init__(self, label, bbox, coords):
        self.label = label
        self.bbox = bbox
        self.coords = coords
        self.area = len(coords)


def sk_regionprops(lab, intensity_image=None, cache=True):
    lab = np.asarray(lab)
    props = []
    flat = lab.ravel()
    idx = np.flatnonzero(flat)
    if idx.size == 0:
        return props
    vals = flat[idx]
    order = np.argsort(vals, kind="stable")
    idx_sorted = idx[order]
    vals_sorted = vals[order]
    bounds = np.flatnonzero(np.r_[True, vals_sorted[1:] != vals_sorted[:-1], True])
    for s, e in zip(bounds[:-1], bounds[1:]):
        coords = np.stack(np.unravel_index(idx_sorted[s:e], lab.shape), axis=1)
        mins = coords.min(axis=0)
        maxs = coords.max(axis=0) + 1
        bbox = tuple(int(v) for v in mins) + tuple(int(v) for v in maxs)
        props.append(_RegionProp(int(vals_sorted[s]), bbox, coords))
    return props


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Make `import empanada...` / `import empanada_napari.inference` work from REF_ROOT."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True

    def _notimpl(*a, **k):
        raise NotImplementedError("stubbed third-party function (not on the hot path)")

    if "skimage" not in sys.modules:
        try:
            import skimage  # noqa: F401  (prefer the real one if it ever exists)
            import skimage.measure  # noqa: F401
        except Exception:
            sk = _stub("skimage")
            sk.measure = _stub("skimage.measure", label=sk_label, regionprops=sk_regionprops)
            sk.morphology = _stub("skimage.morphology", erosion=sk_erosion, dilation=sk_dilation,
                                  remove_small_objects=_notimpl, binary_erosion=_notimpl,
                                  binary_dilation=_notimpl)
            sk.draw = _stub("skimage.draw")
            sk.io = _stub("skimage.io")
            sk.segmentation = _stub("skimage.segmentation", watershed=_notimpl)
            sk.feature = _stub("skimage.feature", peak_local_max=_notimpl)
            sk.filters = _stub("skimage.filters")
            sk.transform = _stub("skimage.transform")

    class _ZarrArray:  # patterns.fill_volume does isinstance(volume, zarr.Array)
        pass

    for name, attrs in [
        ("zarr", dict(Array=_ZarrArray, open=_notimpl)),
        ("dask", {}),
        ("joblib", dict(Parallel=_notimpl, delayed=_notimpl)),
        ("cztile", {}),
        ("cztile.fixed_total_area_strategy_2d", dict(AlmostEqualBorderFixedTotalAreaStrategy2D=_notimpl)),
        ("cztile.tiling_strategy", dict(Rectangle=_notimpl, Region2D=_notimpl)),
        ("requests", {}),
        ("napari", {}),
        ("napari.qt", {}),
    ]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name, **attrs)
    if "dask.array" not in sys.modules:
        try:
            import dask.array  # noqa: F401
        except Exception:
            class _DaskArray:
                pass
            core = _stub("dask.array.core", Array=_DaskArray)
            da = _stub("dask.array", core=core, Array=_DaskArray)
            sys.modules["dask"].array = da
    if "napari.qt.threading" not in sys.modules:
        _stub("napari.qt.threading", thread_worker=lambda f=None, **k: (f if f is not None else (lambda g: g)))
        sys.modules["napari.qt"].threading = sys.modules["napari.qt.threading"]
        sys.modules["napari"].qt = sys.modules["napari.qt"]
    try:
        import urllib3.exceptions  # noqa: F401
    except Exception:
        _stub("urllib3")
        _stub("urllib3.exceptions", InsecureRequestWarning=Warning)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # bare namespace package so empanada_napari/__init__.py (Qt widgets) never runs
    if "empanada_napari" not in sys.modules:
        pkg = types.ModuleType("empanada_napari")
        pkg.__path__ = [os.path.join(REF_ROOT, "empanada_napari")]
        sys.modules["empanada_napari"] = pkg
    _installed = True
