"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200_empanada.h
declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(g.LIB)
    hdr = open(os.path.join(ROOT, "include", "b200_empanada.h")).read()
    names = sorted(set(re.findall(r"\b(be_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.be_version() >= 100


def test_python_binding_covers_header():
    from empanada_napari_b200 import _lib, pdl  # noqa: F401  (pdl registers its signatures)
    hdr = open(os.path.join(ROOT, "include", "b200_empanada.h")).read()
    names = set(re.findall(r"\b(be_[a-z0-9_]+)\s*\(", hdr))
    unbound = sorted(n for n in names if n not in _lib._SIGNATURES and n not in ("be_scan_i32_to_i64",))
    assert not unbound, unbound


def test_no_cpu_fallback_message():
    import pytest
    import torch
    from empanada_napari_b200 import _lib
    from empanada_napari_b200.inference import Engine3d
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.B200EmpanadaError):
        Engine3d({"labels": [1], "class_names": {1: "m"}, "thing_list": [1], "padding_factor": 16,
                  "norms": {"mean": 0.5, "std": 0.1}, "model": None})


def test_engines_raise_without_cuda_and_without_the_library(monkeypatch):
    """No CPU fallback anywhere on the product path: without a CUDA device every engine raises,
    and a missing shared library is an error, not a detour."""
    import torch
    from empanada_napari_b200 import _lib, inference
    cfg = {"class_names": {1: "mito"}, "labels": [1], "thing_list": [1], "padding_factor": 16,
           "norms": {"mean": 0.5, "std": 0.1}, "model": "does-not-matter"}
    if not torch.cuda.is_available():
        for make in (lambda: inference.Engine3d(cfg), lambda: inference.Engine2d(cfg)):
            with pytest.raises(_lib.B200EmpanadaError, match="CUDA"):
                make()
    for make in (lambda: inference.Engine3d(cfg, use_gpu=False), lambda: inference.Engine2d(cfg, use_gpu=False)):
        with pytest.raises(_lib.B200EmpanadaError):
            make()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200_empanada.so")
    with pytest.raises(_lib.B200EmpanadaError, match="no CPU fallback"):
        _lib.call("be_version")
