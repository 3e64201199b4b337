"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200_empanada.h
declares (no compute calls here)."""
import ctypes
import os
import re

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
