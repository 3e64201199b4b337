"""ctypes binding of the C-ABI shared library (include/b200_empanada.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t,
                    c_ulonglong, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200_empanada.so")

_lib = None
RUN_CHUNK = 4096  # elements per CTA of be_runs_count / be_runs_write (csrc/cc_kernels.cu)

P = c_void_p  # device or host pointer passed as an integer address
I, F, D, LL, ULL, SZ = c_int, c_float, c_double, c_longlong, c_ulonglong, c_size_t

_SIGNATURES = {
    "be_version": ([], c_int),
    "be_last_error": ([], c_char_p),
    # post_kernels.cu
    "be_median_push": ([P, I, I, I, I, P, I, I, F, I, P, P, P], I),
    "be_median_flush": ([P, I, I, I, I, I, F, P, P, P], I),
    "be_centers": ([P, I, I, I, F, I, P, I, P, P, P], I),
    "be_group_pixels": ([P, P, I, P, I, I, I, F, P, P], I),
    "be_merge_pan": ([P, P, I, I, I, I, I, I, I, I, I, I, P, P, P], I),
    "be_rank_ids": ([P, I, I, I, I, P], I),
    # run_kernels.cu
    "be_group_flags": ([P, P, P, I, P, I, I, I, I, I, P, P, P], I),
    "be_slice_area": ([P, I, I, I, P, P], I),
    "be_rowruns_count": ([P, P, P, I, I, I, I, I, I, I, I, I, I, P, P, P, P, P, P], I),
    "be_rowruns_write": ([P, P, P, I, I, I, I, I, I, I, I, I, I, P, P, P, P, P, P, P, P], I),
    "be_runs_cc": ([P, P, P, P, P, I, I, I, P, P, P, P], I),
    "be_runs_stats": ([P, P, P, P, I, I, I, P, P], I),
    "be_runs_overlap": ([P, P, P, P, P, I, I, I, I, P, P, ULL, P, P], I),
    "be_runs_paint": ([P, P, P, P, P, P, I, I, I, I, I, I, P, LL, LL, LL, P], I),
    # cc_kernels.cu
    "be_cc_label": ([P, I, I, I, I, I, P, P, P, P, I, P, P], I),
    "be_hash_clear": ([P, P, ULL, P], I),
    "be_pair_overlap": ([P, I, I, I, I, P, P, ULL, P, P], I),
    "be_hash_compact": ([P, P, ULL, P, P, I, P, P], I),
    "be_relabel": ([P, I, I, I, I, P, I, P, LL, LL, LL, P], I),
    "be_runs_count": ([P, LL, LL, P, P], I),
    "be_scan_i32_to_i64": ([P, P, LL, P, SZ, POINTER(SZ), P], I),
    "be_runs_write": ([P, LL, LL, P, P, P, P, LL, P], I),
    "be_sort_runs": ([P, P, P, P, I, P, SZ, POINTER(SZ), P], I),
    "be_up4": ([P, I, I, I, P, P], I),
    "be_resize_linear_u8": ([P, LL, LL, LL, I, I, I, I, I, P, P], I),
    # consensus_kernels.cu
    "be_plane_pairs": ([P, P, P, P, P, P, I, I, I, LL, I, P, P, ULL, P, P], I),
    "be_vote_stats": ([P, P, P, P, P, P, I, I, I, LL, I, P, P, I, P, P, P, ULL, P, P], I),
    "be_vote_paint": ([P, P, P, P, P, P, I, I, I, LL, I, P, P, I, P, P, P, P, I, P, P], I),
    "be_label_hist": ([P, LL, I, I, P, P], I),
    "be_lut_inplace": ([P, LL, P, I, P], I),
    # consensus_runs.cu
    "be_triple_count": ([P, P, P, P, P, P, I, I, I, LL, I, P, P], I),
    "be_triple_write": ([P, P, P, P, P, P, I, I, I, LL, I, P, P, P, P, P, P], I),
    "be_triple_pairs": ([P, P, P, LL, P, P, ULL, P, P], I),
    "be_triple_stats": ([P, P, P, LL, P, P, I, P, P, P, ULL, P, P], I),
    "be_triple_rec_count": ([P, P, P, LL, P, P, I, P, P, P, P], I),
    "be_triple_rec_write": ([P, P, P, LL, I, LL, P, P, I, P, P, P, P, P, P, P], I),
    "be_sort_records": ([P, P, P, P, LL, P, SZ, POINTER(SZ), P], I),
    "be_join_flags": ([P, P, LL, P, P, P, SZ, POINTER(SZ), P], I),
    "be_join_write": ([P, P, P, P, LL, P, P, P, P], I),
    # morph_kernels.cu
    "be_morph3d": ([P, P, I, I, I, I, P], I),
    "be_range_keep": ([P, LL, I, I, P], I),
    "be_runs3d_cc": ([P, P, P, I, I, I, I, P, P, P, P], I),
    "be_fill_holes": ([P, P, I, I, I, P, P, P, P], I),
    # cluster_graph.cpp (host)
    "be_components_clusters": ([I, P, P, P, P, P, P, P, LL, D, D, D, I, P, P], I),
    "be_components_clusters_fetch": ([P, P], I),
    # match_replay.cpp (host)
    "be_match_replay": ([I, P, P, I, P, P, LL, I, I, D, D, I, P, I, P, P, P, I, P], I),
    "be_match_replay_stats": ([P], I),
}


class B200EmpanadaError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200EmpanadaError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = L
    return _lib


def declare(name, argtypes, restype=c_int):
    """Register an additional entry point (used by modules that own their own signatures)."""
    _SIGNATURES[name] = (argtypes, restype)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.argtypes = argtypes
        fn.restype = restype


# Optional per-entry-point device timing (bench.py's live roofline figures): when TIMING is a
# dict, every call is bracketed by CUDA events on the current stream; `timing_summary()`
# synchronises and returns {entry point: (calls, total ms)}. Never enabled inside a timed region.
TIMING = None


def timing_start():
    global TIMING
    TIMING = {}


def timing_summary():
    global TIMING
    import torch
    torch.cuda.synchronize()
    out = {k: (len(v), float(sum(a.elapsed_time(b) for a, b in v))) for k, v in (TIMING or {}).items()}
    TIMING = None
    return out


def call(name, *args):
    """Call an int-returning entry point; raise with be_last_error() on failure."""
    L = lib()
    if TIMING is not None and name != "be_match_replay":
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, name)(*args)
        e1.record()
        TIMING.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(L, name)(*args)
    if rc != 0:
        raise B200EmpanadaError(f"{name} failed ({rc}): {L.be_last_error().decode(errors='replace')}")
    return rc


def ptr(t):
    """Address of a torch tensor / numpy array (or None)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
