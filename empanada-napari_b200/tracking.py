"""Host side of cross-slice instance stitching (piece 5): the sequential matcher/tracker
decisions are replayed by native code (csrc/match_replay.cpp) on the sparse tables the GPU
produced; this module owns the table plumbing and the API-compatible InstanceTracker.

Mirrors empanada/inference/tracker.py:40-159 (InstanceTracker) and filters.py:22-56.
"""
import json
from copy import deepcopy

import numpy as np

from . import _lib

AXIS_NUM = {"xy": 0, "xz": 1, "yz": 2}


def rle_to_string(starts, runs):
    """empanada/array_utils.py:258-271."""
    return " ".join(f"{i} {r}" for i, r in zip(starts, runs))


def string_to_rle(encoding):
    """empanada/array_utils.py:273-287."""
    enc = np.array([int(i) for i in encoding.split(" ")])
    return enc[::2], enc[1::2]


class InstanceTracker:
    """Same public attributes as the reference tracker: `instances` is
    {label: {'box': 6-tuple, 'starts': int64[], 'runs': int64[]}} in first-arrival order."""

    def __init__(self, class_id=None, label_divisor=None, shape3d=None, axis="xy"):
        assert axis in ["xy", "xz", "yz"]
        self.class_id = class_id
        self.label_divisor = label_divisor
        self.shape3d = shape3d
        self.axis = axis
        self.finished = False
        self.instances = {}
        self.axis_nums = {"xy": 0, "xz": 1, "yz": 2}

    def reset(self):
        self.instances = {}

    def finish(self):
        self.finished = True

    def write_to_json(self, savepath):
        if not self.finished:
            self.finish()
        save = {k: v for k, v in self.__dict__.items() if not k.startswith("_") and k != "instances"}
        save = deepcopy(save)
        save["instances"] = self.instances
        inst = {}
        for k, v in save["instances"].items():
            inst[str(k)] = {"box": [int(b) for b in v["box"]],
                            "rle": rle_to_string(v["starts"], v["runs"])}
        save["instances"] = inst
        save["shape3d"] = [int(s) for s in self.shape3d]
        with open(savepath, mode="w") as handle:
            json.dump(save, handle, indent=6)

    def load_from_json(self, fpath):
        with open(fpath, mode="r") as handle:
            d = json.load(handle)
        for k in d["instances"].keys():
            starts, runs = string_to_rle(d["instances"][k]["rle"])
            d["instances"][k]["starts"] = starts
            d["instances"][k]["runs"] = runs
        self.__dict__.update(d)


class PendingTracker(InstanceTracker):
    """Tracker whose tables are completed on first access of `instances` (or of the attached
    device volume): `Engine3d` overlaps the host matcher replay of one plane with the next plane's
    forward pass and resolves the tracker when somebody looks at it."""

    _LAZY = ("instances", "_b200_dense", "_b200_sizes")

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__.pop("instances", None)
        self._resolver = None

    def __getattr__(self, name):  # only called when normal lookup fails
        if name in PendingTracker._LAZY:
            resolver = self.__dict__.get("_resolver")
            if resolver is not None:
                self.__dict__["_resolver"] = None
                resolver()
                return self.__dict__[name]
            if name == "instances":
                self.__dict__["instances"] = {}
                return self.__dict__["instances"]
        raise AttributeError(name)


def remove_small_objects(tracker, min_size=64):
    """filters.py:22-36."""
    for iid in list(tracker.instances.keys()):
        if tracker.instances[iid]["runs"].sum() < min_size:
            del tracker.instances[iid]


def remove_pancakes(tracker, min_span=4):
    """filters.py:38-56."""
    for iid in list(tracker.instances.keys()):
        b = tracker.instances[iid]["box"]
        if any(s < min_span for s in (b[3] - b[0], b[4] - b[1], b[5] - b[2])):
            del tracker.instances[iid]


def match_replay(n_cc, cc_table, pair_keys, pair_vals, class_id, label_divisor, axis_name,
                 iou_thr=0.25, ioa_thr=0.25):
    """Runs forward + backward matching and tracker accumulation on host tables.

    n_cc      int32 [N]            components per slice
    cc_table  int32 [N, cap, 5]    area, y0, x0, y1, x1
    pair_keys uint64 [P]           slice << 40 | prev_cc << 20 | cur_cc
    pair_vals int32 [P]            overlapping pixels
    Returns (lut int32 [N, cap+1], labels int32 [n], sizes int64 [n], boxes int32 [n, 6]) with
    instances in tracker insertion order.
    """
    n_cc = np.ascontiguousarray(n_cc, dtype=np.int32)
    cc_table = np.ascontiguousarray(cc_table, dtype=np.int32)
    pair_keys = np.ascontiguousarray(pair_keys, dtype=np.uint64)
    pair_vals = np.ascontiguousarray(pair_vals, dtype=np.int32)
    n = int(n_cc.shape[0])
    cap = int(cc_table.shape[1])
    lut_stride = cap + 1
    lut = np.zeros((n, lut_stride), dtype=np.int32)
    max_inst = int(n_cc.sum()) + 1
    labels = np.zeros(max_inst, dtype=np.int32)
    sizes = np.zeros(max_inst, dtype=np.int64)
    boxes = np.zeros((max_inst, 6), dtype=np.int32)
    n_inst = np.zeros(1, dtype=np.int32)
    _lib.call("be_match_replay", n, _lib.ptr(n_cc), _lib.ptr(cc_table), cap, _lib.ptr(pair_keys),
              _lib.ptr(pair_vals), int(pair_keys.shape[0]), int(class_id), int(label_divisor),
              float(iou_thr), float(ioa_thr), AXIS_NUM[axis_name], _lib.ptr(lut), lut_stride,
              _lib.ptr(labels), _lib.ptr(sizes), _lib.ptr(boxes), max_inst, _lib.ptr(n_inst))
    k = int(n_inst[0])
    return lut, labels[:k], sizes[:k], boxes[:k]
