"""CPU restatement of the reference's slice-to-slice stitching: 2-D connected components + RLE
(empanada/inference/rle.py), Hungarian IoU matcher (matcher.py), forward/backward matching
(patterns.py), 3-D instance tracker (tracker.py) and the two always-on filters (filters.py).
TEST INFRASTRUCTURE ONLY (see oracle/post.py header).

Third-party arithmetic the reference calls and that is NOT vendored under /root/reference:
  * skimage.measure.label / regionprops (scikit-image >= 0.19, setup.cfg:50; absent here):
    re-stated in `connected_components` / `_regions` from the documented behaviour.
    No reference test pins it -> CC numbering is "parity unpinned" by the reference.
  * scipy.optimize.linear_sum_assignment (SciPy, unpinned; present here): called directly.
Pinned against the reference run in this container: tests/golden/track_*.npz, axis_*.npz.
"""
import math

import numpy as np
from scipy import ndimage as ndi
from scipy.optimize import linear_sum_assignment

from .ranges import (box_iou_pairs, merge_boxes, merge_rles, rle_decode, rle_encode, rle_ioa,
                     rle_iou, rle_to_string, string_to_rle)


# --------------------------------------------------------------------------- 2-D CC + RLE
def connected_components(seg):
    """rle.py:18-24 -> skimage.measure.label(seg): 8-connected components of EQUAL-valued
    non-zero pixels, numbered 1.. in raster order of each component's first pixel."""
    seg = np.asarray(seg)
    out = np.zeros(seg.shape, dtype=np.int64)
    structure = np.ones((3,) * seg.ndim, dtype=bool)
    first = []
    pieces = []
    for v in np.unique(seg):
        if v == 0:
            continue
        lab, n = ndi.label(seg == v, structure=structure)
        if n == 0:
            continue
        flat = lab.ravel()
        idx = np.flatnonzero(flat)
        # ndi.label itself numbers in raster order of first pixel
        fi = np.full(n + 1, -1, dtype=np.int64)
        rev = idx[::-1]
        fi[flat[rev]] = rev
        pieces.append((lab, n))
        for comp in range(1, n + 1):
            first.append((int(fi[comp]), len(pieces) - 1, comp))
    first.sort()
    luts = [np.zeros(n + 1, dtype=np.int64) for (_, n) in pieces]
    for new_id, (_, pi, comp) in enumerate(first, start=1):
        luts[pi][comp] = new_id
    for (lab, n), lut in zip(pieces, luts):
        m = lab > 0
        out[m] = lut[lab[m]]
    return out


def _regions(lab):
    """skimage.measure.regionprops(lab): ascending label; (label, bbox half-open, flat coords)."""
    flat = lab.ravel()
    idx = np.flatnonzero(flat)
    if idx.size == 0:
        return []
    vals = flat[idx]
    order = np.argsort(vals, kind="stable")
    idx_s = idx[order]
    vals_s = vals[order]
    bounds = np.flatnonzero(np.r_[True, vals_s[1:] != vals_s[:-1], True])
    w = lab.shape[1]
    regs = []
    for s, e in zip(bounds[:-1], bounds[1:]):
        f = idx_s[s:e]
        ys = f // w
        xs = f % w
        bbox = (int(ys.min()), int(xs.min()), int(ys.max()) + 1, int(xs.max()) + 1)
        regs.append((int(vals_s[s]), bbox, f))
    return regs


def pan_seg_to_rle_seg(pan_seg, labels, label_divisor, thing_list, force_connected=True):
    """rle.py:26-86."""
    rle_seg = {}
    for label in labels:
        min_id = label * label_divisor
        max_id = min_id + label_divisor
        ins = pan_seg.copy()
        ins[np.logical_or(pan_seg < min_id, pan_seg >= max_id)] = 0
        if force_connected and label in thing_list:
            ins = connected_components(ins)
            ins[ins > 0] += min_id
        attrs = {}
        for lab, bbox, flat in _regions(ins):
            starts, runs = rle_encode(flat)
            attrs[lab] = {"box": bbox, "starts": starts, "runs": runs}
        rle_seg[label] = attrs
    return rle_seg


def rle_seg_to_pan_seg(rle_seg, shape):
    """rle.py:88-118."""
    pan = np.zeros(shape, dtype=np.uint32).ravel()
    for attrs_by_id in rle_seg.values():
        for oid, attrs in attrs_by_id.items():
            for s, r in zip(attrs["starts"], attrs["runs"]):
                pan[s:s + r] = oid
    return pan.reshape(shape)


# --------------------------------------------------------------------------- matcher
def _merge_attrs(a, b):
    """matcher.py:14-28."""
    starts, runs = merge_rles(a["starts"], a["runs"], b["starts"], b["runs"])
    return {"box": merge_boxes(a["box"], b["box"]), "starts": starts, "runs": runs}


def rle_matcher(target, match, iou_thr):
    """matcher.py:136-232 with return_ioa=True."""
    t_labels = np.array([int(k) for k in target.keys()])
    m_labels = np.array([int(k) for k in match.keys()])
    t_attrs = list(target.values())
    m_attrs = list(match.values())
    if len(t_labels) == 0 or len(m_labels) == 0:
        empty = np.array([])
        return (empty, empty), (t_labels, m_labels), empty, empty
    iou = np.zeros((len(t_labels), len(m_labels)), dtype="float")
    ioa = np.zeros((len(t_labels), len(m_labels)), dtype=np.float32)
    pairs, _, _ = box_iou_pairs(np.array([a["box"] for a in t_attrs]),
                                np.array([a["box"] for a in m_attrs]))
    for r1, r2 in pairs:
        a, b = t_attrs[r1], m_attrs[r2]
        iou[r1, r2] = rle_iou(a["starts"], a["runs"], b["starts"], b["runs"])
        ioa[r1, r2] = rle_ioa(a["starts"], a["runs"], b["starts"], b["runs"])
    rows, cols = linear_sum_assignment(iou, maximize=True)
    keep = iou[rows, cols] >= iou_thr
    rows, cols = rows[keep], cols[keep]
    return (t_labels[rows], m_labels[cols]), (t_labels, m_labels), iou[(rows, cols)], ioa


class RLEMatcher:
    """matcher.py:234-326."""

    def __init__(self, class_id, label_divisor, merge_iou_thr=0.25, merge_ioa_thr=0.25,
                 assign_new=True):
        self.class_id = class_id
        self.label_divisor = label_divisor
        self.merge_iou_thr = merge_iou_thr
        self.merge_ioa_thr = merge_ioa_thr
        self.assign_new = assign_new
        self.next_label = class_id * label_divisor + 1
        self.target_rle = None

    def initialize_target(self, target):
        self.target_rle = target
        objs = list(target.keys())
        if len(objs) > 0:
            self.next_label = max(objs) + 1

    def __call__(self, match):
        assert self.target_rle is not None
        matched, all_labels, _, ioa = rle_matcher(self.target_rle, match, self.merge_iou_thr)
        t_labels, m_labels = all_labels
        label_matches = {ml: tl for tl, ml in zip(matched[0], matched[1])}
        out = {}
        for i, (ml, attrs) in enumerate(match.items()):
            if ml in label_matches:
                new_label = label_matches[ml]
            else:
                assert ml == m_labels[i]
                ioa_max = ioa[:, i].max() if len(ioa) > 0 else 0
                if ioa_max >= self.merge_ioa_thr:
                    new_label = t_labels[ioa[:, i].argmax()]
                elif self.assign_new:
                    new_label = self.next_label
                    self.next_label += 1
                else:
                    new_label = ml
            new_label = int(new_label)
            if new_label not in out:
                out[new_label] = attrs
            else:
                out[new_label] = _merge_attrs(out[new_label], attrs)
        self.target_rle = out
        return out


def apply_matchers(rle_seg, matchers):
    """patterns.py:55-66."""
    for m in matchers:
        if m.target_rle is None:
            m.initialize_target(rle_seg[m.class_id])
        else:
            rle_seg[m.class_id] = m(rle_seg[m.class_id])
    return rle_seg


def forward_matching(pan_segs, matchers, labels, label_divisor, thing_list):
    """patterns.py:68-100 (the child process loop) over already-emitted pan_segs."""
    stack = []
    for pan in pan_segs:
        if pan is None:
            continue
        rs = pan_seg_to_rle_seg(pan, labels, label_divisor, thing_list, force_connected=True)
        stack.append(apply_matchers(rs, matchers))
    return stack


def backward_matching(rle_stack, matchers, axis_len):
    """patterns.py:102-121."""
    for m in matchers:
        m.target_rle = None
        m.assign_new = False
    for rev in range(axis_len - 1, -1, -1):
        yield rev, apply_matchers(rle_stack[rev], matchers)


# --------------------------------------------------------------------------- tracker
def to_box3d(i, box, axis):
    """tracker.py:11-23."""
    h1, w1, h2, w2 = box
    if axis == "xy":
        return (i, h1, w1, i + 1, h2, w2)
    if axis == "xz":
        return (h1, i, w1, h2, i + 1, w2)
    return (h1, w1, i, h2, w2, i + 1)


class InstanceTracker:
    """tracker.py:40-159."""

    def __init__(self, class_id=None, label_divisor=None, shape3d=None, axis="xy"):
        assert axis in ["xy", "xz", "yz"]
        self.class_id = class_id
        self.label_divisor = label_divisor
        self.shape3d = shape3d
        self.axis = axis
        self.finished = False
        self.instances = {}
        self.axis_nums = {"xy": 0, "xz": 1, "yz": 2}

    def update(self, instance_rles, index2d):
        assert not self.finished
        ignore = self.axis_nums[self.axis]
        shape2d = tuple(s for i, s in enumerate(self.shape3d) if i != ignore)
        for label, attrs in instance_rles.items():
            box = to_box3d(index2d, attrs["box"], self.axis)
            if self.axis == "xy":
                starts = attrs["starts"] + index2d * math.prod(shape2d)
                runs = attrs["runs"]
            elif self.axis == "xz":
                hc, wc = np.unravel_index(attrs["starts"], shape2d)
                dc = np.repeat([index2d], len(hc))
                starts = np.ravel_multi_index((hc, dc, wc), self.shape3d)
                runs = attrs["runs"]
            else:
                flat = rle_decode(attrs["starts"], attrs["runs"])
                hc, wc = np.unravel_index(flat, shape2d)
                dc = np.repeat([index2d], len(hc))
                starts = np.ravel_multi_index((hc, wc, dc), self.shape3d)
                runs = np.ones_like(starts)
            if label not in self.instances:
                self.instances[label] = {"box": box, "starts": [starts], "runs": [runs]}
            else:
                d = self.instances[label]
                d["box"] = merge_boxes(box, d["box"])
                d["starts"].append(starts)
                d["runs"].append(runs)

    def finish(self):
        for iid in self.instances.keys():
            d = self.instances[iid]
            if isinstance(d["starts"], list):
                starts = np.concatenate(d["starts"])
                if self.axis == "yz":
                    starts, runs = rle_encode(np.sort(starts, kind="stable"))
                else:
                    runs = np.concatenate(d["runs"])
                d["starts"] = starts
                d["runs"] = runs
        self.finished = True


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


def instance_relabel(tracker):
    """empanada_napari/inference.py:31-54."""
    out = {}
    iid = 1
    for attrs in tracker.instances.values():
        cat = np.stack([attrs["starts"], attrs["runs"]], axis=1)
        cat = cat[np.argsort(cat[:, 0], kind="stable")]
        out[iid] = {"box": attrs["box"], "starts": cat[:, 0], "runs": cat[:, 1]}
        iid += 1
    return out


# --------------------------------------------------------------------------- tracker morphology
def rle_seg_to_pan_seg3d(tracker, shape):
    """filters.py:120-152: every instance's runs painted in dictionary order (uint32 volume)."""
    pan = np.zeros(shape, dtype=np.uint32).ravel()
    for object_id, attrs in tracker.instances.items():
        for s, r in zip(attrs["starts"], attrs["runs"]):
            pan[s:s + r] = object_id
    return pan.reshape(shape)


def pan_seg_to_rle_seg3d(pan_seg, labels, label_divisor, thing_list, force_connected=True):
    """filters.py:58-118 (n-dimensional form of rle.py:26-86; returns the flat instance dict)."""
    instance_attrs = {}
    for label in labels:
        lo = label * label_divisor
        hi = lo + label_divisor
        ins = pan_seg.copy()
        ins[np.logical_or(pan_seg < lo, pan_seg >= hi)] = 0
        if force_connected and label in thing_list:
            ins = connected_components(ins)
            ins[ins > 0] += lo
        flat = ins.ravel()
        idx = np.flatnonzero(flat)
        vals = flat[idx]
        order = np.argsort(vals, kind="stable")
        idx_s, vals_s = idx[order], vals[order]
        bounds = np.flatnonzero(np.r_[True, vals_s[1:] != vals_s[:-1], True]) if len(vals_s) else np.zeros(1, int)
        for a, b in zip(bounds[:-1], bounds[1:]):
            coords = np.unravel_index(idx_s[a:b], ins.shape)
            box = tuple(int(c.min()) for c in coords) + tuple(int(c.max()) + 1 for c in coords)
            starts, runs = rle_encode(idx_s[a:b])
            instance_attrs[int(vals_s[a])] = {"box": box, "starts": starts, "runs": runs}
    return instance_attrs


def _cross(ndim):
    from scipy import ndimage as ndi
    return ndi.generate_binary_structure(ndim, 1)


def erode(tracker, shape, labels, label_divisor, thing_list, iterations=1):
    """filters.py:154-163; skimage.morphology.erosion with the default footprint = grey erosion
    with the cross, borders reflected (third party, re-stated: see oracle/ref_shim.py)."""
    from scipy import ndimage as ndi
    mask = rle_seg_to_pan_seg3d(tracker, shape)
    for _ in range(iterations):
        mask = ndi.grey_erosion(mask, footprint=_cross(mask.ndim))
    tracker.instances = pan_seg_to_rle_seg3d(mask, labels, label_divisor, thing_list)
    return tracker


def dilate(tracker, shape, labels, label_divisor, thing_list, iterations=1):
    """filters.py:165-172."""
    from scipy import ndimage as ndi
    mask = rle_seg_to_pan_seg3d(tracker, shape)
    for _ in range(iterations):
        mask = ndi.grey_dilation(mask, footprint=_cross(mask.ndim))
    tracker.instances = pan_seg_to_rle_seg3d(mask, labels, label_divisor, thing_list)
    return tracker


def fill_holes_in_segmentation(tracker, shape, labels, label_divisor, thing_list):
    """filters.py:174-210: slice by slice along axis 0, labels ascending, each inside its bounding
    box of the unmodified slice; every non-zero pixel of the box and every hole takes the label."""
    from scipy.ndimage import binary_fill_holes
    mask3d = rle_seg_to_pan_seg3d(tracker, shape)
    for z in range(mask3d.shape[0]):
        mask = mask3d[z]
        boxes = []
        for l in np.unique(mask):
            if l > 0:
                yy, xx = np.nonzero(mask == l)
                boxes.append((l, yy.min(), xx.min(), yy.max() + 1, xx.max() + 1))
        for l, y0, x0, y1, x1 in boxes:
            tmp = binary_fill_holes(mask[y0:y1, x0:x1].astype(bool))
            mask[y0:y1, x0:x1] = tmp.astype(mask.dtype) * l
        mask3d[z] = mask
    tracker.instances = pan_seg_to_rle_seg3d(mask3d, labels, label_divisor, thing_list)
    return tracker
