"""CPU restatement (numpy) of the reference's per-slice post-processing.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg. The product package never imports anything under oracle/.

Pinned against the reference itself: tests/golden/post_*.npz are produced by
oracle/make_golden.py, which runs the UNMODIFIED reference functions
(empanada/inference/postprocess.py, engines.py) in the build container.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import math

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- input side
def normalize(img, mean, std):
    """empanada_napari/utils.py:170-201 (Preprocessor / normalize): integer image ->
    fp32 `(img - fp32(mean)*max) * (1 / (fp32(std)*max))`, all in fp32."""
    if np.issubdtype(img.dtype, np.floating):
        raise Exception("Input image cannot be float type!")
    max_value = np.iinfo(img.dtype).max
    m = np.array(mean, dtype=F32)
    m *= max_value
    s = np.array(std, dtype=F32)
    s *= max_value
    den = np.reciprocal(s, dtype=F32)
    out = img.astype(F32)
    out -= m
    out *= den
    return out


def factor_pad(img, factor):
    """empanada/inference/postprocess.py:26-36: zero-pad bottom/right (after normalisation)."""
    h, w = img.shape[-2:]
    pb = factor - h % factor if h % factor != 0 else 0
    pr = factor - w % factor if w % factor != 0 else 0
    if pb == 0 and pr == 0:
        return img
    pad = [(0, 0)] * (img.ndim - 2) + [(0, pb), (0, pr)]
    return np.pad(img, pad)


def sigmoid(x):
    """engines.py:28 torch.sigmoid in fp32 (tolerance-checked, not bit-exact)."""
    x = x.astype(F32)
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


# --------------------------------------------------------------------------- median queue
class MedianQueue:
    """empanada/inference/engines.py:47-90 (_MedianQueue) + the push/emit logic of
    PanopticDeepLabRenderEngine3d.__call__/end (engines.py:351-394).

    push(item) returns the item to post-process now (or None while the queue fills);
    the median REPLACES item['sem'] of the queued middle element (recursive filter)."""

    def __init__(self, ks):
        assert ks % 2 == 1, "Kernel size must be odd integer!"
        self.ks = ks
        self.mid = (ks - 1) // 2
        self.q = []

    def reset(self):
        self.q = []

    def push(self, item):
        self.q.append(item)
        if len(self.q) > self.ks:
            self.q.pop(0)
        nq = len(self.q)
        if nq <= self.mid:
            return self.q[-1]
        if nq < self.ks:
            return None
        out = self.q[self.mid]
        stack = np.stack([it["sem"] for it in self.q], axis=0)
        # torch.median over an odd count = middle order statistic
        out["sem"] = np.sort(stack, axis=0)[self.mid]
        return out

    def end(self):
        return list(self.q)[self.mid + 1:]


# --------------------------------------------------------------------------- centres
def find_instance_center(ctr_hmp, threshold=0.1, nms_kernel=7):
    """postprocess.py:39-76. ctr_hmp (h, w) fp32 -> (K, 2) int64 (y, x), row-major order."""
    x = ctr_hmp.astype(F32)
    t = np.where(x > F32(threshold), x, F32(-1.0))
    k = int(nms_kernel)
    pad = k // 2
    h, w = t.shape
    padded = np.full((h + 2 * pad, w + 2 * pad), -np.inf, dtype=F32)
    padded[pad:pad + h, pad:pad + w] = t
    # stride-1 max pool; for even k the output is (h+1, w+1) and the last row/col is dropped
    pooled = np.full((h, w), -np.inf, dtype=F32)
    for dy in range(k):
        for dx in range(k):
            pooled = np.maximum(pooled, padded[dy:dy + h, dx:dx + w])
    t = np.where(t != pooled, F32(-1.0), t)
    return np.argwhere(t > 0).astype(np.int64)


def _dist(cy, cx, ly, lx):
    """torch.norm(ctr - ctr_loc, dim=-1) on CPU: sqrt(fma(dx, dx, fl(dy*dy))) in fp32
    (pinned by tests/golden/post_*.npz; see DESIGN.md 'grouping arithmetic')."""
    dy = (cy - ly).astype(F32)
    dx = (cx - lx).astype(F32)
    dy2 = (dy * dy).astype(F32)
    acc = (dy2.astype(np.float64) + dx.astype(np.float64) * dx.astype(np.float64)).astype(F32)
    return np.sqrt(acc, dtype=F32)


def group_pixels(ctr, offsets, step=1, chunksize=20):
    """postprocess.py:79-169. ctr (K,2) int64, offsets (2,h,w) fp32 -> (h,w) int64 ids."""
    assert ctr.shape[0] > 0
    _, h, w = offsets.shape
    ys = (np.arange(h, dtype=F32) * F32(step)).astype(F32)
    xs = (np.arange(w, dtype=F32) * F32(step)).astype(F32)
    ly = (ys[:, None] + offsets[0].astype(F32)).astype(F32).reshape(-1)
    lx = (xs[None, :] + offsets[1].astype(F32)).astype(F32).reshape(-1)
    c = (F32(step) * ctr.astype(F32)).astype(F32)
    K = c.shape[0]
    if K <= chunksize:
        d = _dist(c[:, 0:1], c[:, 1:2], ly[None, :], lx[None, :])
        ids = 1 + np.argmin(d, axis=0)
    else:
        ids = np.zeros(h * w, dtype=np.int64)
        nearest = np.full(h * w, F32(1e5), dtype=F32)
        prev = 1
        for s in range(0, K, chunksize):
            cc = c[s:s + chunksize]
            d = _dist(cc[:, 0:1], cc[:, 1:2], ly[None, :], lx[None, :])
            mind = d.min(axis=0)
            arg = d.argmin(axis=0)
            upd = mind < nearest
            ids[upd] = prev + arg[upd]
            nearest = np.minimum(nearest, mind)
            prev += cc.shape[0]
    return ids.reshape(h, w).astype(np.int64)


def get_instance_cells(ctr_hmp, offsets, nms_threshold, nms_kernel, coarse_boundaries=True,
                       upsampling=1):
    """engines.py:258-275. ctr_hmp (h4,w4), offsets (2,h4,w4) -> (H,W) fp32 ids."""
    ctr = find_instance_center(ctr_hmp, nms_threshold, nms_kernel)
    step = 4 if coarse_boundaries else 1
    if ctr.shape[0] == 0:
        cells = np.zeros(ctr_hmp.shape, dtype=F32)
    else:
        cells = group_pixels(ctr, offsets, step=step).astype(F32)
    scale = int(upsampling * step)
    if scale != 1:
        cells = np.repeat(np.repeat(cells, scale, axis=0), scale, axis=1)  # nearest, integer scale
    return cells


# --------------------------------------------------------------------------- harden + merge
def harden_seg(sem, confidence_thr):
    """engines.py:115-121. sem (C,H,W) probabilities -> (H,W) int64 class map."""
    if sem.shape[0] > 1:
        return np.argmax(sem, axis=0).astype(np.int64)
    return (sem[0] >= F32(confidence_thr)).astype(np.int64)


def merge_semantic_and_instance(sem_seg, ins_seg, label_divisor, thing_list, stuff_area, void_label):
    """postprocess.py:224-296 on (H,W) int64 arrays."""
    pan = np.zeros_like(sem_seg) + void_label
    thing_seg = ins_seg > 0
    sem_thing = np.zeros_like(sem_seg)
    for c in thing_list:
        sem_thing[sem_seg == c] = 1
    tracker = {}
    for ins_id in np.unique(ins_seg):
        if ins_id == 0:
            continue
        mask = (ins_seg == ins_id) & (sem_thing == 1)
        if not mask.any():
            continue
        vals, counts = np.unique(sem_seg[mask], return_counts=True)
        class_id = int(vals[np.argmax(counts)])  # torch.mode: smallest of the most frequent
        if class_id in tracker:
            new_id = tracker[class_id]
        else:
            tracker[class_id] = 1
            new_id = 1
        tracker[class_id] += 1
        pan[mask] = class_id * label_divisor + new_id
    for class_id in np.unique(sem_seg):
        if int(class_id) in thing_list:
            continue
        smask = (sem_seg == class_id) & (~thing_seg)
        if int(smask.sum()) >= stuff_area:
            pan[smask] = class_id * label_divisor
    return pan


def get_panoptic_seg(sem_hard, instance_cells, label_divisor, thing_list, stuff_area=64,
                     void_label=0):
    """engines.py:278-292."""
    ins = np.zeros_like(sem_hard)
    for c in thing_list:
        ins[sem_hard == c] = 1
    ins = (ins.astype(F32) * instance_cells.astype(F32)).astype(np.int64)
    return merge_semantic_and_instance(sem_hard, ins, label_divisor, thing_list, stuff_area,
                                       void_label)


class RenderEnginePost:
    """Post-model half of PanopticDeepLabRenderEngine(3d) (engines.py:223-394): consumes the
    model's head outputs for one padded slice and emits pan_seg (h, w) int64 or None."""

    def __init__(self, thing_list, label_divisor=1000, stuff_area=64, void_label=0,
                 nms_threshold=0.1, nms_kernel=7, confidence_thr=0.5, median_kernel_size=None,
                 coarse_boundaries=True):
        self.thing_list = list(thing_list)
        self.label_divisor = label_divisor
        self.stuff_area = stuff_area
        self.void_label = void_label
        self.nms_threshold = nms_threshold
        self.nms_kernel = nms_kernel
        self.confidence_thr = confidence_thr
        self.coarse_boundaries = coarse_boundaries
        self.queue = MedianQueue(median_kernel_size) if median_kernel_size else None

    def _finish(self, item, upsampling=1):
        h, w = item["size"]
        cells = get_instance_cells(item["ctr_hmp"], item["offsets"], self.nms_threshold,
                                   self.nms_kernel, self.coarse_boundaries, upsampling)
        hard = harden_seg(item["sem"], self.confidence_thr)
        pan = get_panoptic_seg(hard, cells, self.label_divisor, self.thing_list,
                               self.stuff_area, self.void_label)
        return pan[:h, :w]

    def __call__(self, sem, ctr_hmp, offsets, size, upsampling=1):
        """sem (C,H,W) probabilities, ctr_hmp (h4,w4), offsets (2,h4,w4)."""
        item = {"sem": sem, "ctr_hmp": ctr_hmp, "offsets": offsets, "size": size}
        if self.queue is None:
            return self._finish(item, upsampling)
        out = self.queue.push(item)
        if out is None:
            return None
        return self._finish(out, upsampling)

    def end(self, upsampling=1):
        if self.queue is None:
            return []
        return [self._finish(it, upsampling) for it in self.queue.end()]

    def reset(self):
        if self.queue is not None:
            self.queue.reset()
