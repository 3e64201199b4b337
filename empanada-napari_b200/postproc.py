"""Per-plane GPU post-processing driver: median queue -> centres -> grouping -> merge ->
connected components -> overlap tables -> (host matcher replay) -> relabel -> RLE runs.

Everything between the model's head outputs and the tracker dictionaries of
`Engine3d.infer_on_axis` (empanada_napari/inference.py:526-578) for ONE plane, batched over
slices. torch is used for device buffers and trivial index plumbing; every per-pixel pass is
one of the library's own kernels (csrc/post_kernels.cu, cc_kernels.cu).
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from . import tracking

AXES = {"xy": 0, "xz": 1, "yz": 2}


class CenterOverflow(_lib.B200EmpanadaError):
    """More centres in one slice than `center_cap` slots: the plane is re-run with `needed`."""

    def __init__(self, needed, cap):
        super().__init__(f"{needed} centres in one slice exceed center_cap={cap}")
        self.needed = needed


def _next_pow2(n):
    p = 1
    while p < n:
        p <<= 1
    return p


class PlanePost:
    """State for one `infer_on_axis` pass over `n_slices` slices of size (h, w), padded (H, W).

    Stage 1 (per batch of slices, while the network runs): recursive median + harden -> `hard`,
    centre NMS -> `centers`; the offsets are kept (`off_all`) because grouping is deferred until the
    filtered mask is known and then evaluated only where it matters (csrc/run_kernels.cu).
    Stage 2 (`run_cc`, whole plane): grouping + presence flags, renumbering, row-run extraction,
    union-find over runs, component tables, adjacent-slice overlap table.
    Stage 3: host matcher replay on the tables (tracking.match_replay).
    Stage 4 (`relabel`): the (D,H,W) label volume painted once from the runs."""

    def __init__(self, n_slices, h, w, H, W, *, ks, thing_class, label_divisor, void_label=0,
                 nms_threshold=0.1, nms_kernel=3, confidence_thr=0.5, scale=4, device="cuda:0",
                 center_cap=4096, cc_cap=None, keep_prob=False, hash_cap=None, step=None,
                 semantic=False, stuff_area=64):
        """`scale`: output pixels per head-grid cell (grid step x upsampling, engines.py:263-275);
        `step`: grid step of the head maps in model pixels (4, or 1 with fine boundaries; default
        = scale). `semantic`: the class is a stuff class (engine thing_list = [], 'semantic
        only'): no instances, the whole mask of a slice is one segment if it covers at least
        `stuff_area` pixels of the padded slice (postprocess.py:283-294)."""
        if ks % 2 != 1:
            raise AssertionError("Kernel size must be odd integer!")
        if n_slices < ks:
            raise ValueError(f"stack of {n_slices} slices is shorter than the median kernel ({ks})")
        self.N, self.h, self.w, self.H, self.W = n_slices, h, w, H, W
        self.ks, self.mid = ks, (ks - 1) // 2
        self.cls, self.div, self.void = thing_class, label_divisor, void_label
        self.thr, self.k, self.conf, self.scale = nms_threshold, nms_kernel, confidence_thr, scale
        self.step = scale if step is None else step
        self.semantic, self.stuff_area = semantic, stuff_area
        self.dev = torch.device(device)
        self.h4, self.w4 = H // scale, W // scale
        self.center_cap = center_cap
        self.hash_cap = hash_cap            # initial overlap-table capacity (tests force regrowth)
        d = self.dev
        self.hard = torch.zeros((n_slices, H, W), dtype=torch.uint8, device=d)
        self.prob = torch.zeros((n_slices, H, W), dtype=torch.float32, device=d) if keep_prob else None
        self.off_all = torch.empty((n_slices, 2, self.h4, self.w4), dtype=torch.float32, device=d)
        self.hist = torch.zeros((max(ks - 1, 1), H, W), dtype=torch.float32, device=d)
        self.n_hist = 0
        self.pushed = 0
        self.centers = torch.zeros((n_slices, center_cap), dtype=torch.int32, device=d)
        self.center_counts = torch.zeros(n_slices, dtype=torch.int32, device=d)
        self.cells4 = None
        self.newid = None
        self.runs = None
        self.launches = 0

    # ------------------------------------------------------------------ stage 1: heads in
    def _centers(self, ctr, off, i0):
        B = ctr.shape[0]
        st = stream_ptr()
        scratch = torch.empty(B * ((self.h4 * self.w4 + 1023) // 1024), dtype=torch.int32, device=self.dev)
        call("be_centers", ptr(ctr), B, self.h4, self.w4, float(self.thr), int(self.k),
             ptr(self.centers[i0]), self.center_cap, ptr(self.center_counts[i0:]), ptr(scratch), st)
        self.off_all[i0:i0 + B].copy_(off)
        self.launches += 3

    def push_heads(self, sem, ctr, off, is_prob=False):
        """sem (B,H,W) fp32 logits (or probabilities), ctr (B,h4,w4), off (B,2,h4,w4)."""
        B = sem.shape[0]
        s0 = self.pushed
        assert s0 + B <= self.N
        self.push_semantic(sem, s0, is_prob)
        self._centers(ctr, off, s0)
        self.pushed += B

    # split form of push_heads for slice-sharded planes (multigpu.ShardedEngine3d): the instance
    # heads of a slice and its (recursively median-filtered) semantic slice arrive separately
    def push_instance(self, ctr, off, i0):
        """Centres (and kept offsets) of B slices into local slots [i0, i0 + B)."""
        assert i0 >= 0 and i0 + ctr.shape[0] <= self.N
        self._centers(ctr, off, i0)

    def push_semantic(self, sem, slot0, is_prob=False):
        """Median queue push of B slices; the slice pushed at slot t emits slot t - mid."""
        B = sem.shape[0]
        call("be_median_push", ptr(sem), B, self.H, self.W, self.ks, ptr(self.hist), self.n_hist,
             slot0, float(self.conf), int(is_prob), ptr(self.hard), ptr(self.prob), stream_ptr())
        self.launches += 1
        self.n_hist = min(self.n_hist + B, self.ks - 1)

    def flush_semantic(self):
        """End of the stack: the queue tail is emitted unfiltered into the last slots."""
        if self.ks > 1:
            call("be_median_flush", ptr(self.hist), self.n_hist, self.ks, self.H, self.W, self.N,
                 float(self.conf), ptr(self.hard), ptr(self.prob), stream_ptr())
            self.launches += 1

    def check_centers(self):
        if self.semantic:
            return
        cmax = int(self.center_counts.max().item())
        if cmax > self.center_cap:
            raise CenterOverflow(cmax, self.center_cap)

    def finish_heads(self):
        assert self.pushed == self.N
        self.flush_semantic()
        self.check_centers()

    # ------------------------------------------------------------------ stage 2: groups -> runs -> cc
    def group(self):
        """Nearest-centre ids for every head-grid cell that holds a thing pixel + dense
        renumbering table (`newid`, the per-class running counter of postprocess.py:273-288)."""
        if self.cells4 is not None:
            return
        N, d = self.N, self.dev
        if self.semantic:
            # one "cell id" everywhere; it maps to class * divisor where the slice's mask is large enough
            self.cells4 = torch.ones((N, self.h4, self.w4), dtype=torch.int32, device=d)
            self.newid = torch.zeros((N, self.center_cap + 1), dtype=torch.int32, device=d)
            area = torch.zeros(N, dtype=torch.int32, device=d)
            call("be_slice_area", ptr(self.hard), N, self.H, self.W, ptr(area), stream_ptr())
            self.newid[:, 1] = torch.where(area >= int(self.stuff_area), int(self.cls * self.div), int(self.void))
            self.launches += 1
            return
        self.cells4 = torch.zeros((N, self.h4, self.w4), dtype=torch.int32, device=d)
        self.newid = torch.zeros((N, self.center_cap + 1), dtype=torch.int32, device=d)
        call("be_group_flags", ptr(self.hard), ptr(self.off_all), ptr(self.centers), self.center_cap,
             ptr(self.center_counts), N, self.H, self.W, self.scale, self.step, ptr(self.cells4), ptr(self.newid),
             stream_ptr())
        call("be_rank_ids", ptr(self.newid), N, self.center_cap, int(self.div), int(self.cls), stream_ptr())
        self.launches += 2

    def group_all(self, s0=0, s1=None):
        """Reference-complete `get_instance_cells` output (every cell, background included) for
        slices [s0, s1): used by tests and the per-slice engine API, not by the plane path."""
        s1 = self.N if s1 is None else s1
        out = torch.empty((s1 - s0, self.h4, self.w4), dtype=torch.int32, device=self.dev)
        call("be_group_pixels", ptr(self.off_all[s0]), ptr(self.centers[s0]), self.center_cap,
             ptr(self.center_counts[s0:]), s1 - s0, self.h4, self.w4, float(self.step), ptr(out), stream_ptr())
        return out

    def pan_batch(self, s0, s1):
        """Dense pan_seg (B,h,w) int32 for emitted slices [s0, s1) (the per-slice engine output);
        the plane path itself never materialises it."""
        B = s1 - s0
        cells = self.group_all(s0, s1)
        pan = torch.empty((B, self.h, self.w), dtype=torch.int32, device=self.dev)
        present = torch.empty((B, self.center_cap + 1), dtype=torch.int32, device=self.dev)
        call("be_merge_pan", ptr(self.hard[s0]), ptr(cells), B, self.H, self.W, self.h,
             self.w, self.scale, self.center_cap, int(self.div), int(self.cls), int(self.void),
             ptr(present), ptr(pan), stream_ptr())
        self.launches += 4
        return pan

    def run_cc(self, batch=None):
        """Row runs, connected components + tables for all slices; overlap table between
        neighbouring slices. (`batch` is accepted for compatibility; the run pipeline handles the
        whole plane in one launch sequence.)"""
        N, h, w, d = self.N, self.h, self.w, self.dev
        if self.runs is not None:      # already done (the cell ids were consumed by the first call)
            return
        self.group()
        st = stream_ptr()
        gpr = (w + 15) // 16                 # one thread per 16 pixels of a row (csrc/run_kernels.cu)
        chunks = (h * gpr + 255) // 256
        counts = torch.empty(2 * N * chunks, dtype=torch.int32, device=d)
        n_runs = torch.empty(N, dtype=torch.int32, device=d)
        slice_off = torch.empty(N + 1, dtype=torch.int32, device=d)
        stats = torch.empty(2, dtype=torch.int32, device=d)
        row_ptr = torch.empty((N, h + 1), dtype=torch.int32, device=d)
        lo, hi = self.cls * self.div, (self.cls + 1) * self.div
        src = (ptr(self.hard), ptr(self.cells4), ptr(self.newid), N, self.H, self.W, h, w, self.scale,
               self.center_cap, int(self.void), lo, hi)
        call("be_rowruns_count", *src, ptr(counts), ptr(n_runs), ptr(slice_off), ptr(stats), ptr(row_ptr), st)
        total, max_runs = (int(v) for v in stats.cpu().numpy())
        R = max(total, 1)
        run_yx = torch.empty((R, 2), dtype=torch.int32, device=d)
        run_x1 = torch.empty(R, dtype=torch.int32, device=d)
        run_val = torch.empty(R, dtype=torch.int32, device=d)
        L = torch.empty(R, dtype=torch.int32, device=d)
        run_cc = torch.zeros(R, dtype=torch.int32, device=d)
        self.n_cc = torch.empty(N, dtype=torch.int32, device=d)
        call("be_rowruns_write", *src, ptr(counts), ptr(slice_off), ptr(row_ptr), ptr(run_yx), ptr(run_x1),
             ptr(run_val), ptr(L), st)
        call("be_runs_cc", ptr(row_ptr), ptr(run_yx), ptr(run_x1), ptr(run_val), ptr(slice_off), N, h,
             max_runs, ptr(L), ptr(run_cc), ptr(self.n_cc), st)
        self.launches += 7
        self.runs = dict(row_ptr=row_ptr, yx=run_yx, x1=run_x1, cc=run_cc, slice_off=slice_off,
                         total=total, max_runs=max_runs)
        n_cc_host = self.n_cc.cpu().numpy()
        self._n_cc_host = n_cc_host
        n_cc_max = int(n_cc_host.max()) if N else 0
        if n_cc_max >= (1 << 20):
            raise _lib.B200EmpanadaError("more than 2^20 components in one slice")
        self.cc_cap = max(1, n_cc_max)
        self.cc_table = torch.empty((N, self.cc_cap, 5), dtype=torch.int32, device=d)
        call("be_runs_stats", ptr(run_yx), ptr(run_x1), ptr(run_cc), ptr(slice_off), N, max_runs,
             self.cc_cap, ptr(self.cc_table), st)
        self.launches += 2
        # overlap table (slice s vs s-1)
        total_cc = int(n_cc_host.sum())
        cap = self.hash_cap or _next_pow2(max(1 << 16, 8 * total_cc))
        while True:
            keys = torch.empty(cap, dtype=torch.int64, device=d)
            vals = torch.empty(cap, dtype=torch.int32, device=d)
            overflow = torch.zeros(1, dtype=torch.int32, device=d)
            call("be_hash_clear", ptr(keys), ptr(vals), cap, st)
            call("be_runs_overlap", ptr(row_ptr), ptr(run_yx), ptr(run_x1), ptr(run_cc), ptr(slice_off), N, h,
                 max_runs, 0, ptr(keys), ptr(vals), cap, ptr(overflow), st)
            self.launches += 3
            if int(overflow.item()) == 0:
                break
            cap *= 4
        self.pair_keys, self.pair_vals = self._compact(keys, vals, cap)

    def _compact(self, keys, vals, cap):
        d = self.dev
        out_keys = torch.empty(cap, dtype=torch.int64, device=d)
        out_vals = torch.empty(cap, dtype=torch.int32, device=d)
        cursor = torch.zeros(1, dtype=torch.int32, device=d)
        call("be_hash_compact", ptr(keys), ptr(vals), cap, ptr(out_keys), ptr(out_vals), cap,
             ptr(cursor), stream_ptr())
        self.launches += 1
        n_pairs = int(cursor.item())
        return out_keys[:n_pairs].cpu().numpy().view(np.uint64), out_vals[:n_pairs].cpu().numpy()

    def cc_images(self, s0, s1, add=0):
        """Dense (B,h,w) int32 images of the component ids (+ `add` on labelled pixels) of slices
        [s0, s1): Engine2d output, shard-boundary exchange, tests."""
        r = self.runs
        out = torch.empty((s1 - s0, self.h, self.w), dtype=torch.int32, device=self.dev)
        call("be_runs_paint", ptr(r["row_ptr"]), ptr(r["yx"]), ptr(r["x1"]), ptr(r["cc"]), ptr(r["slice_off"]),
             None, 0, int(add), s0, s1 - s0, self.h, self.w,
             ptr(out) - 4 * s0 * self.h * self.w, self.h * self.w, self.w, 1, stream_ptr())
        self.launches += 1
        return out

    def boundary_pairs(self, prev_cc):
        """Overlap pairs between `prev_cc` (the last component slice of the previous shard, (h, w)
        int32 device tensor) and this shard's first slice, keyed as local slice 0."""
        d, h, w = self.dev, self.h, self.w
        two = torch.cat([prev_cc[None], self.cc_images(0, 1)]).contiguous()
        cap = 1 << 16
        while True:
            keys = torch.empty(cap, dtype=torch.int64, device=d)
            vals = torch.empty(cap, dtype=torch.int32, device=d)
            overflow = torch.zeros(1, dtype=torch.int32, device=d)
            call("be_hash_clear", ptr(keys), ptr(vals), cap, stream_ptr())
            call("be_pair_overlap", ptr(two), h, w, 1, 2, ptr(keys), ptr(vals), cap, ptr(overflow), stream_ptr())
            self.launches += 3
            if int(overflow.item()) == 0:
                break
            cap *= 4
        k, v = self._compact(keys, vals, cap)
        k = k - (np.uint64(1) << np.uint64(40))      # slice index 1 of the pair buffer -> local slice 0
        return k, v

    # ------------------------------------------------------------------ stage 3: host replay
    def replay_inputs(self):
        """Host copies of the component tables (synchronises the stream)."""
        n_cc = self._n_cc_host
        return n_cc, self.cc_table.cpu().numpy()

    def replay_host(self, inputs, axis_name, iou_thr=0.25, ioa_thr=0.25):
        """Pure host part (native code, releases the GIL): safe to run on a worker thread."""
        n_cc, table = inputs
        return tracking.match_replay(n_cc, table, self.pair_keys, self.pair_vals, self.cls,
                                     self.div, axis_name, iou_thr, ioa_thr)

    def replay(self, axis_name, iou_thr=0.25, ioa_thr=0.25):
        return self.replay_host(self.replay_inputs(), axis_name, iou_thr, ioa_thr)

    def semantic_tables(self, axis_name):
        """Tracker tables of a stuff class, in the format of `replay`: no matching takes place
        (patterns.py:55-66 only matches thing classes); every slice contributes its whole mask to
        the single label class * divisor (rle.py:60-83 without connected components,
        tracker.py:61-105)."""
        n_cc, table = self.replay_inputs()
        label = int(self.cls * self.div)
        lut = np.zeros((self.N, table.shape[1] + 1), dtype=np.int32)
        lut[:, 1:] = label
        has = np.flatnonzero(n_cc > 0)
        if has.size == 0:
            return lut, np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros((0, 6), np.int32)
        valid = np.arange(table.shape[1])[None, :] < n_cc[:, None]
        size = int(table[..., 0][valid].sum())
        big = np.iinfo(np.int32).max
        y0 = int(np.where(valid, table[..., 1], big).min()); x0 = int(np.where(valid, table[..., 2], big).min())
        y1 = int(np.where(valid, table[..., 3], -1).max()); x1 = int(np.where(valid, table[..., 4], -1).max())
        i0, i1 = int(has[0]), int(has[-1]) + 1
        box = {"xy": (i0, y0, x0, i1, y1, x1), "xz": (y0, i0, x0, y1, i1, x1), "yz": (y0, x0, i0, y1, x1, i1)}[axis_name]
        return lut, np.array([label], np.int32), np.array([size], np.int64), np.array([box], np.int32)

    # ------------------------------------------------------------------ stage 4: paint
    def relabel(self, lut, axis_name, shape3d):
        """Paint final labels (lut [N, stride] int32, 0 = dropped) into a (D,H,W) device volume:
        every voxel is written exactly once, from the runs."""
        D, Hv, Wv = shape3d
        vol = torch.empty(shape3d, dtype=torch.int32, device=self.dev)
        lut_d = torch.from_numpy(np.ascontiguousarray(lut, dtype=np.int32)).to(self.dev)
        strides = {"xy": (Hv * Wv, Wv, 1), "xz": (Wv, Hv * Wv, 1), "yz": (1, Hv * Wv, Wv)}[axis_name]
        r = self.runs
        call("be_runs_paint", ptr(r["row_ptr"]), ptr(r["yx"]), ptr(r["x1"]), ptr(r["cc"]), ptr(r["slice_off"]),
             ptr(lut_d), int(lut_d.shape[1]), 0, 0, self.N, self.h, self.w, ptr(vol), *strides, stream_ptr())
        self.launches += 1
        self._lut_d = lut_d
        return vol

    def tracker_instances(self, axis_name, shape3d, lut, labels, boxes, vol, batch=64):
        """Eager form of `instances_from_dense` (kept for tests)."""
        return instances_from_dense(vol, axis_name, labels, boxes, batch)


def _extract_runs(img, seg_len):
    """Maximal runs of equal non-zero label over flat indices, split at multiples of seg_len.
    Returns device tensors (labels int32, starts int64, lens int32) in raster order."""
    dev = img.device
    n = img.numel()
    chunks = (n + _lib.RUN_CHUNK - 1) // _lib.RUN_CHUNK
    counts = torch.zeros(2 * (chunks + 1), dtype=torch.int32, device=dev)
    st = stream_ptr()
    call("be_runs_count", ptr(img), n, seg_len, ptr(counts), st)
    c2 = counts.view(2, chunks + 1)
    offsets = torch.zeros((2, chunks + 1), dtype=torch.int64, device=dev)
    torch.cumsum(c2[:, :-1], 1, out=offsets[:, 1:])
    total = int(offsets[0, -1].item())
    labels = torch.empty(total, dtype=torch.int32, device=dev)
    starts = torch.empty(total, dtype=torch.int64, device=dev)
    ends = torch.empty(total, dtype=torch.int64, device=dev)
    if total:
        call("be_runs_write", ptr(img), n, seg_len, ptr(offsets), ptr(labels), ptr(starts), ptr(ends),
             total, st)
    return labels, starts, (ends - starts).to(torch.int32)


def instances_from_dense(vol, axis_name, labels, boxes, batch=64):
    """The reference's `InstanceTracker.instances` dictionary (tracker.py:61-123) rebuilt from
    the plane's dense (D,H,W) label volume: per label, runs in arrival order (reverse slice order
    for xy/xz, where runs are per-slice flat-index runs; sorted 3-D runs for yz)."""
    D, Hv, Wv = vol.shape
    dev = vol.device
    if len(labels) == 0:
        return {}
    if axis_name == "yz":
        lab, st3, ln = _extract_runs(vol, vol.numel())
        seq = st3
    else:
        N = D if axis_name == "xy" else Hv
        h, w = (Hv, Wv) if axis_name == "xy" else (D, Wv)
        hw = h * w
        labs, sts, lns = [], [], []
        for s0 in range(0, N, batch):
            s1 = min(N, s0 + batch)
            img = vol[s0:s1] if axis_name == "xy" else vol[:, s0:s1, :].permute(1, 0, 2).contiguous()
            l_, s_, n_ = _extract_runs(img, hw)
            labs.append(l_); sts.append(s_ + s0 * hw); lns.append(n_)
        lab, stg, ln = torch.cat(labs), torch.cat(sts), torch.cat(lns)
        sl = torch.div(stg, hw, rounding_mode="floor")
        s2d = stg - sl * hw
        seq = (N - 1 - sl) * hw + s2d
        if axis_name == "xy":
            st3 = stg
        else:  # xz: (z, x) of slice `sl` -> (z, sl, x); run lengths kept (tracker.py:80-84)
            z = torch.div(s2d, w, rounding_mode="floor")
            st3 = (z * Hv + sl) * Wv + (s2d - z * w)
    labels = np.asarray(labels)
    max_label = int(max(int(labels.max()), int(lab.max().item()) if lab.numel() else 0))
    rank_lut = torch.full((max_label + 1,), -1, dtype=torch.int64, device=dev)
    rank_lut[torch.from_numpy(labels.astype(np.int64)).to(dev)] = torch.arange(len(labels), device=dev)
    rank = rank_lut[lab.long()]
    keys = (rank << 40) | seq
    n = keys.numel()
    idx = torch.arange(n, dtype=torch.int32, device=dev)
    keys_out = torch.empty_like(keys)
    idx_out = torch.empty_like(idx)
    need = _lib.SZ(0)
    _lib.lib().be_sort_runs(None, None, None, None, n, None, 0, need, None)
    temp = torch.empty(max(int(need.value), 1), dtype=torch.uint8, device=dev)
    call("be_sort_runs", ptr(keys), ptr(keys_out), ptr(idx), ptr(idx_out), n, ptr(temp), temp.numel(),
         None, stream_ptr())
    order = idx_out.long()
    st_sorted = st3[order].cpu().numpy()
    ln_sorted = ln[order].long().cpu().numpy()
    counts = torch.bincount(keys_out >> 40, minlength=len(labels)).cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(counts)])
    instances = {}
    for i, lab_i in enumerate(labels):
        a, b = offs[i], offs[i + 1]
        instances[int(lab_i)] = {"box": tuple(int(v) for v in boxes[i]),
                                 "starts": st_sorted[a:b], "runs": ln_sorted[a:b]}
    return instances


class LazyAttrs(dict):
    """Instance attributes whose RLE ('starts', 'runs') is materialised from the device-resident
    label volume on first use. Only `attrs['box']` is answered without materialising; every other
    way of looking at the dictionary (membership tests, get, iteration, keys / items / values,
    len, copy, equality, pickling, json) sees the complete {'box', 'starts', 'runs'} mapping."""

    def __init__(self, box, owner):
        super().__init__(box=box)
        self._owner = owner

    def _fill(self):
        if self._owner is not None:
            owner, self._owner = self._owner, None
            owner.materialize()

    def __missing__(self, key):
        if key in ("starts", "runs") and self._owner is not None:
            self._fill()
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        self._fill()
        return dict.__contains__(self, key)

    def get(self, key, default=None):
        self._fill()
        return dict.get(self, key, default)

    def keys(self):
        self._fill()
        return dict.keys(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def __len__(self):
        self._fill()
        return dict.__len__(self)

    def __eq__(self, other):
        self._fill()
        return dict.__eq__(self, other)

    __hash__ = None

    def copy(self):
        self._fill()
        return dict(self)

    def __repr__(self):
        self._fill()
        return dict.__repr__(self)

    def __reduce__(self):
        self._fill()
        return (dict, (dict(self),))


class LazyPlane:
    """Owner of the deferred RLE extraction of one plane's tracker."""

    def __init__(self, dense, axis_name, labels, boxes):
        self.dense, self.axis_name, self.labels, self.boxes = dense, axis_name, labels, boxes
        self.attrs = {l: LazyAttrs(tuple(b), self) for l, b in zip(np.asarray(labels).tolist(), np.asarray(boxes).tolist())}
        self.done = False

    def materialize(self):
        if self.done:
            return
        self.done = True
        full = instances_from_dense(self.dense, self.axis_name, self.labels, self.boxes)
        for l, a in full.items():
            attrs = self.attrs.get(l)
            if attrs is not None:
                attrs._owner = None
                dict.__setitem__(attrs, "starts", a["starts"])
                dict.__setitem__(attrs, "runs", a["runs"])
