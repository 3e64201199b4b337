"""Tracker-level morphology on the device-resident label volume:
`filters.erode` / `filters.dilate` / `filters.fill_holes_in_segmentation`
(/root/reference/empanada/inference/filters.py:154-210) as `Engine3d.infer_on_axis` applies them
after tracking and the size filters (empanada_napari/inference.py:560-570). Each step decodes the
tracker to a label volume (here: the volume is already in HBM), transforms it, and re-encodes it
with `filters.pan_seg_to_rle_seg` (filters.py:58-118): labels outside the class range dropped,
26-connected components of equal-valued voxels renumbered class * divisor + 1.. in raster order.
Kernels: csrc/morph_kernels.cu; the re-encoding works on the row runs of the volume.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr
from .postproc import LazyPlane, _extract_runs


def _morph(vol, op, iterations):
    D, H, W = (int(v) for v in vol.shape)
    a, b = vol, torch.empty_like(vol)
    for _ in range(int(iterations)):
        call("be_morph3d", ptr(a), ptr(b), D, H, W, int(op), stream_ptr())
        a, b = b, (a if a is not vol else torch.empty_like(vol))
    return a


def _paint_runs(rows, W, row_ptr, run_row, x0, x1, value, dev):
    """Dense (rows, W) int32 image from row runs (every element written)."""
    n = int(x0.numel())
    out = torch.empty((rows, W), dtype=torch.int32, device=dev)
    yx = torch.stack([run_row.to(torch.int32), x0.to(torch.int32)], dim=1).contiguous() if n else torch.zeros((1, 2), dtype=torch.int32, device=dev)
    x1 = x1.to(torch.int32).contiguous() if n else torch.zeros(1, dtype=torch.int32, device=dev)
    value = value.to(torch.int32).contiguous() if n else torch.zeros(1, dtype=torch.int32, device=dev)
    slice_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    call("be_runs_paint", ptr(row_ptr), ptr(yx), ptr(x1), ptr(value), ptr(slice_off), None, 0, 0, 0, 1, rows, W,
         ptr(out), 0, W, 1, stream_ptr())
    return out


def reencode(vol, class_id, label_divisor, is_thing):
    """`filters.pan_seg_to_rle_seg(mask, [class_id], label_divisor, thing_list)` on a (D,H,W)
    int32 device volume. Returns (new dense volume, labels int64[n] ascending, sizes int64[n],
    boxes int64[n, 6])."""
    dev = vol.device
    D, H, W = (int(v) for v in vol.shape)
    lo, hi = class_id * label_divisor, (class_id + 1) * label_divisor
    work = vol.clone()
    call("be_range_keep", ptr(work), work.numel(), int(lo), int(hi), stream_ptr())
    val, start, length = _extract_runs(work, W)                      # row runs in raster order
    n = int(val.numel())
    rows = D * H
    if n == 0:
        return torch.zeros_like(vol), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros((0, 6), np.int64)
    end = (start + length.long()).contiguous()
    row_ptr = torch.empty(rows + 1, dtype=torch.int32, device=dev)
    L = torch.arange(n, dtype=torch.int32, device=dev)
    root = torch.empty(n, dtype=torch.int32, device=dev)
    call("be_runs3d_cc", ptr(start), ptr(end), ptr(val), n, D, H, W, ptr(row_ptr), ptr(L), ptr(root), stream_ptr())
    run_row = torch.div(start, W, rounding_mode="floor")
    x0 = start - run_row * W
    x1 = x0 + length.long()
    if is_thing:
        # components in raster order of their first voxel = ascending root run index
        is_root = root.long() == torch.arange(n, device=dev)
        rank = torch.cumsum(is_root.to(torch.int64), 0)               # 1-based id at the roots
        comp = rank[root.long()]
        new_val = comp + lo
        n_out = int(rank[-1].item())
        labels = np.arange(1, n_out + 1, dtype=np.int64) + lo
        group = comp - 1
    else:
        # a stuff class keeps its label values (no connected components, filters.py:105-107)
        uniq, group = torch.unique(val.long(), return_inverse=True)
        new_val = val.long()
        n_out = int(uniq.numel())
        labels = uniq.cpu().numpy().astype(np.int64)
    z = torch.div(run_row, H, rounding_mode="floor")
    y = run_row - z * H
    sizes = torch.zeros(n_out, dtype=torch.int64, device=dev).index_add_(0, group, length.long())
    big = torch.iinfo(torch.int64).max
    mins = torch.full((n_out, 3), big, dtype=torch.int64, device=dev)
    maxs = torch.full((n_out, 3), -1, dtype=torch.int64, device=dev)
    lo3 = torch.stack([z, y, x0], dim=1)
    hi3 = torch.stack([z + 1, y + 1, x1], dim=1)
    idx = group[:, None].expand(-1, 3)
    mins.scatter_reduce_(0, idx, lo3, reduce="amin")
    maxs.scatter_reduce_(0, idx, hi3, reduce="amax")
    dense = _paint_runs(rows, W, row_ptr, run_row, x0, x1, new_val, dev).view(D, H, W)
    return dense, labels, sizes.cpu().numpy(), torch.cat([mins, maxs], dim=1).cpu().numpy()


def fill_holes(vol):
    """`fill_holes_in_segmentation` (filters.py:174-210) on a (D,H,W) int32 device volume, in place
    on a copy: slice by slice along axis 0, labels in ascending order, each inside its bounding
    box of the unmodified slice."""
    dev = vol.device
    D, H, W = (int(v) for v in vol.shape)
    out = vol.clone()
    val, start, length = _extract_runs(vol, W)
    if int(val.numel()) == 0:
        return out
    row = torch.div(start, W, rounding_mode="floor")
    x0 = start - row * W
    x1 = x0 + length.long()
    z = torch.div(row, H, rounding_mode="floor")
    y = row - z * H
    span = int(val.max().item()) + 1
    key = z * span + val.long()                                       # ascending: slice, then label
    uniq, group = torch.unique(key, return_inverse=True)
    m = int(uniq.numel())
    big = torch.iinfo(torch.int64).max
    mins = torch.full((m, 2), big, dtype=torch.int64, device=dev)
    maxs = torch.full((m, 2), -1, dtype=torch.int64, device=dev)
    idx = group[:, None].expand(-1, 2)
    mins.scatter_reduce_(0, idx, torch.stack([y, x0], dim=1), reduce="amin")
    maxs.scatter_reduce_(0, idx, torch.stack([y + 1, x1], dim=1), reduce="amax")
    zs = torch.div(uniq, span, rounding_mode="floor")
    labels = (uniq - zs * span).to(torch.int32).contiguous()
    boxes = torch.cat([mins, maxs], dim=1).to(torch.int32).contiguous()
    slice_off = torch.searchsorted(zs.contiguous(), torch.arange(D + 1, device=dev)).to(torch.int32).contiguous()
    scratch = torch.empty((D, H, W), dtype=torch.uint8, device=dev)
    call("be_fill_holes", ptr(out), ptr(scratch), D, H, W, ptr(slice_off), ptr(labels), ptr(boxes), stream_ptr())
    return out


def apply_to_tracker(tracker, dense, class_id, label_divisor, is_thing, erosion=0, dilation=0, holes=False):
    """The morphology block of `Engine3d.infer_on_axis`: erode, dilate, fill holes (each followed
    by the re-encoding). Replaces the tracker's instances / label volume / sizes and returns the
    new dense volume."""
    steps = []
    if erosion > 0:
        steps.append(lambda v: _morph(v, 0, erosion))
    if dilation > 0:
        steps.append(lambda v: _morph(v, 1, dilation))
    if holes:
        steps.append(fill_holes)
    for step in steps:
        dense, labels, sizes, boxes = reencode(step(dense), class_id, label_divisor, is_thing)
        plane = LazyPlane(dense, "yz", labels, boxes)      # "yz": maximal flat runs sorted by start = rle_encode of raster coords
        tracker.instances = plane.attrs
        tracker._b200_dense = dense
        tracker._b200_sizes = {int(l): int(s) for l, s in zip(labels, sizes)}
    return dense
