"""CPU restatement of the reference's orchestration: Engine3d.infer_on_axis
(empanada_napari/inference.py:491-578), Engine2d.infer (:281-325, no-tiling branch) and the
model-facing half of the render engines (empanada/inference/engines.py:300-394).
TEST INFRASTRUCTURE ONLY (see oracle/post.py header).

`heads_fn(index, image_f32_padded)` stands for `model(image, render_steps, interpolate_ins)`:
it returns (sem_logits (C,H,W), ctr_hmp (h4,w4), offsets (2,h4,w4)) as fp32 numpy arrays.
"""
import numpy as np

from . import post
from .tracking import (InstanceTracker, RLEMatcher, backward_matching, connected_components,
                       forward_matching, remove_pancakes, remove_small_objects)
from .ranges import numpy_fill_instances

AXES = {"xy": 0, "xz": 1, "yz": 2}


def take_slice(volume, idx, axis):
    """empanada/array_utils.py:10-27."""
    sl = [slice(None)] * volume.ndim
    sl[axis] = idx
    return volume[tuple(sl)]


def infer_on_axis(volume, axis_name, heads_fn, model_config, label_divisor=1000,
                  median_kernel_size=3, stuff_area=64, void_label=0, nms_threshold=0.1,
                  nms_kernel=3, confidence_thr=0.5, min_size=500, min_extent=4,
                  save_panoptic=True, dtype=np.int32, fine_boundaries=False, semantic_only=False,
                  inference_scale=1, label_erosion=0, label_dilation=0, fill_holes_in_segmentation=False):
    """`fine_boundaries=True`: `heads_fn` returns FULL-resolution ctr_hmp / offsets (the model's
    `interpolate_ins=True` output) and pixels are grouped with step 1 (engines.py:263-275).
    `semantic_only`: engine thing_list = [] (inference.py:365-368). `inference_scale` s: slices go
    through `resize_by_factor` (volume_dataset.py:42-47) and `heads_fn` stands for the model with
    2 + log2(s) render steps: sem_logits at s times the padded down-sampled size."""
    from .transforms import resize_by_factor
    axis = AXES[axis_name]
    labels = model_config["labels"]
    thing_list = [] if semantic_only else model_config["thing_list"]
    pf = model_config["padding_factor"]
    norms = model_config["norms"]
    eng = post.RenderEnginePost(thing_list, label_divisor, stuff_area, void_label, nms_threshold,
                                nms_kernel, confidence_thr, median_kernel_size, not fine_boundaries)
    trackers = [InstanceTracker(l, label_divisor, volume.shape, axis_name) for l in labels]
    matchers = [RLEMatcher(c, label_divisor, 0.25, 0.25) for c in thing_list]
    pan_segs = []
    n = volume.shape[axis]
    for i in range(n):
        img = take_slice(volume, i, axis)
        h, w = img.shape
        img = resize_by_factor(img, inference_scale)
        x = post.factor_pad(post.normalize(img, norms["mean"], norms["std"]), pf)
        sem_logits, ctr, off = heads_fn(i, x)
        sem = post.sigmoid(sem_logits) if sem_logits.shape[0] == 1 else _softmax(sem_logits)
        pan_segs.append(eng(sem, ctr, off, (h, w), inference_scale))
    pan_segs.extend(eng.end(inference_scale))
    rle_stack = forward_matching(pan_segs, matchers, labels, label_divisor, thing_list)
    for index, rle_seg in backward_matching(rle_stack, matchers, n):
        for tr in trackers:
            tr.update(rle_seg[tr.class_id], index)
    for tr in trackers:
        tr.finish()
        remove_small_objects(tr, min_size=min_size)
        remove_pancakes(tr, min_span=min_extent)
    # tracker morphology (empanada_napari/inference.py:560-570)
    from . import tracking as _t
    if label_erosion > 0:
        for tr in trackers:
            _t.erode(tr, volume.shape, labels, label_divisor, thing_list, iterations=label_erosion)
    if label_dilation > 0:
        for tr in trackers:
            _t.dilate(tr, volume.shape, labels, label_divisor, thing_list, iterations=label_dilation)
    if fill_holes_in_segmentation:
        for tr in trackers:
            _t.fill_holes_in_segmentation(tr, volume.shape, labels, label_divisor, thing_list)
    stack = None
    if save_panoptic:
        stack = np.zeros(volume.shape, dtype=dtype)
        for tr in trackers:
            numpy_fill_instances(stack, tr.instances)
    return stack, trackers


def _softmax(x):
    e = np.exp(x - x.max(axis=0, keepdims=True))
    return (e / e.sum(axis=0, keepdims=True)).astype(np.float32)


def engine2d_infer(image, heads_fn, model_config, label_divisor=1000, nms_threshold=0.1,
                   nms_kernel=3, confidence_thr=0.3, stuff_area=64, void_label=0, fine_boundaries=False,
                   semantic_only=False, inference_scale=1):
    """Engine2d.infer, no tiling (empanada_napari/inference.py:319-325,263-279)."""
    from .transforms import resize_by_factor
    thing_list = [] if semantic_only else model_config["thing_list"]
    norms = model_config["norms"]
    h, w = image.shape
    image = resize_by_factor(image, inference_scale)
    x = post.factor_pad(post.normalize(image, norms["mean"], norms["std"]), model_config["padding_factor"])
    sem_logits, ctr, off = heads_fn(0, x)
    sem = post.sigmoid(sem_logits) if sem_logits.shape[0] == 1 else _softmax(sem_logits)
    eng = post.RenderEnginePost(thing_list, label_divisor, stuff_area, void_label, nms_threshold,
                                nms_kernel, confidence_thr, None, not fine_boundaries)
    pan = eng(sem, ctr, off, (h, w), inference_scale).astype(np.int32)
    for label in thing_list:
        lo = label * label_divisor
        hi = lo + label_divisor
        ins = pan.copy()
        ins[np.logical_or(pan < lo, pan >= hi)] = 0
        ins = connected_components(ins).astype(np.int32)
        ins[ins > 0] += lo
        pan[ins > 0] = ins[ins > 0]
    return pan
