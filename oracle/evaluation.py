"""CPU restatement of the reference's evaluation of 3-D run-length-encoded segmentations:
`Evaluator` (empanada/evaluation/evaluator.py:24-122), `f1` / `f1_50` / `f1_75`
(evaluation/instance_metrics.py:3-54,231-235) and semantic `iou`
(evaluation/semantic_metrics.py:4-26). TEST INFRASTRUCTURE ONLY (see oracle/post.py header):
the end-to-end agreement gate of BASELINE.md section 4 scores the CUDA path's trackers against
the fp32 oracle's trackers with these functions. Pinned by tests/golden/eval_cases.json, produced
by oracle/make_golden.py from the UNMODIFIED reference Evaluator.
"""
import json

import numpy as np

from .ranges import merge_rles, rle_iou, string_to_rle
from .tracking import rle_matcher


def f1(gt_matched, gt_unmatched, pred_matched, pred_unmatched, matched_ious, iou_thr=0.5):
    """instance_metrics.py:3-54."""
    fn = len(gt_unmatched)
    fp = len(pred_unmatched)
    tp = np.count_nonzero(matched_ious >= iou_thr)
    failed = np.count_nonzero(matched_ious < iou_thr)
    fp += failed
    fn += failed
    if tp + fp + fn == 0:
        return 1
    return tp / (tp + 0.5 * fp + 0.5 * fn)


def f1_50(**kw):
    return f1(**kw, iou_thr=0.5)


def f1_75(**kw):
    return f1(**kw, iou_thr=0.75)


def iou(gt_rle, pred_rle):
    """semantic_metrics.py:4-26."""
    if len(gt_rle) == 0 and len(pred_rle) == 0:
        return 1
    if len(gt_rle) == 0 or len(pred_rle) == 0:
        return 0
    return rle_iou(gt_rle[:, 0], gt_rle[:, 1], pred_rle[:, 0], pred_rle[:, 1])


def _merge_encodings_for_semantic(pred_encodings):
    """evaluator.py:6-22 (including its single-instance quirk: one prediction -> [[-1, -1]])."""
    if len(pred_encodings) > 1:
        pred_runs = np.concatenate([np.stack(string_to_rle(enc), axis=1) for enc in pred_encodings])
        return np.stack(merge_rles(pred_runs[:, 0], pred_runs[:, 1]), axis=1)
    return np.array([[-1, -1]])


def _instances_from_json(d):
    """json instance dicts hold 'box' + 'rle' strings (tracker.py:125-147); `unpack_rle_attrs`
    (rle.py:120-150) decodes them for the matcher."""
    out = {}
    for k, v in d.items():
        if "rle" in v:
            starts, runs = string_to_rle(v["rle"])
        else:
            starts, runs = v["starts"], v["runs"]
        out[k] = {"box": v["box"], "starts": np.asarray(starts), "runs": np.asarray(runs)}
    return out


class Evaluator:
    """evaluator.py:24-122."""

    def __init__(self, semantic_metrics=None, instance_metrics=None, panoptic_metrics=None):
        self.semantic_metrics = semantic_metrics
        self.instance_metrics = instance_metrics
        self.panoptic_metrics = panoptic_metrics

    def __call__(self, gt_json_fpath, pred_json_fpath, return_instances=False):
        with open(gt_json_fpath, mode="r") as f:
            gt_json = json.load(f)
        with open(pred_json_fpath, mode="r") as f:
            pred_json = json.load(f)
        assert gt_json["class_id"] == pred_json["class_id"], "Prediction and ground truth classes must match!"
        gt_encodings = [v["rle"] for v in gt_json["instances"].values()]
        pred_encodings = [v["rle"] for v in pred_json["instances"].values()]
        results = {}
        if self.semantic_metrics is not None:
            gt_indices = np.concatenate([np.stack(string_to_rle(enc), axis=1) for enc in gt_encodings])
            pred_indices = _merge_encodings_for_semantic(pred_encodings)
            results.update({name: func(gt_indices, pred_indices) for name, func in self.semantic_metrics.items()})
        inst = {}
        if self.instance_metrics is not None or self.panoptic_metrics is not None:
            (gt_matched, pred_matched), (gt_labels, pred_labels), matched_ious, _ = rle_matcher(
                _instances_from_json(gt_json["instances"]), _instances_from_json(pred_json["instances"]), 0.5)
            inst = {"gt_matched": gt_matched, "pred_matched": pred_matched,
                    "gt_unmatched": np.setdiff1d(gt_labels, gt_matched),
                    "pred_unmatched": np.setdiff1d(pred_labels, pred_matched), "matched_ious": matched_ious}
            for metrics in (self.instance_metrics, self.panoptic_metrics):
                if metrics is not None:
                    results.update({name: func(**inst) for name, func in metrics.items()})
        return (results, inst) if return_instances else results
